// runtime.h — process-wide enforcement state behind the exported hooks: configuration from the plugin's env
// contract, the shared region, the process-local allocation table and the semantic implementation of every
// intercepted driver call. hook.cc only adapts exported symbol names onto these methods; cabi.cc exposes the
// same objects to host tools.
//
// Reference counterparts: libvgpu.c (cuInit@0x15f5a, preInit@0x15c14, postInit@0x15d63), allocator.c
// (add_chunk@0x40007, remove_chunk@0x40871, check_memory_type@0x407f2), memory.c (cuMemAlloc_v2@0x3180a ...
// cuLaunchKernel@0x370bf), utils.c (set_task_pid@0x16a7f, try_lock_unified_lock@0x1612c).
#pragma once
#include <cuda.h>

#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "region.h"

namespace vgpu {

class SwapEngine;
class Limiter;

constexpr size_t kIpcSize = 2u << 20;  // IPCSIZE .data@0x610a8: allocations above this take the swap switch

enum class AllocKind : uint8_t { Device, Managed, Pitch, Swap, Async };

struct Alloc {
    size_t size;       // bytes charged to the quota (requested bytes, no rounding — Appendix E)
    int dev;           // CUDA ordinal it was charged to
    AllocKind kind;
};

struct Config {
    bool oversubscribe = false;   // CUDA_OVERSUBSCRIBE=true  (server.go:356-358)
    int util_policy = 0;          // GPU_CORE_UTILIZATION_POLICY: 0 default, 1 force, 2 disable (get_utilization_switch@0x45307)
    bool active_oom_killer = false;
    int priority = 1;             // CUDA_TASK_PRIORITY
    std::string region_path;      // CUDA_DEVICE_MEMORY_SHARED_CACHE, default /tmp/cudevshr.cache
    uint64_t mem_limit[VGPU_MAX_DEVICES] = {};
    uint64_t sm_limit[VGPU_MAX_DEVICES] = {};
    uint64_t virtual_limit[VGPU_MAX_DEVICES] = {};  // swap mode only: CUDA_DEVICE_MEMORY_VIRTUAL_LIMIT[_i]; 0 = host pool bound
    // swap mode only. false (default): the gpumem limit bounds RESIDENT bytes, live bytes may exceed it (SURVEY.md §8d
    // cfg 3). true (VGPU_SWAP_LIMIT_MODE=virtual): the reference's meaning — the limit is a hard cap on live bytes
    // (oom_check still applies under CUDA_OVERSUBSCRIBE), and paging only starts when the DEVICE runs short
    bool limit_is_virtual = false;
    static Config from_env();
};

class Runtime {
   public:
    static Runtime &get();

    // region + slot; safe before cuInit and in non-CUDA processes (nothing GPU-side is touched)
    bool ensure_initialized();
    bool initialized() const { return inited_.load(std::memory_order_acquire); }

    // ---- intercepted calls (semantics + reference return codes; see hook.cc for the symbol mapping)
    CUresult init(unsigned flags);                                        // cuInit
    CUresult mem_alloc(CUdeviceptr *dptr, size_t bytes);                  // cuMemAlloc_v2
    CUresult mem_alloc_managed(CUdeviceptr *dptr, size_t bytes, unsigned flags);
    CUresult mem_alloc_pitch(CUdeviceptr *dptr, size_t *pitch, size_t width, size_t height, unsigned elem);
    CUresult mem_free(CUdeviceptr dptr);                                  // cuMemFree_v2
    CUresult mem_get_info(size_t *free_b, size_t *total_b);               // cuMemGetInfo_v2
    CUresult device_total_mem(size_t *bytes, CUdevice dev);               // cuDeviceTotalMem_v2
    CUresult primary_ctx_retain(CUcontext *ctx, CUdevice dev);            // cuDevicePrimaryCtxRetain
    CUresult ctx_create(CUcontext *ctx, unsigned flags, CUdevice dev);    // cuCtxCreate_v2
    CUresult launch_kernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                           unsigned smem, CUstream st, void **params, void **extra);
    CUresult launch_kernel_ex(const CUlaunchConfig *cfg, CUfunction f, void **params, void **extra);
    CUresult launch_cooperative(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                unsigned bz, unsigned smem, CUstream st, void **params);
    // per-thread-default-stream twins of the three launches above (ptsz: a null stream means CU_STREAM_PER_THREAD)
    CUresult launch_kernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                unsigned smem, CUstream st, void **params, void **extra);
    CUresult launch_kernel_ex_ptsz(const CUlaunchConfig *cfg, CUfunction f, void **params, void **extra);
    CUresult launch_cooperative_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                     unsigned bz, unsigned smem, CUstream st, void **params);
    // ---- coverage beyond the reference: the reference forwards these unaccounted / unlimited
    // (cuMemAllocAsync@0x37e52, cuMemCreate@0x37ba1, cuGraphLaunch); VGPU_REFERENCE_COVERAGE=1 restores that.
    CUresult mem_alloc_async(CUdeviceptr *dptr, size_t bytes, CUmemoryPool pool, bool from_pool, CUstream st, bool ptsz);
    CUresult mem_free_async(CUdeviceptr dptr, CUstream st, bool ptsz);
    CUresult mem_create(CUmemGenericAllocationHandle *h, size_t bytes, const CUmemAllocationProp *prop, unsigned long long flags);
    CUresult mem_release(CUmemGenericAllocationHandle h);
    CUresult mem_map(CUdeviceptr ptr, size_t size, size_t offset, CUmemGenericAllocationHandle h, unsigned long long flags);
    CUresult mem_unmap(CUdeviceptr ptr, size_t size);
    CUresult graph_launch(CUgraphExec g, CUstream st, bool ptsz);
    // check_oom() of the reference (oom_check(dev, 0)): used by host-alloc style hooks
    // cuModuleUnload: CUfunction handles of the module die with it; forget their cached parameter layouts
    void forget_function_layouts();
    bool check_oom();
    // memcpy/memset family: make the touched device ranges resident before the real call (swap mode only)
    // a device range a copy / fill / query is about to touch; `writes` = the operation writes it (a range that is only
    // read stays clean: evicting it later needs no copy)
    void touch_range(CUdeviceptr p, size_t bytes, CUstream st, bool writes = true);
    void touch_range2(CUdeviceptr dst, size_t dbytes, CUdeviceptr src, size_t sbytes, CUstream st);   // dst written, src read
    // cuMemAdvise / cuMemPrefetchAsync on a swappable range (the reference's swappable memory is UVM-managed, where both
    // are meaningful): true when the pointer is the swap engine's and the call has been handled
    bool swap_advise(CUdeviceptr p, CUmem_advise advice);
    bool swap_prefetch(CUdeviceptr p, bool to_device);
    // Explicitly built graphs (cuGraphAdd*Node / *SetParams): a replay cannot be admitted node by node, so the swappable
    // operands named by a node are made resident when the node is defined and stay pinned (same rule as stream capture).
    CUresult pin_graph_kernel(CUfunction f, void **params, void **extra);
    CUresult pin_graph_ptrs(const CUdeviceptr *p, size_t n);
    SwapEngine *engine_of(CUdeviceptr p);         // the engine (of whichever device) whose arena holds p, or null
    bool swap_address_range(CUdeviceptr p, CUdeviceptr *base, size_t *size);   // true: p is a swappable buffer, answered from the table
    void touch_done(CUstream st);   // after the real copy has been enqueued: unpins + records the use
    // Batched copies (cuMemcpyBatchAsync / cuMemcpy3DBatchAsync, CUDA 12.8): every swappable operand of the batch is admitted
    // together. false = they do not fit the resident quota at once (the caller then issues the copies one by one).
    bool touch_batch(const CUdeviceptr *written, size_t nw, const CUdeviceptr *read, size_t nr, CUstream st);

    // NVML view: nvmlDeviceGetMemoryInfo under the quota (nvml/hook.c:L327-334)
    bool nvml_memory_view(int nvml_index, unsigned long long *total, unsigned long long *free_b, unsigned long long *used);

    // ---- introspection (C-ABI, tests)
    Region *region() { return region_.get(); }
    const Config &config() const { return cfg_; }
    bool reference_coverage_mode() const;
    int check_memory_type(CUdeviceptr p);   // 2 = tracked device memory, 1 = not (check_memory_type@0x407f2)
    size_t table_size();
    uint64_t context_size() const { return context_size_; }
    SwapEngine *swap(int dev);
    Limiter *limiter() { return limiter_.get(); }
    void on_fork_child();
    void on_exit();

   private:
    Runtime() = default;
    int current_device();
    void post_init(bool late);              // postInit@0x15d63
    void measure_context_size();            // set_task_pid@0x16a7f
    void start_memory_monitor(int sm_limit_percent);   // set_gpu_device_memory_monitor@0x42301 + active_oom_killer@0x41e55
    void wait_running();                    // wait_status_self(1) loop of every wrapper
    bool track(CUdeviceptr base, size_t size, int dev, AllocKind kind);
    CUresult swap_alloc(CUdeviceptr *dptr, size_t bytes, int dev);
    bool charge(int dev, size_t bytes);     // quota check + accounting of a NON-swappable allocation (both modes)
    void uncharge(int dev, size_t bytes);

    std::atomic<bool> inited_{false};
    std::atomic<bool> post_inited_{false};
    std::mutex init_mu_;
    Config cfg_;
    std::unique_ptr<Region> region_;
    bool fail_closed_ = false;                   // a quota is configured but the region could not be opened: refuse device allocations
    int32_t pid_ = 0;
    uint64_t context_size_ = 0;
    bool pid_found_ = false;
    bool ctx_charged_[VGPU_MAX_DEVICES] = {};     // context_size added to this device's lane (once per device, see charge_context_once)
    std::mutex ctx_mu_;
    void charge_context_once(CUdevice dev);

    std::mutex table_mu_;                   // the reference's single allocator mutex (mutex@0x61180)
    std::map<CUdeviceptr, Alloc> table_;    // base -> alloc; ordered for range classification
    // cuMemCreate handles charged to the quota (same mutex). The physical memory behind a handle lives until the handle is
    // released AND its last mapping is gone (the CUDA samples release right after cuMemMap), so the charge does too.
    struct PhysAlloc { size_t size; int dev; int maps; bool released; };
    std::map<CUmemGenericAllocationHandle, PhysAlloc> phys_;
    std::map<CUdeviceptr, std::pair<size_t, CUmemGenericAllocationHandle>> vmaps_;   // application cuMemMap ranges onto charged handles

    std::mutex swap_mu_;
    std::unique_ptr<SwapEngine> swap_[VGPU_MAX_DEVICES];
    std::unique_ptr<Limiter> limiter_;
};

}  // namespace vgpu
