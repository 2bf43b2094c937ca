// swap.h — the virtual-device-memory swap engine (one per device/context).
//
// Behind the same switch as the reference's oversubscription (CUDA_OVERSUBSCRIBE=true -> allocmode 0,
// postInit libvgpu.so@0x15d63; large allocations then go through cuMemoryAllocate@0x315da), but where the
// reference hands the problem to CUDA UVM (cuMemAllocManaged; victim choice and page traffic inside the NVIDIA
// kernel driver on host cores), this engine keeps explicit control:
//   * every swappable allocation is a stable virtual range (cuMemAddressReserve arena) backed by VMM physical
//     handles that are mapped only while the buffer is resident;
//   * a device-resident allocation table (32-byte rows) records residency and a logical LRU clock;
//   * when an admission (kernel launch / memcpy / new allocation) needs more physical memory than the container's
//     resident quota allows, a GPU victim scan picks exact-LRU victims, a TMA pack kernel compacts them into a
//     staging ring in HBM, and a copy stream drains the ring to pinned host memory while the next chunk is packed;
//     page-in is the mirror image (pinned host -> staging ring -> unpack kernel into the re-mapped range);
//   * pack, unpack, scan, D2H and H2D each have their own stream, ordered only by the events that express real
//     data dependencies, so the two link directions and the HBM compaction all overlap;
//   * nothing on the host ever waits for a PCIe transfer except for ring back-pressure: the application stream is
//     ordered behind the page-in with events.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "kernels.h"
#include "vgpu_region.h"
#include "kmod.h"

namespace vgpu {

struct SwapConfig {
    uint64_t resident_cap = 0;            // bytes of physical backing the container may hold (its gpumem quota); 0 = device free memory
    uint64_t virtual_cap = 0;             // live swappable bytes allowed; 0 = bounded by host pool only
    uint64_t host_pool_cap = 0;           // pinned bytes allowed; 0 = unbounded
    size_t chunk_bytes = 32u << 20;       // staging slot size (one DMA)
    int ring_slots = 4;                   // staging slots per direction
    size_t slab_bytes = 1ull << 30;       // pinned pool growth unit
    uint64_t arena_bytes = 1ull << 40;    // virtual address arena
    bool profile = false;                 // bracket pack/unpack launches with events (bench roofline)
    uint32_t scan_lookahead = 8;          // a scan selects this many times the bytes needed; the surplus is consumed by later evictions
    uint32_t trace = 0;                   // VGPU_SWAP_TRACE=n: timeline of n steady-state misses (diagnostics)
    bool async_unmap = false;             // VGPU_SWAP_ASYNC_UNMAP=1: victims are unmapped by a reaper thread instead of the admitting
                                          // thread. Measured neutral on B200/driver 580 (VMM calls from two threads serialise in the driver:
                                          // 68 vs 70 GB/s, profiles/README.md), kept as the building block of the prefetch pipeline
    uint64_t spare_bytes = 128u << 20;    // physical memory the engine may hold beyond the quota while victims await their unmap
    static SwapConfig from_env(uint64_t resident_cap, uint64_t virtual_cap);
};

struct SwapStats {
    uint64_t page_out_bytes = 0, page_in_bytes = 0;
    uint64_t evictions = 0, faults = 0, admissions = 0;
    uint64_t pack_launches = 0, unpack_launches = 0, scan_launches = 0, scans = 0, scan_cache_hits = 0;
    uint64_t resident_bytes = 0, live_bytes = 0, host_bytes = 0, entries = 0;
    uint64_t phys_creates = 0, phys_reuses = 0;
    uint64_t host_slabs = 0, host_slabs_local = 0;   // pinned slabs allocated / of those on the GPU's NUMA node
    double pack_ms = 0, unpack_ms = 0;     // CUDA-event brackets around each launch, only when profiling (include the
                                           // host's event->launch gap when the stream is idle)
    double pack_span_ms = 0, unpack_span_ms = 0;   // exact execution spans from in-kernel %globaltimer stamps (profiling)
    // where the calling thread's time goes inside ensure_resident/alloc (ns): victim scan incl. its sync, waiting for
    // the last pack of a batch, VMM calls (unmap/map/setaccess/create), staging-ring back-pressure, whole admissions
    uint64_t host_scan_ns = 0, host_packsync_ns = 0, host_vmm_ns = 0, host_ring_ns = 0, host_admit_ns = 0;
    uint64_t pack_bytes = 0, unpack_bytes = 0;
};

class SwapEngine {
   public:
    // Must be called with the target context current. Returns nullptr (and logs) on failure.
    static SwapEngine *create(int dev, const SwapConfig &cfg);
    ~SwapEngine();

    CUresult alloc(CUdeviceptr *dptr, size_t bytes);
    CUresult free(CUdeviceptr dptr);            // CUDA_ERROR_INVALID_VALUE when not one of ours
    bool owns(CUdeviceptr p) const { return p >= arena_ && p < arena_ + cfg_.arena_bytes; }
    int lookup(CUdeviceptr p) const;            // row index of the allocation containing p, or -1

    // Makes every listed row resident and orders `stream` after the page-ins. Rows stay pinned (not evictable)
    // until note_use() is called with the same list after the real launch has been enqueued.
    // stream == kHostWait: block the calling thread until the rows are usable instead of ordering a stream behind the
    // page-ins (stream capture: a capturing stream may not wait on outside work). Rows acquired this way and never
    // passed to note_use() stay pinned resident — what a captured kernel's operands need.
    static inline CUstream kHostWait = reinterpret_cast<CUstream>(~uintptr_t(0));
    CUresult ensure_resident(const int *rows, int n, CUstream stream);
    void note_use(const int *rows, int n, CUstream stream);
    // Scans kernel parameter bytes for pointers into the arena; appends distinct row indices.
    void collect_rows(const void *param, size_t bytes, std::vector<int> *rows) const;

    SwapStats stats();
    void set_profile(bool on) { cfg_.profile = on; }
    // the quota left for swappable memory shrinks/grows with the container's non-swappable bytes (context, small buffers)
    void set_resident_cap(uint64_t cap) {
        std::lock_guard<std::mutex> g(mu_);
        quota_cap_ = cap;
        // under physical pressure (see map_row) the working cap stays at what the device could actually give
        cfg_.resident_cap = pressure_ ? std::min(cap, cfg_.resident_cap) : cap;
    }
    // publish the counters into the container's shared region (vgpu_region.h extension block) after every call that
    // changes them; nullptr = do not publish
    void set_shared_record(vgpu_swap_record_t *rec) { std::lock_guard<std::mutex> g(mu_); shared_ = rec; publish_locked(); }
    CUresult drain();                          // wait for all side-stream work (tests / shutdown)
    const SwapConfig &config() const { return cfg_; }
    // device memory the engine itself holds next to the application's resident buffers: both staging rings (the table,
    // scan scratch and span words are a few hundred KiB and not counted). The hook takes it out of the room it gives
    // the engine, so that the container's PHYSICAL footprint stays within its gpumem quota.
    uint64_t device_overhead() const { return 2ull * cfg_.ring_slots * cfg_.chunk_bytes; }
    uint64_t live_bytes() const { return live_bytes_; }

    // diagnostics (VGPU_SWAP_TRACE=n): host timestamps and timed GPU events of the first n missing admissions after
    // warm-up, printed as JSON lines by dump_trace() — where do the two DMA queues idle?
    void dump_trace(FILE *f);

    // test hook: copy of the host mirror of the table
    std::vector<VgpuEntry> snapshot_table();

   private:
    struct Side {                // host-only companion of a table row
        size_t mapped = 0;
        CUmemGenericAllocationHandle handle = 0;
        bool has_handle = false;
        uint64_t host_off = 0;   // pinned pool location while paged out
        bool has_host = false;
        CUevent ready = nullptr; // pending page-in completion (owned by ready_pool_)
        uint64_t use_seq = 0;    // sequence number of the last-use event
        int pins = 0;
        uint64_t va_off = 0;
        bool evicting = false;   // packed and logically paged out, but its unmap is still queued at the reaper
        int out_slot = -1;       // staging slot of the last page-out chunk of this row ...
        uint64_t out_seq = 0;    // ... and that slot's use counter at the time (stale => the D2H is known complete)
    };
    struct Slab { unsigned char *host = nullptr; size_t bytes = 0; std::map<uint64_t, uint64_t> free; };
    struct Slot { CUdeviceptr buf = 0; CUevent busy = nullptr; bool used = false; uint64_t seq = 0; };
    struct PendingHost { uint64_t off, len; CUevent done; };

    SwapEngine() = default;
    bool init(int dev, const SwapConfig &cfg);
    int new_row();
    void mark_dirty(int row);
    CUresult sync_table(CUstream s);
    CUresult make_room(uint64_t need_mapped, bool finish = true);
    CUresult page_out(const std::vector<uint32_t> &victims, bool finish = true);
    CUresult page_out_finish();
    struct InRun { unsigned char *src; uint64_t pos, len; };
    struct InJob { Slot *slot = nullptr; std::vector<PackSegment> segs; std::vector<InRun> runs; std::vector<int> done_rows; uint64_t bytes = 0; };
    CUresult page_in_plan(const std::vector<int> &rows);
    CUresult page_in_stage(const std::vector<int> &rows);
    CUresult page_in_finish(const std::vector<int> &rows);
    CUresult in_issue_copies(InJob &j);
    CUresult map_row(int row);
    void unmap_row(int row);
    CUresult get_phys(size_t mapped, CUmemGenericAllocationHandle *h);
    void trim_phys_pool(uint64_t need);
    bool host_alloc(size_t bytes, uint64_t *off);
    void release_host_range(uint64_t off, uint64_t len);
    void reap_pending_host(bool wait);
    CUevent get_event();
    unsigned char *host_ptr(uint64_t off);
    bool va_alloc(size_t bytes, uint64_t *off);
    void va_free(uint64_t off, size_t bytes);
    CUevent use_event(uint64_t seq);
    Slot &acquire_slot(std::vector<Slot> &ring, int *cursor);
    void prof_begin(CUstream s, CUevent *a);
    void prof_end(CUstream s, CUevent a, bool unpack, uint64_t bytes);
    void harvest_prof(bool wait);

    mutable std::mutex mu_;
    int dev_ = 0;
    int numa_node_ = -1;                            // NUMA node of the GPU (pinned slabs are allocated there)
    SwapConfig cfg_;
    const Kernels *k_ = nullptr;
    size_t gran_ = 2u << 20;
    CUdeviceptr arena_ = 0;
    std::map<uint64_t, uint64_t> va_free_;          // offset -> len
    std::vector<int32_t> owner_;                    // granule -> row (or -1)

    std::vector<VgpuEntry> rows_;                   // host mirror (authoritative)
    std::vector<Side> side_;
    std::vector<int> free_rows_;
    CUdeviceptr d_tbl_ = 0;
    uint32_t tbl_cap_ = 0;
    uint32_t dirty_lo_ = UINT32_MAX, dirty_hi_ = 0;
    VgpuEntry *h_tbl_stage_ = nullptr;              // pinned upload buffer
    uint64_t tick_ = 0;

    uint64_t resident_mapped_ = 0, live_bytes_ = 0, host_used_ = 0;
    std::multimap<size_t, CUmemGenericAllocationHandle> phys_pool_;
    uint64_t phys_pool_bytes_ = 0;
    std::vector<Slab> slabs_;

    // independent queues: a pack never waits behind an unpack that is itself waiting for its H2D, and a victim
    // scan never waits behind either
    CUstream s_scan_ = nullptr, s_pack_ = nullptr, s_unpack_ = nullptr, s_out_ = nullptr, s_in_ = nullptr;
    std::vector<Slot> ring_out_, ring_in_;
    int cur_out_ = 0, cur_in_ = 0;
    std::vector<CUevent> use_ring_;                 // last-use events, indexed by seq % size
    uint64_t use_seq_ = 0;
    std::vector<CUevent> ready_free_;
    std::unique_ptr<VictimScanner> scanner_;
    std::vector<PendingHost> pending_host_;
    std::vector<uint32_t> out_pending_;             // victims packed but not yet unmapped (page_out_finish)
    // reaper: waits for a batch's last pack, unmaps the victims (VMM calls cost 0.1-1 ms each under load on B200) and
    // returns their physical handles to the pool while the admitting thread is already mapping the incoming rows
    struct ReapJob { std::vector<uint32_t> rows; std::vector<CUdeviceptr> bases; std::vector<size_t> mapped; CUevent packed; };
    std::thread reaper_;
    std::mutex rq_mu_;
    std::condition_variable rq_cv_;
    std::condition_variable_any reap_cv_;          // waited on with mu_ held
    std::deque<ReapJob> rq_;
    bool reaper_stop_ = false, reaper_busy_ = false;
    CUcontext ctx_ = nullptr;
    bool ctx_warned_ = false;
    uint64_t evicting_mapped_ = 0;
    void reaper_main();
    void wait_not_evicting(int row);
    std::vector<InJob> in_jobs_;                    // page-in plan of the admission in progress
    struct TraceRec { uint64_t t_begin = 0, t_packs = 0, t_staged = 0, t_unmapped = 0, t_mapped = 0, t_end = 0; std::vector<CUevent> d2h, h2d; };
    std::vector<TraceRec> trace_;
    uint32_t trace_want_ = 0, trace_skip_ = 0;
    bool trace_dumped_ = false;
    CUevent trace_base_ = nullptr;
    uint64_t trace_base_ns_ = 0;
    TraceRec *tr_ = nullptr;                        // record of the admission in progress (or null)
    void trace_mark(std::vector<CUevent> *v, CUstream s);
    // victims selected by the last scan beyond what was needed then, in LRU order. They stay the exact LRU prefix for
    // as long as they are untouched (anything touched or created since carries a larger tick), so consuming them
    // is equivalent to re-scanning; an entry whose row changed is simply skipped.
    struct Cand { uint32_t row; uint64_t touch, base; };
    std::deque<Cand> victim_cache_;
    uint32_t scan_lookahead_ = 8;
    SwapStats st_;
    vgpu_swap_record_t *shared_ = nullptr;
    // Physical pressure: the quota (quota_cap_) promises more than the device can give right now — other containers of an
    // overcommitted GPU (DeviceMemoryScaling > 1, server.go:356) hold the rest. The working cap (cfg_.resident_cap) is
    // lowered to what was obtainable and probed back up every few admissions.
    uint64_t quota_cap_ = 0;
    bool pressure_ = false;
    uint32_t pressure_probe_ = 0;
    uint64_t pressure_events_ = 0;
    void publish_locked();
    struct Prof { CUevent a, b; bool unpack; uint64_t bytes; };
    CUdeviceptr d_span_ = 0;                        // profiling: {min start, max end} per launch, pre-set to {~0, 0}
    uint32_t span_cap_ = 0, span_next_ = 0, span_read_ = 0;
    std::vector<uint8_t> span_unpack_;
    CUdeviceptr next_span(bool unpack);
    void harvest_spans();
    std::vector<Prof> prof_;
};

}  // namespace vgpu
