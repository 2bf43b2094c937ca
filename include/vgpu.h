/*
 * vgpu.h — C ABI of the B200-native vGPU enforcement library (libvgpu.so / libvgpu_core.so).
 *
 * Two consumers:
 *  1. Host-side tools written in the reference's language (Go, via cgo): the node monitor and the device plugin.
 *     The reference has no FFI here — cmd/vGPUmonitor/cudevshr.go:18-58 re-declares the shared-region struct by
 *     hand and mmaps the file itself (mmapcachefile, cudevshr.go:112-127), and feedback.go:197-255 pokes
 *     recentKernel/utilizationSwitch straight into the mapping. vgpu_region_* replaces that hand copy
 *     (INTEGRATION.md shows the cgo stub).
 *  2. Tests and the benchmark, which drive the sm_100a kernels and the swap engine through plain pointers and
 *     sizes (no torch types, no CUDA headers needed): device addresses are uint64_t, streams are void* (CUstream).
 *
 * The application-facing boundary is NOT this header: applications keep calling the CUDA driver API / NVML, which
 * libvgpu.so interposes exactly like the reference hook (k8s-device-plugin_b200/csrc/hook.cc).
 *
 * All functions returning int return a CUresult-compatible code (0 = success) unless stated otherwise.
 */
#ifndef VGPU_H
#define VGPU_H

#include <stddef.h>
#include <stdint.h>

#include "vgpu_region.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *vgpu_version(void);

/* ---- limits (reference: get_limit_from_env libvgpu.so@0x40d00; producer server.go:343-345 "<MiB>m") */
uint64_t vgpu_parse_limit(const char *value);

/* ---- shared region (reference: try_create_shrreg@0x44138, lock_shrreg@0x437d3, add/rm_gpu_device_memory_usage
 * @0x42a0e/@0x42de1, get_gpu_memory_usage@0x420bd; Go consumer cmd/vGPUmonitor/{cudevshr,feedback,metrics}.go) */
typedef struct vgpu_region_handle vgpu_region_handle_t;

typedef struct vgpu_proc_usage {
    int32_t pid, hostpid, status, _pad;
    vgpu_device_memory_t used[VGPU_MAX_DEVICES];
} vgpu_proc_usage_t;

typedef struct vgpu_region_snapshot {
    int32_t initialized;                       /* magic matched */
    int32_t proc_num;
    int32_t utilization_switch, recent_kernel, priority, _pad;
    uint64_t device_num;
    uint64_t limit[VGPU_MAX_DEVICES];
    uint64_t sm_limit[VGPU_MAX_DEVICES];
    uint64_t usage_total[VGPU_MAX_DEVICES];    /* sum over processes of used[d].total (getDeviceUsedMemory, cudevshr.go:98-110) */
    char uuids[VGPU_MAX_DEVICES][VGPU_UUID_LEN];
} vgpu_region_snapshot_t;

/* create != 0: create+initialise when absent (mem_limits/sm_limits: 16 entries each or NULL). */
int vgpu_region_open(const char *path, int create, const uint64_t *mem_limits, const uint64_t *sm_limits, int priority,
                     vgpu_region_handle_t **out);
void vgpu_region_close(vgpu_region_handle_t *h);
int vgpu_region_snapshot(vgpu_region_handle_t *h, vgpu_region_snapshot_t *out);          /* lock-free, like the monitor */
int vgpu_region_proc(vgpu_region_handle_t *h, int index, vgpu_proc_usage_t *out);        /* index < proc_num */
int vgpu_region_claim(vgpu_region_handle_t *h, int32_t pid);                              /* returns slot or -1 */
void vgpu_region_release(vgpu_region_handle_t *h, int32_t pid);
int vgpu_region_try_add(vgpu_region_handle_t *h, int32_t pid, int dev, uint64_t bytes, int type, int enforce); /* 1 ok, 0 quota breach */
void vgpu_region_sub(vgpu_region_handle_t *h, int32_t pid, int dev, uint64_t bytes, int type);
uint64_t vgpu_region_usage(vgpu_region_handle_t *h, int dev);
/* monitor write-back (feedback.go:153-156,207-251); pass INT32_MIN to leave a field unchanged */
int vgpu_region_set_feedback(vgpu_region_handle_t *h, int32_t recent_kernel, int32_t utilization_switch);
int vgpu_region_set_hostpid(vgpu_region_handle_t *h, int32_t pid, int32_t hostpid);
/* publish a device identity in lane `dev` (the hook does this itself from NVML: put_device_info, multiprocess_memory_limit.c:L150) */
int vgpu_region_set_uuid(vgpu_region_handle_t *h, int dev, const char *uuid);
void *vgpu_region_raw(vgpu_region_handle_t *h);                                          /* vgpu_shared_region_t* */
/* swap counters of the container on device `dev`, summed over its processes (extension block, vgpu_region.h). Returns 0,
 * or -1 when the file carries no extension block (created by the reference hook). out->pid = number of records summed. */
int vgpu_region_swap_counters(vgpu_region_handle_t *h, int dev, vgpu_swap_record_t *out);

/* ---- node monitor feedback (reference: Observe cmd/vGPUmonitor/feedback.go:197-255, CheckBlocking :165-179, CheckPriority
 * :181-195). One pass over the regions of every container on the node: decrements recentKernel, counts active tasks per GPU
 * UUID and priority, then blocks lower-priority containers (recentKernel = -1) while a higher-priority one is active and
 * switches the core limiter on only where a GPU is actually contended (utilizationSwitch). Returns the number of regions
 * whose words were changed. */
int vgpu_monitor_observe(vgpu_region_handle_t **regions, int n);

/* ---- sm_100a kernels, directly (a CUDA context must be current on the calling thread) */
typedef struct vgpu_seg { uint64_t src, dst, bytes; } vgpu_seg_t;
/* pack / unpack: copy every segment src->dst on `stream` (TMA path for 16-byte-aligned segments). */
int vgpu_pack(const vgpu_seg_t *segs, size_t nseg, void *stream);
/* launch geometry of the TMA pack kernel (0 = keep): tile bytes (multiple of 16), ring stages (2..8), CTAs per SM */
int vgpu_pack_config(uint32_t tile_bytes, uint32_t stages, uint32_t ctas_per_sm);

typedef struct vgpu_entry {   /* one row of the allocation table; identical to VgpuEntry in csrc/kernels.h */
    uint64_t base, size, last_touch;
    uint32_t state, host_slot;
} vgpu_entry_t;
#define VGPU_ENTRY_FREE 0u
#define VGPU_ENTRY_RESIDENT 1u
#define VGPU_ENTRY_PAGED_OUT 2u
#define VGPU_ENTRY_PINNED 4u
/* Exact-LRU victim scan over a DEVICE-resident table of n rows: the shortest prefix, in (last_touch, index)
 * order, of RESIDENT rows whose sizes sum to >= need. out_idx (host memory, out_cap entries) receives the row
 * indices in ascending index order. Synchronises the stream. */
int vgpu_victim_scan(uint64_t d_table, uint32_t n, uint64_t need, uint64_t max_touch, void *stream, uint32_t *out_idx,
                     uint32_t out_cap, uint32_t *out_count, uint64_t *freed, int *insufficient);

/* synthetic workload kernels (SURVEY.md §8d cfg 3): word j of buffer i = splitmix64((i << 32) + j) */
int vgpu_wl_fill(uint64_t dptr, uint64_t nwords, uint64_t buf_index, void *stream);
int vgpu_wl_touch(uint64_t dptr, uint64_t nwords, void *stream);                         /* x += 1 */
/* x += 1 on nptr buffers of nwords each, reached through a table of device pointers that itself lives in device memory */
int vgpu_wl_touch_indirect(uint64_t d_table, uint32_t nptr, uint64_t nwords, void *stream);
int vgpu_wl_verify(uint64_t dptr, uint64_t nwords, uint64_t buf_index, uint64_t added, uint64_t d_mismatch_counter, void *stream);

/* ---- swap engine (new functionality behind CUDA_OVERSUBSCRIBE; reference switch cuMemoryAllocate@0x315da) */
typedef struct vgpu_swap vgpu_swap_t;
typedef struct vgpu_swap_config {
    uint64_t resident_cap;     /* bytes of physical backing allowed (the gpumem quota); 0 = device free memory */
    uint64_t virtual_cap;      /* live bytes allowed; 0 = bounded by the host pool */
    uint64_t host_pool_cap;    /* pinned bytes allowed; 0 = unbounded */
    uint64_t chunk_bytes;      /* staging slot size; 0 = 32 MiB */
    uint32_t ring_slots;       /* staging slots per direction; 0 = 4 */
    uint32_t profile;          /* 1 = time pack/unpack launches with events */
    uint64_t prefetch_bytes;   /* how far the pager runs ahead of the application; 0 = default (VGPU_SWAP_PREFETCH_MB), ~0 = off */
    uint64_t copy_bytes;       /* piece size of the direct DMA copies; 0 = default (VGPU_SWAP_COPY_MB) */
} vgpu_swap_config_t;
typedef struct vgpu_swap_stats {
    uint64_t page_out_bytes, page_in_bytes, evictions, faults, admissions;
    uint64_t pack_launches, unpack_launches, scan_launches, scans;
    uint64_t resident_bytes, live_bytes, host_bytes, entries, phys_creates, phys_reuses;
    uint64_t pack_bytes, unpack_bytes;         /* bytes moved by the staged path's TMA kernels */
    double pack_ms, unpack_ms;
    uint64_t scan_cache_hits;   /* evictions served from the previous scan's surplus (no new scan) */
    /* APPLICATION-thread time inside admissions (ns): total, blocked until the pager had issued the page-in, and inside VMM
     * calls (zero by construction: the pager thread owns them) */
    uint64_t host_admit_ns, host_wait_ns, host_vmm_ns;
    /* pager-thread time (ns): VMM calls, victim scans, waiting for the last pack of a staged batch, staging back-pressure,
     * all steps that made progress; and the number of cuMemUnmap + cuMemSetAccess calls after batching */
    uint64_t pager_vmm_ns, pager_scan_ns, pager_packsync_ns, pager_ring_ns, pager_busy_ns, vmm_calls;
    double pack_span_ms, unpack_span_ms;   /* exact in-kernel %globaltimer execution spans (profiling), valid after vgpu_swap_drain */
    uint64_t direct_out_bytes, direct_in_bytes;   /* bytes moved by plain DMA between a buffer's own range and its pinned block */
    uint64_t prefetch_issued, prefetch_hits, prefetch_wasted;   /* rows paged in ahead of need / touched afterwards / evicted untouched */
    uint64_t demand_waits;      /* admissions that had to block for the pager */
    uint64_t clean_evictions;   /* evictions without a copy (pinned block still valid) */
    uint64_t host_slabs, host_slabs_local;   /* pinned slabs allocated / of those on the GPU's NUMA node */
    /* diagnostics: where the pager's time goes (ns) — cuMemUnmap, cuMemSetAccess, copy/event enqueue calls, event polling,
     * re-taking the engine lock, and busy time per step (freed rows, reap, demand, prefetch, evict-ahead) */
    uint64_t pager_unmap_ns, pager_setaccess_ns, pager_issue_ns, pager_poll_ns, pager_lock_ns, pager_step_ns[5];
    uint64_t vmm_slow_calls, vmm_slow_ns, vmm_max_ns;   /* VMM calls that took > 2 ms (stalls inside the driver), their total, the worst */
    uint64_t inplace_uses;      /* host-backed mode: operands of oversized launches that were used where they were (host-mapped) */
} vgpu_swap_stats_t;
int vgpu_swap_create(int dev, const vgpu_swap_config_t *cfg, vgpu_swap_t **out);
void vgpu_swap_destroy(vgpu_swap_t *s);
int vgpu_swap_alloc(vgpu_swap_t *s, uint64_t bytes, uint64_t *dptr);
int vgpu_swap_free(vgpu_swap_t *s, uint64_t dptr);
/* make the buffers containing ptrs[] resident and order `stream` behind the page-ins; pair with vgpu_swap_release */
int vgpu_swap_acquire(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream);
int vgpu_swap_release(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream);
/* like vgpu_swap_release for work that only READ the buffers (they stay clean: evicting them later needs no copy) */
int vgpu_swap_release_ro(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream);
/* cuMemAdvise(SET/UNSET_READ_MOSTLY) for a swappable buffer: kernel launches no longer mark it dirty */
int vgpu_swap_advise_read_mostly(vgpu_swap_t *s, uint64_t ptr, int on);
/* cuMemPrefetchAsync for a swappable buffer: to_device != 0 queues a page-in with the pager; 0 ("to the host") makes the
 * buffer the first victim when room is needed */
int vgpu_swap_prefetch(vgpu_swap_t *s, uint64_t ptr, int to_device);
/* keep a buffer resident for good (operands the argument scan cannot see: device-side pointer tables) */
int vgpu_swap_pin(vgpu_swap_t *s, uint64_t ptr, int on);
int vgpu_swap_stats(vgpu_swap_t *s, vgpu_swap_stats_t *out);
int vgpu_swap_drain(vgpu_swap_t *s);
int vgpu_swap_set_profile(vgpu_swap_t *s, int on);   /* event brackets + in-kernel spans around pack/unpack launches from now on */
int vgpu_swap_table(vgpu_swap_t *s, vgpu_entry_t *out, uint32_t cap, uint32_t *n);

/* ---- gpucores limiter (reference: rate_limiter@0x4591a / utilization_watcher@0x46710) */
typedef struct vgpu_limiter vgpu_limiter_t;
typedef struct vgpu_limiter_stats {
    uint64_t launches, stamps, groups, busy_ns, throttle_ns, wall_ns;
    int32_t limit_percent, _pad;
} vgpu_limiter_stats_t;
int vgpu_limiter_create(int percent, vgpu_limiter_t **out);
void vgpu_limiter_destroy(vgpu_limiter_t *l);
void vgpu_limiter_before_launch(vgpu_limiter_t *l, void *stream);
void vgpu_limiter_after_launch(vgpu_limiter_t *l, void *stream);
int vgpu_limiter_stats(vgpu_limiter_t *l, vgpu_limiter_stats_t *out);

/* ---- introspection of the in-process hook runtime (valid in a process that has libvgpu.so preloaded) */
int vgpu_runtime_swap_stats(int dev, vgpu_swap_stats_t *out);       /* nonzero when no engine exists on dev */
int vgpu_runtime_limiter_stats(vgpu_limiter_stats_t *out);
int vgpu_runtime_set_swap_profile(int dev, int on);
int vgpu_runtime_swap_pin(uint64_t dptr, int on);                     /* vgpu_swap_pin on the hook's engine of the current device */
uint64_t vgpu_runtime_context_size(void);
int vgpu_runtime_check_memory_type(uint64_t dptr);                    /* check_memory_type@0x407f2: 2 tracked, 1 not */

#ifdef __cplusplus
}
#endif
#endif /* VGPU_H */
