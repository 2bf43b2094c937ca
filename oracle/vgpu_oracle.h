/*
 * vgpu_oracle.h — CPU restatement of the reference hook's hot-path arithmetic. TEST INFRASTRUCTURE:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it (never the product).
 *
 * PARITY PIN: the reference's own tests hold no vectors for this path, and its
 * source (submodule HAMi-core) is absent; the restatement is therefore pinned against the
 * reference's shipped BINARY lib/nvidia/libvgpu.so executed here on oracle/fake_driver (the
 * same driver-API trace through both; tests/test_oracle_pin.py, fixtures in tests/golden/).
 * Each function cites the binary address it restates (libvgpu.so@0x…, orig src:line).
 */
#ifndef VGPU_ORACLE_H
#define VGPU_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VO_IPCSIZE (2u << 20)   /* .data@0x610a8: allocations above this take the allocmode switch */

typedef struct vo_state vo_state_t;

/* get_limit_from_env@0x40d00 (multiprocess_memory_limit.c:L101-111) on the VALUE string */
uint64_t vo_parse_limit(const char *value);

/* one container, one process, device 0. ctx_bytes = NVML usedGpuMemory seen by set_task_pid@0x16a7f
 * (added as type 0 by cuDevicePrimaryCtxRetain, context.c:L72-86); dev_total = real device memory. */
vo_state_t *vo_create(uint64_t limit, uint64_t ctx_bytes, uint64_t dev_total);
void vo_destroy(vo_state_t *);

/* return codes are the reference's: 0 ok, -1 quota breach in cuMemAlloc_v2 (add_chunk@0x4005d),
 * 2 quota breach in managed/pitch (cuMemAllocManaged@0x31eab, cuMemAllocPitch_v2@0x321ef) */
int vo_alloc(vo_state_t *, uint64_t id, uint64_t size);                 /* cuMemAlloc_v2@0x3180a */
int vo_alloc_managed(vo_state_t *, uint64_t id, uint64_t size);         /* cuMemAllocManaged@0x31cb1 */
int vo_alloc_pitch(vo_state_t *, uint64_t id, uint64_t width, uint64_t height, unsigned elem); /* @0x31fa5 */
int vo_free(vo_state_t *, uint64_t id);                                 /* cuMemFree_v2@0x322f1 -> remove_chunk@0x40871 */
int vo_free_untracked(vo_state_t *, uint64_t addr);                     /* addr 0 -> 0, else -1 */
int vo_mem_get_info(vo_state_t *, uint64_t *free_b, uint64_t *total_b); /* cuMemGetInfo_v2@0x367dc */
uint64_t vo_total_mem(vo_state_t *);                                    /* cuDeviceTotalMem_v2@0x2d3f8 */
/* counters of device 0: context, module, buffer, offset, total (Appendix A used[] lane) */
void vo_counters(const vo_state_t *, uint64_t out[5]);

/* multiprocess_utilization_watcher.c: setspec@0x45d5b, delta@0x45c7b — int32 arithmetic, wraps like the binary */
int32_t vo_total_cuda_cores(int32_t sm_num, int32_t max_threads_per_sm);
int32_t vo_delta(int32_t sm_num, int32_t max_threads_per_sm, int32_t total_cores,
                 int32_t up_limit, int32_t user_current, int32_t share);

/* ---- spec oracles for the NEW swap kernels (the reference has no counterpart: its swap is UVM,
 * cuMemoryAllocate@0x315da). They restate DESIGN.md's definitions so the CUDA kernels can be checked bit-exactly. */
typedef struct {            /* one row of the device-resident allocation table (32 B) */
    uint64_t base;          /* device VA */
    uint64_t size;          /* requested bytes */
    uint64_t last_touch;    /* logical launch tick of the last kernel that referenced it */
    uint32_t state;         /* VO_ST_* */
    uint32_t host_slot;     /* index into the pinned pool when paged out */
} vo_entry_t;
#define VO_ST_FREE 0u
#define VO_ST_RESIDENT 1u
#define VO_ST_PAGED_OUT 2u
#define VO_ST_PINNED 4u     /* flag: resident and not evictable (in use by the launch being admitted) */

/* Exact LRU victim choice: among rows with state == VO_ST_RESIDENT (not pinned), ordered by
 * (last_touch, index) ascending, the shortest prefix whose size sum >= need. Writes the row indices in
 * ascending INDEX order; returns the count, or -1 when even evicting every candidate is not enough
 * (then out holds all candidates). *freed = size sum of the chosen rows. */
int64_t vo_select_victims(const vo_entry_t *tbl, uint64_t n, uint64_t need, uint32_t *out, uint64_t *freed);

typedef struct { uint64_t src_off, dst_off, bytes; } vo_seg_t;
/* pack: dst[dst_off .. +bytes) = src[src_off .. +bytes) for each segment (gather into staging);
 * unpack is the same call with the roles of the offsets swapped by the caller. */
void vo_pack(uint8_t *dst, const uint8_t *src, const vo_seg_t *segs, uint64_t nseg);

/* splitmix64 — the fill pattern of the swap workload (SURVEY.md §8d cfg 3): word j of buffer i = splitmix64((i<<32)+j) */
uint64_t vo_splitmix64(uint64_t x);

#ifdef __cplusplus
}
#endif
#endif
