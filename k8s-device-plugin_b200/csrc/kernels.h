// kernels.h — parameter/record layouts shared by kernels.cu and the host-side launch code.
#pragma once
#include <stdint.h>

#define VGPU_PACK_MAX_STAGES 8             /* upper bound of the shared-memory ring depth of vgpu_pack_tma */
/* Default geometry from scripts/pack_sweep.py on B200 (profiles/r01_pack_sweep.json): 8 KiB tiles x 3 stages x 3 CTAs/SM
 * reaches 6543 GB/s = 0.996 of the measured HBM copy peak with only 24 KiB of shared memory per CTA (72 KiB per SM),
 * so a pack, an unpack and an application kernel co-reside on every SM instead of queueing for shared memory
 * (the 6 x 32 KiB single-CTA ring first tried took 192 KiB and measured 0.914; 32 KiB x 2 x 2 is the top at 1.008). */
#define VGPU_PACK_STAGES 3
#define VGPU_PACK_TILE_BYTES (8u * 1024u)
#define VGPU_PACK_CTAS_PER_SM 3            /* persistent CTAs per SM (each drives its own ring from one elected thread) */
#define VGPU_PACK_MAX_SEG 96               /* segments per launch; keeps the kernel parameter block < 4 KiB */
#define VGPU_SCAN_DIGIT_BITS 11
#define VGPU_SCAN_BINS (1 << VGPU_SCAN_DIGIT_BITS)

/* allocation-table row states (device-resident table; oracle/vgpu_oracle.h mirrors these for the CPU checker) */
#define VGPU_ST_FREE 0u
#define VGPU_ST_RESIDENT 1u
#define VGPU_ST_PAGED_OUT 2u
#define VGPU_ST_PINNED 4u                  /* flag OR-ed onto RESIDENT: in use by an admission in progress, not evictable */

typedef struct VgpuEntry {                 /* 32 bytes, 32-byte aligned rows */
    uint64_t base;                         /* device VA (stable across page-out/page-in) */
    uint64_t size;                         /* requested bytes */
    uint64_t last_touch;                   /* logical launch tick of the last kernel/memcpy that referenced the buffer */
    uint32_t state;
    uint32_t host_slot;                    /* page index into the pinned pool while paged out */
} VgpuEntry;

typedef struct VgpuPackSeg {
    uint64_t src;                          /* device address */
    uint64_t dst;                          /* device address */
    uint64_t bytes;
    uint64_t tile_begin;                   /* index of this segment's first tile within the launch */
} VgpuPackSeg;

typedef struct VgpuPackParams {
    uint32_t nseg;
    uint32_t tile_bytes;
    uint64_t total_tiles;
    uint32_t stages;                       /* ring depth used by this launch (<= VGPU_PACK_MAX_STAGES) */
    uint32_t _pad;
    uint64_t span;                         /* optional device address of {min start, max end} %globaltimer words (0 = off) */
    VgpuPackSeg seg[VGPU_PACK_MAX_SEG];
} VgpuPackParams;

#define VGPU_SCAN_SMALL_ROWS_PER_THREAD 8u  /* single-CTA scan: 1024 threads x 8 rows held in registers */
#define VGPU_SCAN_SMALL_MAX_ROWS 8192u
#define VGPU_SCAN_MAX_CTAS 1024u       /* grid of the scan launches is capped at min(this, 4 x SMs) */
typedef struct VgpuScanState {
    uint64_t prefix;                       /* selected high digits of K* so far */
    uint64_t need_left;
    uint64_t need;
    uint64_t cand_bytes;                   /* total candidate bytes (valid when insufficient) */
    uint64_t out_freed;                    /* size sum of the emitted victims */
    uint32_t done_ctas;
    uint32_t insufficient;
    uint32_t out_count;
    uint32_t chain_flag;
    uint32_t chain_offset;
    uint32_t _pad;
    uint64_t hist[VGPU_SCAN_BINS];
    /* ordered emit: per-CTA victim counts, turned into exclusive offsets by the last CTA of the count launch */
    uint32_t cta_off[VGPU_SCAN_MAX_CTAS];
    uint64_t cta_bytes[VGPU_SCAN_MAX_CTAS];
    /* vgpu_victim_persist: second histogram (the digit passes alternate between the two, so that the CTA that picks a digit
     * can clear the one it just read without a second grid barrier) and the grid barrier's arrival counter / generation.
     * Invariant between launches: both histograms and bar_count are zero. */
    uint64_t hist2[VGPU_SCAN_BINS];
    uint32_t bar_count, bar_gen;
} VgpuScanState;
#define VGPU_SCAN_PERSIST_ROWS_PER_CTA (1024u * VGPU_SCAN_SMALL_ROWS_PER_THREAD)   /* rows a CTA of vgpu_victim_persist keeps in registers */
