// kernels.cu — every device kernel of the library, compiled to ONE sm_100a cubin that is embedded in the shared
// objects and loaded with cuModuleLoadData into the application's own context (the hook is a driver-API
// library; it deliberately has no cudart dependency).
//
// The reference has no device code at all (zero .cu files; its "swap" is cuMemAllocManaged + the UVM
// driver, cuMemoryAllocate libvgpu.so@0x315da). These kernels are the new swap engine's data path:
//   vgpu_pack_tma      pack/compact (and, with src/dst exchanged, unpack/scatter) copy: global -> shared -> global
//                      with 1-D TMA bulk copies (cp.async.bulk), mbarrier-tracked, one elected thread per CTA,
//                      persistent over 148 SMs. Pure data movement: HBM-bound, 2 bytes of HBM traffic per byte moved.
//   vgpu_pack_generic  same contract for segments that are not 16-byte aligned (LSU path, any alignment).
//   vgpu_victim_*      exact LRU victim choice over the device-resident allocation table: weighted radix select on
//                      the key (last_touch, index) + ordered compaction. 32 bytes of HBM read per table row.
//   vgpu_stamp         %globaltimer stamps for the device-timestamped gpucores token bucket.
//   vgpu_wl_*          the synthetic alloc+touch workload (fill / read-modify-write / verify) used by bench + tests.
#include <cstdint>

#include "kernels.h"

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ------------------------------------------------------------------------------------------------ pack / unpack
// Tile t of the launch belongs to the segment s with seg[s].tile_begin <= t < seg[s+1].tile_begin; inside the
// segment it covers bytes [ (t - tile_begin) * tile_bytes, +tile_bytes ) clipped to seg.bytes.
__device__ __forceinline__ int find_seg(const VgpuPackParams &p, uint64_t tile) {
    int lo = 0, hi = static_cast<int>(p.nseg) - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (p.seg[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// One elected thread per CTA drives a VGPU_PACK_STAGES-deep ring of tile buffers in shared memory:
//   load(i)  : cp.async.bulk global->smem, arrival on full[stage]      (up to STAGES-1 loads in flight)
//   store(i) : cp.async.bulk smem->global, bulk_group                  (stage recycled once the store has READ it)
// No thread ever touches the payload: both directions run on the TMA unit, the SM only issues descriptors.
extern "C" __global__ void __launch_bounds__(32, 1) vgpu_pack_tma(const __grid_constant__ VgpuPackParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int ST = static_cast<int>(p.stages);
    const uint32_t tile_bytes = p.tile_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);              // up to VGPU_PACK_MAX_STAGES barriers
    unsigned char *buf = smem + 128;                                   // ST tiles
    if (threadIdx.x != 0) return;
    const uint64_t t_start = p.span ? globaltimer() : 0;

    for (int s = 0; s < ST; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();

    const uint64_t total = p.total_tiles;
    const uint64_t stride = gridDim.x;
    const uint64_t first = blockIdx.x;
    if (first >= total) return;
    const uint64_t my_tiles = (total - first + stride - 1) / stride;
    unsigned long long *span = reinterpret_cast<unsigned long long *>(p.span);

    auto tile_desc = [&](uint64_t k, const unsigned char *&src, unsigned char *&dst, uint32_t &len) {
        uint64_t t = first + k * stride;
        int s = find_seg(p, t);
        uint64_t off = (t - p.seg[s].tile_begin) * static_cast<uint64_t>(tile_bytes);
        uint64_t rem = p.seg[s].bytes - off;
        len = rem < tile_bytes ? static_cast<uint32_t>(rem) : tile_bytes;
        src = reinterpret_cast<const unsigned char *>(p.seg[s].src) + off;
        dst = reinterpret_cast<unsigned char *>(p.seg[s].dst) + off;
    };
    auto issue_load = [&](uint64_t k) {
        const unsigned char *src; unsigned char *dst; uint32_t len;
        tile_desc(k, src, dst, len);
        int st = static_cast<int>(k % ST);
        mbar_expect_tx(&full[st], len);
        bulk_g2s(buf + static_cast<size_t>(st) * tile_bytes, src, len, &full[st]);
    };

    const uint64_t pre = my_tiles < static_cast<uint64_t>(ST - 1) ? my_tiles : static_cast<uint64_t>(ST - 1);
    for (uint64_t k = 0; k < pre; k++) issue_load(k);

    for (uint64_t k = 0; k < my_tiles; k++) {
        const unsigned char *src; unsigned char *dst; uint32_t len;
        tile_desc(k, src, dst, len);
        int st = static_cast<int>(k % ST);
        mbar_wait(&full[st], static_cast<uint32_t>((k / ST) & 1));
        bulk_s2g(dst, buf + static_cast<size_t>(st) * tile_bytes, len);
        bulk_commit();
        // stage (k-1)%ST is reusable once store(k-1) has finished reading shared memory; store(k) may stay pending
        uint64_t nk = k + ST - 1;
        if (nk < my_tiles) {
            bulk_wait_read<1>();
            issue_load(nk);
        }
    }
    bulk_wait_all();
    if (span) {   // exact execution span of this launch under whatever else shares the GPU: min start / max end over CTAs
        atomicMin(&span[0], static_cast<unsigned long long>(t_start));
        atomicMax(&span[1], static_cast<unsigned long long>(globaltimer()));
    }
}

// Any-alignment fallback: all threads of the CTA copy one tile at a time through registers. Used for segments whose
// src/dst/bytes are not multiples of 16 (the swap engine never produces such segments; the C ABI accepts them).
extern "C" __global__ void __launch_bounds__(256) vgpu_pack_generic(const __grid_constant__ VgpuPackParams p) {
    const uint32_t tile_bytes = p.tile_bytes;
    for (uint64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int s = find_seg(p, t);
        uint64_t off = (t - p.seg[s].tile_begin) * static_cast<uint64_t>(tile_bytes);
        uint64_t rem = p.seg[s].bytes - off;
        uint32_t len = rem < tile_bytes ? static_cast<uint32_t>(rem) : tile_bytes;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(p.seg[s].src) + off;
        unsigned char *dst = reinterpret_cast<unsigned char *>(p.seg[s].dst) + off;
        if (((reinterpret_cast<uintptr_t>(src) ^ reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
            // same phase: byte head up to a 16-byte boundary, uint4 body, byte tail
            uint32_t head = static_cast<uint32_t>((16u - (reinterpret_cast<uintptr_t>(src) & 15u)) & 15u);
            if (head > len) head = len;
            for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
            uint32_t body = (len - head) >> 4;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
            for (uint32_t i = threadIdx.x; i < body; i += blockDim.x) d4[i] = s4[i];
            uint32_t done = head + (body << 4);
            for (uint32_t i = done + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
        } else {
            for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ victim scan
// Exact LRU: candidates are rows with state == VGPU_ST_RESIDENT; order = (last_touch, index) ascending; the
// victims are the shortest prefix of that order whose size sum reaches `need`. Implemented as a byte-weighted
// radix select on key = (last_touch << idx_bits) | index, most significant digit first (11-bit digits), one launch
// per digit; the last CTA of each launch (atomic ticket) picks the digit, so no host round trip between passes.
__device__ __forceinline__ uint64_t row_key(const VgpuEntry &e, uint32_t idx, uint32_t idx_bits) {
    return (e.last_touch << idx_bits) | static_cast<uint64_t>(idx);
}
__device__ __forceinline__ VgpuEntry load_row(const VgpuEntry *tbl, uint32_t i) {
    // 32-byte row = two 16-byte loads; a warp reads 1 KB contiguous
    const uint4 *p = reinterpret_cast<const uint4 *>(tbl + i);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    VgpuEntry e;
    e.base = (static_cast<uint64_t>(a.y) << 32) | a.x;
    e.size = (static_cast<uint64_t>(a.w) << 32) | a.z;
    e.last_touch = (static_cast<uint64_t>(b.y) << 32) | b.x;
    e.state = b.z;
    e.host_slot = b.w;
    return e;
}

extern "C" __global__ void vgpu_victim_init(VgpuScanState *st, uint64_t need) {
    for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x) st->hist[i] = 0;
    if (threadIdx.x == 0) {
        st->prefix = 0; st->need_left = need; st->need = need; st->done_ctas = 0; st->insufficient = 0;
        st->out_count = 0; st->out_freed = 0; st->chain_flag = 0; st->chain_offset = 0; st->cand_bytes = 0;
    }
}

// digit = bits [shift, shift+width) of the key; rows take part when key >> (shift+width) == prefix
extern "C" __global__ void __launch_bounds__(256) vgpu_victim_hist(const VgpuEntry *__restrict__ tbl, uint32_t n,
                                                                   VgpuScanState *st, uint32_t idx_bits,
                                                                   uint32_t shift, uint32_t width) {
    __shared__ unsigned long long hist[VGPU_SCAN_BINS];
    __shared__ bool last;
    if (st->insufficient) return;   // uniform: written by a previous launch
    for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint64_t prefix = st->prefix;
    const uint32_t hi = shift + width;
    const uint32_t mask = (1u << width) - 1u;
    // four independent row loads in flight per thread before any of them is consumed (memory-level parallelism: one
    // 32-byte row per thread per trip leaves the SM far below the bytes in flight HBM3e needs)
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        VgpuEntry e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t i = i0 + u * stride;
            if (i < n) e[u] = load_row(tbl, i); else e[u].state = VGPU_ST_FREE;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t i = i0 + u * stride;
            if (e[u].state != VGPU_ST_RESIDENT) continue;
            uint64_t key = row_key(e[u], i, idx_bits);
            uint64_t top = hi >= 64 ? 0 : (key >> hi);
            if (top != prefix) continue;
            atomicAdd(&hist[static_cast<uint32_t>(key >> shift) & mask], static_cast<unsigned long long>(e[u].size));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x)
        if (hist[i]) atomicAdd(reinterpret_cast<unsigned long long *>(&st->hist[i]), hist[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&st->done_ctas, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last CTA: pick the digit. warp 0: lane l owns bins [l*64, l*64+64)
    if (threadIdx.x < 32) {
        constexpr int PER = VGPU_SCAN_BINS / 32;
        volatile uint64_t *gh = st->hist;
        uint64_t mine = 0;
        for (int j = 0; j < PER; j++) mine += gh[threadIdx.x * PER + j];
        uint64_t incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (static_cast<int>(threadIdx.x) >= d) incl += o;
        }
        uint64_t excl = incl - mine;
        uint64_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint64_t need = st->need_left;
        if (total < need) {
            if (threadIdx.x == 0) { st->insufficient = 1; st->cand_bytes = total; }
        } else if (excl < need && need <= incl) {   // exactly one lane (need >= 1)
            uint64_t cum = excl;
            int b = threadIdx.x * PER;
            for (int j = 0; j < PER; j++, b++) {
                uint64_t h = gh[b];
                if (cum + h >= need) break;
                cum += h;
            }
            st->prefix = (st->prefix << width) | static_cast<uint64_t>(b);
            st->need_left = need - cum;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x) st->hist[i] = 0;
    if (threadIdx.x == 0) st->done_ctas = 0;
}

// Ordered compaction of rows with key <= K* (all candidates when insufficient), ascending row index, in two launches
// and without any inter-CTA spinning: CTA b owns rows [b*chunk, (b+1)*chunk).
//   count: every CTA publishes how many of its rows qualify (and their bytes); the last CTA to finish (atomic ticket)
//          turns the per-CTA counts into exclusive offsets and the totals.
//   emit : every CTA recounts per thread, scans inside the block and writes behind its offset.
__device__ __forceinline__ bool victim_row(const VgpuEntry &e, uint32_t i, uint32_t idx_bits, bool all, bool none, uint64_t kstar) {
    if (e.state != VGPU_ST_RESIDENT || none) return false;
    return all || row_key(e, i, idx_bits) <= kstar;
}

extern "C" __global__ void __launch_bounds__(256) vgpu_victim_count(const VgpuEntry *__restrict__ tbl, uint32_t n,
                                                                    VgpuScanState *st, uint32_t idx_bits, uint32_t chunk) {
    __shared__ uint32_t blk_cnt;
    __shared__ unsigned long long blk_bytes;
    __shared__ bool last;
    __shared__ uint32_t part[256];
    const bool all = st->insufficient != 0;
    const uint64_t kstar = st->prefix;
    const bool none = (st->need == 0);
    const uint32_t lo = blockIdx.x * chunk;
    const uint32_t hi = min(n, lo + chunk);
    if (threadIdx.x == 0) { blk_cnt = 0; blk_bytes = 0; }
    __syncthreads();
    uint32_t cnt = 0;
    unsigned long long bytes = 0;
    for (uint32_t i0 = lo + threadIdx.x; i0 < hi; i0 += 4 * blockDim.x) {  // counting needs no order: coalesced stride, 4 loads in flight
        VgpuEntry e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t i = i0 + u * blockDim.x;
            if (i < hi) e[u] = load_row(tbl, i); else e[u].state = VGPU_ST_FREE;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (victim_row(e[u], i0 + u * blockDim.x, idx_bits, all, none, kstar)) { cnt++; bytes += e[u].size; }
    }
    for (int d = 16; d; d >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, d); bytes += __shfl_down_sync(0xffffffffu, bytes, d); }
    if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&blk_cnt, cnt); atomicAdd(&blk_bytes, bytes); }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->cta_off[blockIdx.x] = blk_cnt;
        st->cta_bytes[blockIdx.x] = blk_bytes;
        __threadfence();
        last = (atomicAdd(&st->done_ctas, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // last CTA: exclusive scan of gridDim.x (<= 1024) counts; thread t owns entries [t*per, (t+1)*per)
    volatile uint32_t *off = st->cta_off;
    volatile uint64_t *cb = st->cta_bytes;
    const uint32_t g = gridDim.x;
    const uint32_t per = (g + blockDim.x - 1) / blockDim.x;
    const uint32_t a = min(g, threadIdx.x * per), b = min(g, a + per);
    uint32_t mine = 0;
    unsigned long long mybytes = 0;
    for (uint32_t k = a; k < b; k++) { mine += off[k]; mybytes += cb[k]; }
    part[threadIdx.x] = mine;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {                                    // Hillis-Steele over 256 partials
        uint32_t v = threadIdx.x >= (uint32_t)d ? part[threadIdx.x - d] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - mine;
    for (uint32_t k = a; k < b; k++) { uint32_t c = off[k]; off[k] = run; run += c; }
    if (mybytes) atomicAdd(reinterpret_cast<unsigned long long *>(&st->out_freed), mybytes);
    if (threadIdx.x == blockDim.x - 1) st->out_count = part[threadIdx.x];
    if (threadIdx.x == 0) st->done_ctas = 0;
}

extern "C" __global__ void __launch_bounds__(256) vgpu_victim_emit(const VgpuEntry *__restrict__ tbl, uint32_t n,
                                                                   VgpuScanState *st, uint32_t idx_bits,
                                                                   uint32_t chunk, uint32_t *__restrict__ out_idx,
                                                                   uint32_t out_cap) {
    __shared__ uint32_t warp_cnt[8];
    const bool all = st->insufficient != 0;
    const uint64_t kstar = st->prefix;
    const bool none = (st->need == 0);
    const uint32_t lo = blockIdx.x * chunk;
    const uint32_t hi = min(n, lo + chunk);
    // each thread owns a contiguous run of rows so that thread order == index order
    const uint32_t per = (chunk + blockDim.x - 1) / blockDim.x;
    const uint32_t tlo = min(hi, lo + threadIdx.x * per);
    const uint32_t thi = min(hi, tlo + per);
    uint32_t cnt = 0;
    for (uint32_t i = tlo; i < thi; i++) {
        VgpuEntry e = load_row(tbl, i);
        if (victim_row(e, i, idx_bits, all, none, kstar)) cnt++;
    }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) warp_cnt[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t w = 0; w < wid; w++) woff += warp_cnt[w];
    uint32_t o = st->cta_off[blockIdx.x] + woff + (incl - cnt);
    if (!cnt) return;
    for (uint32_t i = tlo; i < thi; i++) {
        VgpuEntry e = load_row(tbl, i);
        if (victim_row(e, i, idx_bits, all, none, kstar)) { if (o < out_cap) out_idx[o] = i; o++; }
    }
}

// Whole scan in ONE launch for the tables a container really has (<= 8192 rows; SURVEY.md §8a6: 10^2-10^4 live
// entries): one CTA of 1024 threads keeps its 8 rows per thread in registers (key + size), runs every radix digit
// against a shared-memory histogram with __syncthreads between digits, then scans and emits in row order. Same
// selection rule and outputs as the multi-launch path above; replaces 7 launches (init, 4 digits, count, emit) by one.
extern "C" __global__ void __launch_bounds__(1024) vgpu_victim_small(const VgpuEntry *__restrict__ tbl, uint32_t n, VgpuScanState *st,
                                                                     uint64_t need, uint32_t idx_bits, uint32_t key_bits,
                                                                     uint32_t *__restrict__ out_idx, uint32_t out_cap) {
    constexpr int R = VGPU_SCAN_SMALL_ROWS_PER_THREAD;
    __shared__ unsigned long long hist[VGPU_SCAN_BINS];
    __shared__ uint64_t s_prefix, s_need_left, s_cand;
    __shared__ uint32_t s_insufficient;
    __shared__ uint32_t warp_cnt[32];
    __shared__ unsigned long long s_bytes;
    uint64_t key[R], size[R];
    uint32_t valid = 0;
    const uint32_t base = threadIdx.x * R;
#pragma unroll
    for (int u = 0; u < R; u++) {
        uint32_t i = base + u;
        key[u] = 0; size[u] = 0;
        if (i < n) {
            VgpuEntry e = load_row(tbl, i);
            if (e.state == VGPU_ST_RESIDENT) { valid |= 1u << u; key[u] = row_key(e, i, idx_bits); size[u] = e.size; }
        }
    }
    if (threadIdx.x == 0) { s_prefix = 0; s_need_left = need; s_insufficient = 0; s_bytes = 0; s_cand = 0; }
    __syncthreads();
    for (int hi = (int)key_bits; hi > 0; hi -= VGPU_SCAN_DIGIT_BITS) {
        const uint32_t shift = hi > VGPU_SCAN_DIGIT_BITS ? (uint32_t)(hi - VGPU_SCAN_DIGIT_BITS) : 0u;
        const uint32_t width = (uint32_t)hi - shift;
        const uint32_t mask = (1u << width) - 1u;
        for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint64_t prefix = s_prefix;
#pragma unroll
        for (int u = 0; u < R; u++) {
            if (!(valid >> u & 1u)) continue;
            uint64_t top = hi >= 64 ? 0 : (key[u] >> hi);
            if (top != prefix) continue;
            atomicAdd(&hist[static_cast<uint32_t>(key[u] >> shift) & mask], static_cast<unsigned long long>(size[u]));
        }
        __syncthreads();
        if (threadIdx.x < 32) {       // warp 0 picks the digit: lane l owns bins [l*64, l*64+64)
            constexpr int PER = VGPU_SCAN_BINS / 32;
            volatile unsigned long long *gh = hist;
            uint64_t mine = 0;
            for (int j = 0; j < PER; j++) mine += gh[threadIdx.x * PER + j];
            uint64_t incl = mine;
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (static_cast<int>(threadIdx.x) >= d) incl += o;
            }
            uint64_t excl = incl - mine;
            uint64_t total = __shfl_sync(0xffffffffu, incl, 31);
            uint64_t nl = s_need_left;
            if (total < nl) {
                if (threadIdx.x == 0) { s_insufficient = 1; s_cand = total; }
            } else if (excl < nl && nl <= incl) {
                uint64_t cum = excl;
                int b = threadIdx.x * PER;
                for (int j = 0; j < PER; j++, b++) {
                    uint64_t h = gh[b];
                    if (cum + h >= nl) break;
                    cum += h;
                }
                s_prefix = (prefix << width) | static_cast<uint64_t>(b);
                s_need_left = nl - cum;
            }
        }
        __syncthreads();
        if (s_insufficient) break;    // uniform: read after the barrier
    }
    const bool all = s_insufficient != 0;
    const uint64_t kstar = s_prefix;
    uint32_t cnt = 0;
    unsigned long long bytes = 0;
    uint32_t pick = 0;
#pragma unroll
    for (int u = 0; u < R; u++)
        if ((valid >> u & 1u) && (all || key[u] <= kstar)) { pick |= 1u << u; cnt++; bytes += size[u]; }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    for (int d = 16; d; d >>= 1) bytes += __shfl_down_sync(0xffffffffu, bytes, d);
    if (lane == 31) warp_cnt[wid] = incl;
    if (lane == 0 && bytes) atomicAdd(&s_bytes, bytes);
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; w++) { if (w < wid) woff += warp_cnt[w]; total += warp_cnt[w]; }
    uint32_t o = woff + (incl - cnt);
#pragma unroll
    for (int u = 0; u < R; u++)
        if (pick >> u & 1u) { if (o < out_cap) out_idx[o] = base + u; o++; }
    if (threadIdx.x == 0) {
        st->prefix = kstar; st->need = need; st->need_left = s_need_left; st->cand_bytes = s_cand;
        st->insufficient = s_insufficient; st->out_count = total; st->out_freed = s_bytes; st->done_ctas = 0;
    }
}

// Small control-plane copies (table rows up, scan results down) done by the SMs straight from / into pinned host memory:
// a cuMemcpyAsync of a few KiB would queue behind tens of MiB of page traffic on the copy engines (they serve copies in
// submission order across streams) and stall the pager for milliseconds; a load/store over PCIe does not.
extern "C" __global__ void vgpu_copy16(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint64_t n16) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Tables above 8192 rows, ONE launch: the multi-launch path above costs a launch per digit plus two for the
// ordered compaction and re-reads the 32-byte rows from L2/HBM in every pass (7 launches, 117 us cold for 1 Mi rows).
// Here every CTA of a co-resident grid (cooperative launch, one CTA of 1024 threads per SM, up to 8192 rows each) loads its
// slice ONCE into registers (key + size, like vgpu_victim_small) and keeps it there across the digit passes:
//   per digit : shared-memory histogram of the slice -> non-zero bins added to a global histogram -> grid barrier; the LAST
//               CTA to arrive picks the digit (2048 bins, two per thread, block scan), clears the histogram it read and
//               only then releases the barrier. The passes alternate between two global histograms, so one barrier per
//               digit suffices.
//   emit      : per-CTA count of the selected rows -> grid barrier -> every CTA sums the counts of the CTAs before it and
//               writes its rows' indices in order (thread order == index order).
// Same selection rule and outputs as the other two paths (exact LRU prefix, ascending index).
__device__ __forceinline__ void grid_arrive_and_wait(VgpuScanState *st, unsigned nctas, unsigned &gen, bool *is_last_smem) {
    __threadfence();                                 // this thread's histogram atomics / count writes, before the CTA reports in
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned ticket = atomicAdd(&st->bar_count, 1u);
        *is_last_smem = (ticket == nctas - 1);
        if (!*is_last_smem) {
            while (*reinterpret_cast<volatile unsigned *>(&st->bar_gen) == gen) { __nanosleep(32); }
            __threadfence();
        }
    }
    gen++;
    __syncthreads();
}
__device__ __forceinline__ void grid_release(VgpuScanState *st) {      // by the last arriver, after its exclusive work
    __syncthreads();
    if (threadIdx.x == 0) { st->bar_count = 0; __threadfence(); atomicAdd(&st->bar_gen, 1u); }
}

extern "C" __global__ void __launch_bounds__(1024, 1) vgpu_victim_persist(const VgpuEntry *__restrict__ tbl, uint32_t n, VgpuScanState *st,
                                                                          uint64_t need, uint32_t idx_bits, uint32_t key_bits,
                                                                          uint32_t *__restrict__ out_idx, uint32_t out_cap) {
    constexpr int R = VGPU_SCAN_SMALL_ROWS_PER_THREAD;
    __shared__ unsigned long long hist[VGPU_SCAN_BINS];
    __shared__ unsigned long long warp_sum[32];
    __shared__ uint32_t warp_cnt[32];
    __shared__ bool is_last;
    __shared__ unsigned long long s_bytes;
    const unsigned nctas = gridDim.x;
    unsigned gen = 0;
    if (threadIdx.x == 0) gen = *reinterpret_cast<volatile unsigned *>(&st->bar_gen);   // no CTA passes a barrier before all have arrived
    uint64_t key[R], size[R];
    uint32_t valid = 0;
    const uint32_t base = blockIdx.x * VGPU_SCAN_PERSIST_ROWS_PER_CTA + threadIdx.x * R;
#pragma unroll
    for (int u = 0; u < R; u++) {
        uint32_t i = base + u;
        key[u] = 0; size[u] = 0;
        if (i < n) {
            VgpuEntry e = load_row(tbl, i);
            if (e.state == VGPU_ST_RESIDENT) { valid |= 1u << u; key[u] = row_key(e, i, idx_bits); size[u] = e.size; }
        }
    }
    uint64_t prefix = 0, need_left = need;
    uint32_t insufficient = 0;
    int pass = 0;
    for (int hi = (int)key_bits; hi > 0 && !insufficient; hi -= VGPU_SCAN_DIGIT_BITS, pass++) {
        const uint32_t shift = hi > VGPU_SCAN_DIGIT_BITS ? (uint32_t)(hi - VGPU_SCAN_DIGIT_BITS) : 0u;
        const uint32_t width = (uint32_t)hi - shift;
        const uint32_t mask = (1u << width) - 1u;
        unsigned long long *ghist = reinterpret_cast<unsigned long long *>((pass & 1) ? st->hist2 : st->hist);
        for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x) hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < R; u++) {
            if (!(valid >> u & 1u)) continue;
            uint64_t top = hi >= 64 ? 0 : (key[u] >> hi);
            if (top != prefix) continue;
            atomicAdd(&hist[static_cast<uint32_t>(key[u] >> shift) & mask], static_cast<unsigned long long>(size[u]));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < VGPU_SCAN_BINS; i += blockDim.x)
            if (hist[i]) atomicAdd(&ghist[i], hist[i]);
        grid_arrive_and_wait(st, nctas, gen, &is_last);
        if (is_last) {
            // pick the digit: thread t owns bins 2t and 2t+1 of the global histogram
            volatile unsigned long long *gh = ghist;
            unsigned long long h0 = gh[2 * threadIdx.x], h1 = gh[2 * threadIdx.x + 1];
            unsigned long long mine = h0 + h1, incl = mine;
            const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            for (int d = 1; d < 32; d <<= 1) { unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += o; }
            if (lane == 31) warp_sum[wid] = incl;
            __syncthreads();
            unsigned long long woff = 0, total = 0;
            for (unsigned w = 0; w < 32; w++) { if (w < wid) woff += warp_sum[w]; total += warp_sum[w]; }
            incl += woff;
            unsigned long long excl = incl - mine;
            if (threadIdx.x == 0) { st->insufficient = total < need_left ? 1u : 0u; if (total < need_left) st->cand_bytes = total; }
            if (total >= need_left && excl < need_left && need_left <= incl) {       // exactly one thread (need_left >= 1)
                unsigned b = 2 * threadIdx.x;
                unsigned long long cum = excl;
                if (cum + h0 < need_left) { cum += h0; b++; }
                st->prefix = (prefix << width) | static_cast<uint64_t>(b);
                st->need_left = need_left - cum;
            }
            gh[2 * threadIdx.x] = 0; gh[2 * threadIdx.x + 1] = 0;     // leave the histogram clean for the pass after next
            __threadfence();
            grid_release(st);
        }
        __syncthreads();
        // every CTA picks up the decision (written before the release, read after the barrier)
        prefix = *reinterpret_cast<volatile uint64_t *>(&st->prefix);
        need_left = *reinterpret_cast<volatile uint64_t *>(&st->need_left);
        insufficient = *reinterpret_cast<volatile uint32_t *>(&st->insufficient);
    }
    const bool all = insufficient != 0;
    const bool none = need == 0;
    const uint64_t kstar = prefix;
    uint32_t cnt = 0, pick = 0;
    unsigned long long bytes = 0;
#pragma unroll
    for (int u = 0; u < R; u++)
        if (!none && (valid >> u & 1u) && (all || key[u] <= kstar)) { pick |= 1u << u; cnt++; bytes += size[u]; }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    for (int d = 16; d; d >>= 1) bytes += __shfl_down_sync(0xffffffffu, bytes, d);
    if (threadIdx.x == 0) s_bytes = 0;
    __syncthreads();
    if (lane == 31) warp_cnt[wid] = incl;
    if (lane == 0 && bytes) atomicAdd(&s_bytes, bytes);
    __syncthreads();
    uint32_t woff = 0, cta_total = 0;
    for (uint32_t w = 0; w < 32; w++) { if (w < wid) woff += warp_cnt[w]; cta_total += warp_cnt[w]; }
    if (threadIdx.x == 0) { st->cta_off[blockIdx.x] = cta_total; st->cta_bytes[blockIdx.x] = s_bytes; }
    grid_arrive_and_wait(st, nctas, gen, &is_last);
    if (is_last) grid_release(st);                    // nothing exclusive to do: the counts are all there
    __syncthreads();
    // offset of this CTA = counts of the CTAs before it (<= 148 values: one warp)
    __shared__ uint32_t s_off;
    __shared__ unsigned long long s_tot_bytes;
    __shared__ uint32_t s_tot;
    if (threadIdx.x < 32) {
        uint32_t before = 0, tot = 0;
        unsigned long long tb = 0;
        volatile uint32_t *co = st->cta_off;
        volatile uint64_t *cb = st->cta_bytes;
        for (uint32_t c = threadIdx.x; c < nctas; c += 32) { uint32_t v = co[c]; if (c < blockIdx.x) before += v; tot += v; tb += cb[c]; }
        for (int d = 16; d; d >>= 1) { before += __shfl_down_sync(0xffffffffu, before, d); tot += __shfl_down_sync(0xffffffffu, tot, d); tb += __shfl_down_sync(0xffffffffu, tb, d); }
        if (threadIdx.x == 0) { s_off = before; s_tot = tot; s_tot_bytes = tb; }
    }
    __syncthreads();
    uint32_t o = s_off + woff + (incl - cnt);
#pragma unroll
    for (int u = 0; u < R; u++)
        if (pick >> u & 1u) { if (o < out_cap) out_idx[o] = base + u; o++; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->need = need; st->out_count = s_tot; st->out_freed = s_tot_bytes; st->done_ctas = 0;
        if (!insufficient) st->cand_bytes = 0;
    }
}

// ------------------------------------------------------------------------------------------------ limiter stamp
extern "C" __global__ void vgpu_stamp(uint64_t *slot) {
    uint64_t t = globaltimer();
    *reinterpret_cast<volatile uint64_t *>(slot) = t;
    __threadfence_system();
}

// ------------------------------------------------------------------------------------------------ workload
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// word j of buffer i = splitmix64((i << 32) + j)   (SURVEY.md §8d cfg 3)
extern "C" __global__ void vgpu_wl_fill(uint64_t *buf, uint64_t nwords, uint64_t buf_index) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nwords; j += (uint64_t)gridDim.x * blockDim.x)
        buf[j] = splitmix64((buf_index << 32) + j);
}
// the "touch": read-modify-write x += 1 over the whole buffer, 16 bytes per thread per step
extern "C" __global__ void vgpu_wl_touch(uint64_t *buf, uint64_t nwords) {
    uint64_t n2 = nwords >> 1;
    ulonglong2 *b2 = reinterpret_cast<ulonglong2 *>(buf);
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n2; j += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 v = b2[j];
        v.x += 1; v.y += 1;
        b2[j] = v;
    }
    if ((nwords & 1) && blockIdx.x == 0 && threadIdx.x == 0) buf[nwords - 1] += 1;
}
// the touch through a POINTER TABLE in device memory (what cuBLAS batched GEMM, torch._foreach_* and NCCL do with their
// operands): the launch parameters hold the table's address only, so no argument scan can see the buffers
extern "C" __global__ void vgpu_wl_touch_indirect(uint64_t *const *__restrict__ table, uint32_t nptr, uint64_t nwords) {
    for (uint32_t b = blockIdx.y; b < nptr; b += gridDim.y) {
        uint64_t *buf = table[b];
        for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nwords; j += (uint64_t)gridDim.x * blockDim.x) buf[j] += 1;
    }
}
extern "C" __global__ void vgpu_wl_verify(const uint64_t *buf, uint64_t nwords, uint64_t buf_index, uint64_t added,
                                          unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nwords; j += (uint64_t)gridDim.x * blockDim.x)
        bad += (buf[j] != splitmix64((buf_index << 32) + j) + added);
    if (bad) atomicAdd(mismatches, bad);
}
extern "C" __global__ void vgpu_wl_empty() {}
