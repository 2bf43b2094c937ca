// kmod.cc — see kmod.h.
#include "kmod.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "driver.h"
#include "log.h"

extern "C" {
extern const unsigned char vgpu_kernels_cubin[];      // generated: bin2c of kernels.cu's sm_100a cubin
extern const unsigned long long vgpu_kernels_cubin_size;
}

namespace vgpu {

PackConfig &pack_config() {
    static PackConfig c = [] {
        PackConfig d;
        if (const char *e = std::getenv("VGPU_PACK_TILE_KB")) d.tile_bytes = (uint32_t)std::atoi(e) * 1024u;
        if (const char *e = std::getenv("VGPU_PACK_STAGES")) d.stages = (uint32_t)std::atoi(e);
        if (const char *e = std::getenv("VGPU_PACK_CTAS_PER_SM")) d.ctas_per_sm = (uint32_t)std::atoi(e);
        return d;
    }();
    return c;
}

static std::mutex g_mu;
static std::map<CUcontext, Kernels *> g_by_ctx;

const Kernels *kernels_for_current_ctx() {
    const DriverTable &d = drv();
    if (!d.loaded) { LOG_ERROR("CUDA driver not loaded: kernels unavailable"); return nullptr; }
    CUcontext ctx = nullptr;
    if (d.cuCtxGetCurrent(&ctx) != CUDA_SUCCESS || !ctx) { LOG_ERROR("no current CUDA context: kernels unavailable"); return nullptr; }
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_by_ctx.find(ctx);
    if (it != g_by_ctx.end()) return it->second;
    Kernels *k = new Kernels();
    CUresult r = d.cuModuleLoadData(&k->mod, vgpu_kernels_cubin);
    if (r != CUDA_SUCCESS) {
        LOG_ERROR("cuModuleLoadData(sm_100a cubin, %llu bytes) failed: %d %s — this library requires a B200 (sm_100a) device",
                  vgpu_kernels_cubin_size, (int)r, cu_err(r));
        delete k;
        g_by_ctx[ctx] = nullptr;
        return nullptr;
    }
    struct { const char *name; CUfunction *fn; } tab[] = {
        {"vgpu_pack_tma", &k->pack_tma}, {"vgpu_pack_generic", &k->pack_generic},
        {"vgpu_victim_init", &k->victim_init}, {"vgpu_victim_hist", &k->victim_hist}, {"vgpu_victim_emit", &k->victim_emit}, {"vgpu_victim_count", &k->victim_count}, {"vgpu_victim_small", &k->victim_small}, {"vgpu_victim_persist", &k->victim_persist},
        {"vgpu_stamp", &k->stamp}, {"vgpu_copy16", &k->copy16}, {"vgpu_wl_fill", &k->wl_fill}, {"vgpu_wl_touch", &k->wl_touch},
        {"vgpu_wl_verify", &k->wl_verify}, {"vgpu_wl_empty", &k->wl_empty}, {"vgpu_wl_touch_indirect", &k->wl_touch_indirect},
    };
    for (auto &t : tab) {
        r = d.cuModuleGetFunction(t.fn, k->mod, t.name);
        if (r != CUDA_SUCCESS) { LOG_ERROR("kernel %s missing from cubin: %d", t.name, (int)r); delete k; g_by_ctx[ctx] = nullptr; return nullptr; }
    }
    CUdevice dev = 0;
    d.cuCtxGetDevice(&dev);
    d.cuDeviceGetAttribute(&k->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    if (k->sm_count <= 0) k->sm_count = 148;
    int coop = 0;
    d.cuDeviceGetAttribute(&coop, CU_DEVICE_ATTRIBUTE_COOPERATIVE_LAUNCH, dev);
    k->cooperative = coop != 0 && d.cuLaunchCooperativeKernel != nullptr;
    const int smem = 227 * 1024;   // opt in to the maximum; each launch requests 128 + stages * tile bytes
    r = d.cuFuncSetAttribute(k->pack_tma, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
    if (r != CUDA_SUCCESS) LOG_ERROR("cuFuncSetAttribute(max dynamic smem %d) failed: %d", smem, (int)r);
    g_by_ctx[ctx] = k;
    return k;
}

static CUresult launch_pack_set(const Kernels *k, CUfunction fn, bool tma, const PackSegment *segs, const uint32_t *which,
                                size_t n, CUstream stream, int *launches, CUdeviceptr span) {
    const DriverTable &d = drv();
    size_t i = 0;
    PackConfig cfg = pack_config();
    if (cfg.stages < 2) cfg.stages = 2;
    if (cfg.stages > VGPU_PACK_MAX_STAGES) cfg.stages = VGPU_PACK_MAX_STAGES;
    if (cfg.tile_bytes < 1024 || (cfg.tile_bytes & 15u)) cfg.tile_bytes = VGPU_PACK_TILE_BYTES;
    while (128u + cfg.stages * cfg.tile_bytes > 227u * 1024u) cfg.stages--;
    if (cfg.ctas_per_sm < 1) cfg.ctas_per_sm = 1;
    const uint32_t tile = tma ? cfg.tile_bytes : 32u * 1024u;
    while (i < n) {
        VgpuPackParams p;
        p.tile_bytes = tile;
        p.stages = cfg.stages;
        p._pad = 0;
        p.span = span;
        p.nseg = 0;
        uint64_t tiles = 0;
        while (i < n && p.nseg < VGPU_PACK_MAX_SEG) {
            const PackSegment &s = segs[which[i]];
            VgpuPackSeg &o = p.seg[p.nseg++];
            o.src = s.src; o.dst = s.dst; o.bytes = s.bytes; o.tile_begin = tiles;
            tiles += (s.bytes + tile - 1) / tile;
            i++;
        }
        p.total_tiles = tiles;
        if (tiles == 0) continue;
        void *args[] = {&p};
        CUresult r;
        if (tma) {
            // persistent grid: a multiple of the SM count, every CTA loops over tiles t = blockIdx, blockIdx + grid, ...
            uint64_t want = (uint64_t)k->sm_count * cfg.ctas_per_sm;
            unsigned grid = (unsigned)(tiles < want ? tiles : want);
            r = d.cuLaunchKernel(fn, grid, 1, 1, 32, 1, 1, 128 + cfg.stages * cfg.tile_bytes, stream, args, nullptr);
        } else {
            uint64_t want = tiles < (uint64_t)k->sm_count * 8 ? tiles : (uint64_t)k->sm_count * 8;
            r = d.cuLaunchKernel(fn, (unsigned)want, 1, 1, 256, 1, 1, 0, stream, args, nullptr);
        }
        if (r != CUDA_SUCCESS) { LOG_ERROR("pack launch failed: %d %s", (int)r, cu_err(r)); return r; }
        if (launches) (*launches)++;
    }
    return CUDA_SUCCESS;
}

CUresult launch_pack(const Kernels *k, const PackSegment *segs, size_t nseg, CUstream stream, int *launches_out, CUdeviceptr span) {
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    std::vector<uint32_t> al, un;
    for (size_t i = 0; i < nseg; i++) {
        if (segs[i].bytes == 0) continue;
        bool aligned = ((segs[i].src | segs[i].dst | segs[i].bytes) & 15u) == 0;
        (aligned ? al : un).push_back((uint32_t)i);
    }
    CUresult r = CUDA_SUCCESS;
    if (!al.empty()) r = launch_pack_set(k, k->pack_tma, true, segs, al.data(), al.size(), stream, launches_out, span);
    if (r == CUDA_SUCCESS && !un.empty()) r = launch_pack_set(k, k->pack_generic, false, segs, un.data(), un.size(), stream, launches_out, 0);
    return r;
}

CUresult launch_copy16(const Kernels *k, CUdeviceptr dst, CUdeviceptr src, size_t bytes, CUstream stream) {
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    if (bytes == 0) return CUDA_SUCCESS;
    if ((dst | src | bytes) & 15u) return CUDA_ERROR_INVALID_VALUE;
    uint64_t n16 = bytes / 16;
    unsigned grid = (unsigned)std::min<uint64_t>((n16 + 255) / 256, (uint64_t)k->sm_count * 4);
    void *a[] = {&dst, &src, &n16};
    return drv().cuLaunchKernel(k->copy16, grid, 1, 1, 256, 1, 1, 0, stream, a, nullptr);
}

VictimScanner::~VictimScanner() {
    const DriverTable &d = drv();
    if (d_state_) d.cuMemFree_v2(d_state_);
    if (d_out_) d.cuMemFree_v2(d_out_);
    if (h_state_) d.cuMemFreeHost(h_state_);
    if (h_out_) d.cuMemFreeHost(h_out_);
}

CUresult VictimScanner::init(const Kernels *k, uint32_t max_rows) {
    const DriverTable &d = drv();
    k_ = k;
    cap_ = max_rows;
    CUresult r;
    if ((r = d.cuMemAlloc_v2(&d_state_, sizeof(VgpuScanState))) != CUDA_SUCCESS) return r;
    if ((r = d.cuMemsetD8_v2(d_state_, 0, sizeof(VgpuScanState))) != CUDA_SUCCESS) return r;   // vgpu_victim_persist's invariant: histograms and barrier start at zero
    if ((r = d.cuMemAlloc_v2(&d_out_, ((size_t)cap_ * 4 + 15) & ~(size_t)15)) != CUDA_SUCCESS) return r;
    if ((r = d.cuMemHostAlloc(&h_state_, 64, CU_MEMHOSTALLOC_DEVICEMAP)) != CUDA_SUCCESS) return r;
    if ((r = d.cuMemHostAlloc((void **)&h_out_, ((size_t)cap_ * 4 + 15) & ~(size_t)15, CU_MEMHOSTALLOC_DEVICEMAP)) != CUDA_SUCCESS) return r;
    // the results travel down by kernel stores into pinned memory, not by a copy engine (see vgpu_copy16)
    if ((r = d.cuMemHostGetDevicePointer_v2(&dh_state_, h_state_, 0)) != CUDA_SUCCESS) return r;
    if ((r = d.cuMemHostGetDevicePointer_v2(&dh_out_, h_out_, 0)) != CUDA_SUCCESS) return r;
    return CUDA_SUCCESS;
}

static bool multilaunch_only() { static const bool v = std::getenv("VGPU_SCAN_MULTILAUNCH") != nullptr; return v; }   // experiments / tests of the multi-launch path
static uint32_t bit_length(uint64_t v) { uint32_t b = 0; while (v) { b++; v >>= 1; } return b; }

CUresult VictimScanner::scan(CUdeviceptr d_tbl, uint32_t n, uint64_t need, uint64_t max_touch, CUstream stream,
                             std::vector<uint32_t> *victims, uint64_t *freed, bool *insufficient, int *launches_out) {
    const DriverTable &d = drv();
    victims->clear();
    if (freed) *freed = 0;
    if (insufficient) *insufficient = false;
    if (n == 0 || need == 0) return CUDA_SUCCESS;
    if (n > cap_) return CUDA_ERROR_INVALID_VALUE;
    uint32_t idx_bits = bit_length(n - 1);
    if (idx_bits == 0) idx_bits = 1;
    uint32_t touch_bits = bit_length(max_touch);
    if (touch_bits == 0) touch_bits = 1;
    if (idx_bits + touch_bits > 64) return CUDA_ERROR_INVALID_VALUE;
    uint32_t key_bits = idx_bits + touch_bits;
    int launches = 0;
    CUresult r;
    if (n <= VGPU_SCAN_SMALL_MAX_ROWS && k_->victim_small) {
        // the whole scan in one launch: a single 1024-thread CTA with its rows in registers (kernels.cu)
        void *a[] = {&d_tbl, &n, &d_state_, &need, &idx_bits, &key_bits, &d_out_, &cap_};
        if ((r = d.cuLaunchKernel(k_->victim_small, 1, 1, 1, 1024, 1, 1, 0, stream, a, nullptr)) != CUDA_SUCCESS) return r;
        launches++;
    } else {
    bool persisted = false;
    if (k_->cooperative && k_->victim_persist && !persist_refused_ && n <= (uint64_t)k_->sm_count * VGPU_SCAN_PERSIST_ROWS_PER_CTA && !multilaunch_only()) {
        // one cooperative launch, the table slice of every CTA stays in registers across the digit passes (kernels.cu)
        unsigned grid = (unsigned)((n + VGPU_SCAN_PERSIST_ROWS_PER_CTA - 1) / VGPU_SCAN_PERSIST_ROWS_PER_CTA);
        void *a[] = {&d_tbl, &n, &d_state_, &need, &idx_bits, &key_bits, &d_out_, &cap_};
        r = d.cuLaunchCooperativeKernel(k_->victim_persist, grid, 1, 1, 1024, 1, 1, 0, stream, a);
        if (r == CUDA_SUCCESS) { launches++; persisted = true; }
        else {
            // a context that cannot co-schedule the grid (MPS, a partitioned device): nothing has run — use the launch-per-digit
            // path from now on
            persist_refused_ = true;
            LOG_WARN("cooperative launch of the single-launch victim scan refused (%d %s): using the multi-launch path", (int)r, cu_err(r));
        }
    }
    if (!persisted) {
    {
        void *a[] = {&d_state_, &need};
        if ((r = d.cuLaunchKernel(k_->victim_init, 1, 1, 1, 256, 1, 1, 0, stream, a, nullptr)) != CUDA_SUCCESS) return r;
        launches++;
    }
    // one CTA per 2048 rows, capped at 4 CTAs per SM and at the per-CTA offset table of the ordered emit
    uint32_t rows_per_cta = 256 * 8;
    uint32_t grid = (n + rows_per_cta - 1) / rows_per_cta;
    uint32_t max_grid = (uint32_t)k_->sm_count * 8;   // 8 x 256 threads fill an SM
    if (grid > max_grid) grid = max_grid;
    if (grid > VGPU_SCAN_MAX_CTAS) grid = VGPU_SCAN_MAX_CTAS;
    if (grid == 0) grid = 1;
    for (int hi = (int)key_bits; hi > 0; hi -= VGPU_SCAN_DIGIT_BITS) {
        uint32_t shift = hi > VGPU_SCAN_DIGIT_BITS ? (uint32_t)(hi - VGPU_SCAN_DIGIT_BITS) : 0u;
        uint32_t width = (uint32_t)hi - shift;
        void *a[] = {&d_tbl, &n, &d_state_, &idx_bits, &shift, &width};
        if ((r = d.cuLaunchKernel(k_->victim_hist, grid, 1, 1, 256, 1, 1, 0, stream, a, nullptr)) != CUDA_SUCCESS) return r;
        launches++;
    }
    uint32_t chunk = (n + grid - 1) / grid;
    {
        void *a[] = {&d_tbl, &n, &d_state_, &idx_bits, &chunk};
        if ((r = d.cuLaunchKernel(k_->victim_count, grid, 1, 1, 256, 1, 1, 0, stream, a, nullptr)) != CUDA_SUCCESS) return r;
        launches++;
    }
    {
        void *a[] = {&d_tbl, &n, &d_state_, &idx_bits, &chunk, &d_out_, &cap_};
        if ((r = d.cuLaunchKernel(k_->victim_emit, grid, 1, 1, 256, 1, 1, 0, stream, a, nullptr)) != CUDA_SUCCESS) return r;
        launches++;
    }
    }
    }
    if ((r = launch_copy16(k_, dh_state_, d_state_, 64, stream)) != CUDA_SUCCESS) return r;
    launches += 2;                                        // the two result copies below are kernels too (vgpu_copy16)
    if (launches_out) *launches_out += launches;
    // optimistic first page of indices; the rest (rare) after the count is known
    uint32_t first = cap_ < 1024 ? cap_ : 1024;
    if ((r = launch_copy16(k_, dh_out_, d_out_, ((size_t)first * 4 + 15) & ~(size_t)15, stream)) != CUDA_SUCCESS) return r;
    if ((r = d.cuStreamSynchronize(stream)) != CUDA_SUCCESS) return r;
    const VgpuScanState *hs = static_cast<const VgpuScanState *>(h_state_);
    uint32_t cnt = hs->out_count;
    if (cnt > first) {
        size_t from = (size_t)first * 4 & ~(size_t)15, to = ((size_t)cnt * 4 + 15) & ~(size_t)15;
        if ((r = launch_copy16(k_, dh_out_ + from, d_out_ + from, to - from, stream)) != CUDA_SUCCESS) return r;
        if ((r = d.cuStreamSynchronize(stream)) != CUDA_SUCCESS) return r;
    }
    victims->assign(h_out_, h_out_ + cnt);
    if (freed) *freed = hs->out_freed;
    if (insufficient) *insufficient = hs->insufficient != 0;
    return CUDA_SUCCESS;
}

}  // namespace vgpu
