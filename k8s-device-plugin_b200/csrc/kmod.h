// kmod.h — the embedded sm_100a cubin (kernels.cu) loaded into the CURRENT context with the driver API, plus
// host-side launch helpers for the pack/unpack and victim-scan kernels.
#pragma once
#include <cuda.h>

#include <cstdint>
#include <vector>

#include "kernels.h"

namespace vgpu {

struct Kernels {
    CUmodule mod = nullptr;
    CUfunction pack_tma = nullptr, pack_generic = nullptr;
    CUfunction victim_init = nullptr, victim_hist = nullptr, victim_emit = nullptr, victim_count = nullptr, victim_small = nullptr, victim_persist = nullptr;
    CUfunction stamp = nullptr, copy16 = nullptr;
    CUfunction wl_fill = nullptr, wl_touch = nullptr, wl_verify = nullptr, wl_empty = nullptr, wl_touch_indirect = nullptr;
    int sm_count = 0;
    bool cooperative = false;                    // the device supports cooperative launches (vgpu_victim_persist)
};

// Loads (once per CUcontext) and returns the kernels for the calling thread's current context. Returns nullptr and
// logs an ERROR when there is no current context or the cubin cannot be loaded (wrong architecture): callers fail
// loudly — there is no CPU fallback anywhere in the product path.
const Kernels *kernels_for_current_ctx();

struct PackSegment { CUdeviceptr src, dst; uint64_t bytes; };

// launch geometry of vgpu_pack_tma (process-wide; tuned by scripts/pack_sweep.py, overridable for experiments)
struct PackConfig { uint32_t tile_bytes = VGPU_PACK_TILE_BYTES; uint32_t stages = VGPU_PACK_STAGES; uint32_t ctas_per_sm = VGPU_PACK_CTAS_PER_SM; };
PackConfig &pack_config();

// Enqueues the copies described by segs on `stream` (src -> dst for every segment). 16-byte-aligned segments go
// through vgpu_pack_tma, the rest through vgpu_pack_generic. launches_out (optional) += number of kernel launches.
CUresult launch_pack(const Kernels *k, const PackSegment *segs, size_t nseg, CUstream stream, int *launches_out = nullptr,
                     CUdeviceptr span = 0);   // span: optional {min start, max end} %globaltimer slot, pre-set to {~0, 0}

// dst/src: device addresses (device memory or the device view of pinned host memory), bytes a multiple of 16. Done by a
// kernel, not by a copy engine: see vgpu_copy16 in kernels.cu.
CUresult launch_copy16(const Kernels *k, CUdeviceptr dst, CUdeviceptr src, size_t bytes, CUstream stream);

// Exact-LRU victim selection on the GPU (see kernels.cu). Owns its device/pinned scratch.
class VictimScanner {
   public:
    ~VictimScanner();
    CUresult init(const Kernels *k, uint32_t max_rows);
    // d_tbl: device table of n rows. need >= 1 bytes. max_touch: upper bound of last_touch values (sizes the key).
    // Synchronises `stream`. victims (ascending row index); *freed = their size sum; returns CUDA_SUCCESS, and
    // *insufficient = true when even all candidates do not reach `need` (then victims = all candidates).
    CUresult scan(CUdeviceptr d_tbl, uint32_t n, uint64_t need, uint64_t max_touch, CUstream stream,
                  std::vector<uint32_t> *victims, uint64_t *freed, bool *insufficient, int *launches_out = nullptr);

   private:
    const Kernels *k_ = nullptr;
    CUdeviceptr d_state_ = 0, d_out_ = 0, dh_state_ = 0, dh_out_ = 0;   // dh_*: device view of the pinned result buffers
    uint32_t cap_ = 0;
    bool persist_refused_ = false;   // the context refused the cooperative launch once: stay on the multi-launch path
    void *h_state_ = nullptr;   // pinned: VgpuScanState header readback
    uint32_t *h_out_ = nullptr; // pinned
};

}  // namespace vgpu
