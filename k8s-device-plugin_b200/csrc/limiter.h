// limiter.h — gpucores (SM percentage) limiter: a token bucket of GPU-busy nanoseconds, debited with DEVICE
// timestamps (%globaltimer stamps written by a one-thread kernel into pinned host memory) and refilled at
// limit% of wall time, enforced on the cuLaunchKernel intercept.
//
// Reference: rate_limiter libvgpu.so@0x4591a + utilization_watcher@0x46710 (multiprocess_utilization_watcher.c):
// a bucket of "cuda cores" decremented by gridDim products and refilled every 120 ms from NVML's 1-second
// per-process smUtil through delta()@0x45c7b — int32 arithmetic that overflows on B200 (148 SMs x 2048 threads;
// SURVEY.md Appendix E), a feedback loop with a >1 s lag, and a 10 ms nanosleep spin when empty. The quota
// (CUDA_DEVICE_SM_LIMIT, percent) and the monitor handshake (recent_kernel / utilization_switch / priority,
// cmd/vGPUmonitor/feedback.go:197-255) are kept; the control law is replaced.
#pragma once
#include <cuda.h>

#include <cstdint>
#include <deque>
#include <map>
#include <mutex>

#include "kmod.h"
#include "vgpu_region.h"

namespace vgpu {

struct LimiterStats {
    uint64_t launches = 0, stamps = 0, groups = 0;
    uint64_t busy_ns = 0;       // device-measured busy time debited so far
    uint64_t throttle_ns = 0;   // host time spent waiting in before_launch
    uint64_t wall_ns = 0;       // since creation
    int limit_percent = 0;
};

class Limiter {
   public:
    // percent in (0,100) enables throttling; anything else makes before/after no-ops apart from the monitor handshake
    Limiter(int percent, vgpu_shared_region_t *region, int util_policy);
    ~Limiter();
    void before_launch(CUstream st);
    void after_launch(CUstream st, bool heavy = false);
    LimiterStats stats();
    bool active() const { return active_; }

   private:
    struct Group { CUstream st; int begin_idx; int end_idx; uint64_t prev_end; bool has_begin; int launches; uint64_t idle_ns; };
    struct PerStream { uint64_t last_end = 0; int since_end = 0; uint64_t last_launch_ns = 0; bool open = false; int begin_idx = -1; };
    bool ensure_ring();
    int stamp(CUstream st);
    void harvest();
    void refill(uint64_t now);
    bool enabled_now() const;

    bool active_ = false;
    int percent_ = 100;
    int util_policy_ = 0;
    vgpu_shared_region_t *region_ = nullptr;
    std::mutex mu_;
    volatile uint64_t *ring_ = nullptr;   // pinned, device-visible
    CUdeviceptr d_ring_ = 0;
    int ring_n_ = 1024, next_ = 0;
    std::deque<Group> pending_;
    std::map<CUstream, PerStream> streams_;
    double bucket_ns_ = 0, burst_ns_ = 5e6;
    uint64_t last_refill_ = 0, t0_ = 0;
    int stride_ = 1, max_inflight_ = 2;
    double avg_busy_per_launch_ns_ = 0;   // from groups measured without a pause inside
    LimiterStats st_;
};

}  // namespace vgpu
