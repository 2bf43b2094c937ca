// limiter.cc — see limiter.h.
#include "limiter.h"

#include <time.h>
#include <unistd.h>

#include <cstdlib>

#include "driver.h"
#include "log.h"

namespace vgpu {

static uint64_t now_ns() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static void sleep_ns(uint64_t ns) {
    struct timespec ts{(time_t)(ns / 1000000000ull), (long)(ns % 1000000000ull)};
    nanosleep(&ts, nullptr);
}

Limiter::Limiter(int percent, vgpu_shared_region_t *region, int util_policy)
    : percent_(percent), util_policy_(util_policy), region_(region) {
    active_ = percent > 0 && percent < 100;   // rate_limiter returns early for sm_limit >= 100 or == 0 (@0x4591a)
    if (const char *e = std::getenv("VGPU_LIMITER_BURST_US")) burst_ns_ = std::atof(e) * 1e3;
    bucket_ns_ = burst_ns_;
    t0_ = last_refill_ = now_ns();
    st_.limit_percent = percent;
}

Limiter::~Limiter() {
    if (ring_ && drv().loaded) drv().cuMemFreeHost((void *)ring_);
}

bool Limiter::enabled_now() const {
    if (!active_) return false;
    // get_utilization_switch@0x45307: GPU_CORE_UTILIZATION_POLICY force -> on, disable -> off, default -> monitor's word
    if (util_policy_ == 1) return true;
    if (util_policy_ == 2) return false;
    return !region_ || __atomic_load_n(&region_->utilization_switch, __ATOMIC_RELAXED) != 0;
}

bool Limiter::ensure_ring() {
    if (ring_) return true;
    const DriverTable &d = drv();
    void *p = nullptr;
    if (d.cuMemHostAlloc(&p, (size_t)ring_n_ * 8, CU_MEMHOSTALLOC_DEVICEMAP | CU_MEMHOSTALLOC_PORTABLE) != CUDA_SUCCESS) {
        LOG_ERROR("limiter: pinned stamp ring allocation failed; core limit NOT enforced");
        active_ = false;
        return false;
    }
    ring_ = static_cast<volatile uint64_t *>(p);
    for (int i = 0; i < ring_n_; i++) ring_[i] = 0;
    if (d.cuMemHostGetDevicePointer_v2(&d_ring_, p, 0) != CUDA_SUCCESS) d_ring_ = (CUdeviceptr)(uintptr_t)p;
    return true;
}

int Limiter::stamp(CUstream st) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) { active_ = false; LOG_ERROR("limiter: kernels unavailable; core limit NOT enforced"); return -1; }
    int idx = next_;
    next_ = (next_ + 1) % ring_n_;
    ring_[idx] = 0;
    CUdeviceptr slot = d_ring_ + (size_t)idx * 8;
    void *args[] = {&slot};
    CUresult r = drv().cuLaunchKernel(k->stamp, 1, 1, 1, 1, 1, 1, 0, st, args, nullptr);
    if (r != CUDA_SUCCESS) { LOG_WARN("limiter: stamp launch failed: %d", (int)r); return -1; }
    st_.stamps++;
    return idx;
}

void Limiter::refill(uint64_t now) {
    bucket_ns_ += (double)(now - last_refill_) * percent_ / 100.0;
    if (bucket_ns_ > burst_ns_) bucket_ns_ = burst_ns_;
    last_refill_ = now;
}

void Limiter::harvest() {
    for (auto it = pending_.begin(); it != pending_.end();) {
        uint64_t e = ring_[it->end_idx];
        if (e == 0) { ++it; continue; }
        uint64_t b = it->has_begin ? ring_[it->begin_idx] : it->prev_end;
        if (it->has_begin && b == 0) { ++it; continue; }
        PerStream &ps = streams_[it->st];
        if (!it->has_begin && ps.last_end > b) b = ps.last_end;
        uint64_t busy = e > b ? e - b : 0;
        const bool estimated = it->idle_ns && busy > 0;
        if (estimated) {
            // span = work + the host's pause: keep what preceded the pause, cap what followed it at the recent average
            uint64_t est = (uint64_t)(avg_busy_per_launch_ns_ * (it->launches ? it->launches : 1));
            uint64_t before_pause = busy > it->idle_ns ? busy - it->idle_ns : 0;
            uint64_t capped = before_pause + (it->idle_ns < est ? it->idle_ns : est);
            if (capped < busy) busy = capped;
        } else if (it->launches) {
            double per = (double)busy / it->launches;
            avg_busy_per_launch_ns_ = avg_busy_per_launch_ns_ == 0 ? per : 0.8 * avg_busy_per_launch_ns_ + 0.2 * per;
        }
        ps.last_end = e;
        bucket_ns_ -= (double)busy;
        st_.busy_ns += busy;
        st_.groups++;
        // keep the measuring overhead proportional: long kernels -> stamp every launch and keep at most one group
        // in flight (tight duty cycle); short kernels -> stamp every `stride` launches
        if (busy > 1000000ull) { stride_ = 1; max_inflight_ = 1; }
        else if (estimated) { /* not a measurement: no evidence for widening the stride */ }
        else if (busy < 100000ull * (uint64_t)stride_ && stride_ < 32) { stride_ *= 2; max_inflight_ = 4; }
        else if (busy > 400000ull && stride_ > 1) stride_ /= 2;
        it = pending_.erase(it);
    }
}

void Limiter::before_launch(CUstream st) {
    // monitor handshake first (all modes): a higher-priority task blocks us with recent_kernel = -1
    if (region_) {
        // a word shared without a lock with the node monitor and with every thread of every process of the container
        // (the reference does plain loads/stores on it too): relaxed atomics keep it a defined access
        while (__atomic_load_n(&region_->recent_kernel, __ATOMIC_RELAXED) < 0) sleep(1);
        __atomic_store_n(&region_->recent_kernel, 2, __ATOMIC_RELAXED);
    }
    if (!enabled_now()) return;
    std::lock_guard<std::mutex> g(mu_);
    if (!ensure_ring()) return;
    st_.launches++;
    uint64_t t_in = now_ns();
    refill(t_in);
    harvest();
    // throttle: wait until the bucket is positive and the in-flight measurement backlog is bounded. A group that is
    // still open on this stream is closed FIRST: its end stamp then runs right behind the kernels already queued, so
    // the time this thread is about to sleep is not billed as GPU time when the group is finally measured.
    if (bucket_ns_ <= 0 || (int)pending_.size() >= max_inflight_) {
        PerStream &cur = streams_[st];
        if (cur.open && cur.since_end > 0) {
            int e = stamp(st);
            if (e >= 0) pending_.push_back(Group{st, cur.begin_idx, e, cur.last_end, cur.begin_idx >= 0, cur.since_end, 0});
            cur.open = false;
        }
    }
    while (bucket_ns_ <= 0 || (int)pending_.size() >= max_inflight_) {
        if (!pending_.empty()) {
            sleep_ns(20000);
        } else {
            double need = -bucket_ns_ * 100.0 / percent_;
            sleep_ns((uint64_t)(need < 20000 ? 20000 : (need > 2e6 ? 2e6 : need)));
        }
        refill(now_ns());
        harvest();
    }
    uint64_t t_out = now_ns();
    st_.throttle_ns += t_out - t_in;
    PerStream &ps = streams_[st];
    if (!ps.open) {
        // New group. If an earlier group of this stream is still unmeasured, this launch queues directly behind it and
        // that group's end stamp is this group's start; otherwise the stream may have gone idle -> stamp the start.
        ps.open = true;
        ps.since_end = 0;
        bool behind_pending = false;
        for (auto &gq : pending_) if (gq.st == st) behind_pending = true;
        ps.begin_idx = behind_pending ? -1 : stamp(st);
    } else if (t_out - ps.last_launch_ns > 200000) {
        // the APPLICATION paused inside an open group (no end stamp was queued behind its last launch). Close the group
        // now; its span then contains the pause, so harvest() replaces the part of the span that lies after the last
        // launch by an estimate from the recent busy-per-launch average (idle_ns = how long the host was away).
        int e = stamp(st);
        if (e >= 0) pending_.push_back(Group{st, ps.begin_idx, e, ps.last_end, ps.begin_idx >= 0, ps.since_end, t_out - ps.last_launch_ns});
        ps.begin_idx = -1;
        ps.since_end = 0;
        // An estimate was just substituted for a measurement. Measure the next launches one by one (begin/end stamps
        // around each) so that the stride only grows back on evidence that they really are short — otherwise a pause
        // inside a group of LONG launches (the host blocked in a synchronise while the GPU worked) under-bills the group,
        // the low figure keeps the stride high, and the limiter never recovers.
        stride_ = 1;
    }
}

void Limiter::after_launch(CUstream st, bool heavy) {
    if (!enabled_now()) return;
    std::lock_guard<std::mutex> g(mu_);
    if (!ring_) return;
    PerStream &ps = streams_[st];
    if (!ps.open) return;
    ps.last_launch_ns = now_ns();
    // heavy = one call that enqueues an unknown amount of work (a graph launch): always measured on its own
    if (++ps.since_end >= stride_ || heavy) {
        int e = stamp(st);
        if (e >= 0) pending_.push_back(Group{st, ps.begin_idx, e, ps.last_end, ps.begin_idx >= 0, ps.since_end, 0});
        ps.open = false;
    }
}

LimiterStats Limiter::stats() {
    std::lock_guard<std::mutex> g(mu_);
    if (ring_) harvest();
    LimiterStats s = st_;
    s.wall_ns = now_ns() - t0_;
    return s;
}

}  // namespace vgpu
