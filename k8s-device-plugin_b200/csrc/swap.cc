// swap.cc — see swap.h.
#include "swap.h"

#include <sched.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "driver.h"
#include "log.h"

namespace vgpu {

#define CU_TRY(expr)                                                                     \
    do {                                                                                 \
        CUresult _r = (expr);                                                            \
        if (_r != CUDA_SUCCESS) {                                                        \
            LOG_ERROR("%s failed: %d %s", #expr, (int)_r, cu_err(_r));                   \
            return _r;                                                                   \
        }                                                                                \
    } while (0)

static uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
static uint64_t mono_ns() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec; }
struct ScopedNs { uint64_t *acc; uint64_t t0; explicit ScopedNs(uint64_t *a) : acc(a), t0(mono_ns()) {} ~ScopedNs() { *acc += mono_ns() - t0; } };
static uint64_t env_u64(const char *name, uint64_t dflt) {
    const char *e = std::getenv(name);
    return e && *e ? std::strtoull(e, nullptr, 0) : dflt;
}

SwapConfig SwapConfig::from_env(uint64_t resident_cap, uint64_t virtual_cap) {
    SwapConfig c;
    c.resident_cap = resident_cap;
    c.virtual_cap = virtual_cap;
    c.host_pool_cap = env_u64("VGPU_SWAP_HOST_POOL_MB", 0) << 20;
    c.chunk_bytes = (size_t)env_u64("VGPU_SWAP_CHUNK_MB", 32) << 20;
    c.ring_slots = (int)env_u64("VGPU_SWAP_RING", 4);
    c.slab_bytes = (size_t)env_u64("VGPU_SWAP_SLAB_MB", 1024) << 20;
    c.arena_bytes = env_u64("VGPU_SWAP_ARENA_GB", 1024) << 30;
    c.profile = env_u64("VGPU_SWAP_PROFILE", 0) != 0;
    c.scan_lookahead = (uint32_t)env_u64("VGPU_SWAP_SCAN_LOOKAHEAD", c.scan_lookahead);
    c.prefetch_bytes = env_u64("VGPU_SWAP_PREFETCH_MB", c.prefetch_bytes >> 20) << 20;
    c.copy_bytes = (size_t)env_u64("VGPU_SWAP_COPY_MB", c.copy_bytes >> 20) << 20;
    c.batch_rows = (uint32_t)env_u64("VGPU_SWAP_BATCH_ROWS", c.batch_rows);
    if (std::getenv("VGPU_SWAP_HEADROOM_MB")) c.headroom_bytes = env_u64("VGPU_SWAP_HEADROOM_MB", 0) << 20;
    c.host_backed = env_u64("VGPU_SWAP_HOST_BACKED", 0) != 0;
    if (c.ring_slots < 2) c.ring_slots = 2;
    if (c.chunk_bytes < (1u << 20)) c.chunk_bytes = 1u << 20;
    if (c.copy_bytes < (1u << 20)) c.copy_bytes = 1u << 20;
    if (c.batch_rows < 1) c.batch_rows = 1;
    return c;
}

// NUMA node of the GPU's PCIe root (sysfs), or -1. On HGX B200 boxes (2 sockets) page traffic that lands in the far
// socket's DRAM crosses the inter-socket link and loses more than half of the host-link bandwidth (profiles/README.md),
// so pinned slabs are allocated with a memory policy preferring the GPU's own node, wherever the caller's thread runs.
static int gpu_numa_node(int dev) {
    char bus[32] = {0};
    CUdevice d = dev;
    auto getbus = reinterpret_cast<CUresult (*)(char *, int, CUdevice)>(real_cuda_symbol("cuDeviceGetPCIBusId"));
    if (!getbus || getbus(bus, sizeof bus, d) != CUDA_SUCCESS) return -1;
    for (char *p = bus; *p; p++) if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    char path[128];
    const char *bdf = bus;
    if (std::strlen(bus) > 12) bdf = bus + (std::strlen(bus) - 12);   // "00000000:9c:00.0" -> "0000:9c:00.0"
    std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = std::fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
// set_mempolicy without libnuma: MPOL_PREFERRED = 1, MPOL_DEFAULT = 0
static void prefer_node(int node) {
    if (node < 0 || node >= 64) return;
    unsigned long mask = 1ul << node;
    syscall(SYS_set_mempolicy, 1, &mask, sizeof(mask) * 8);
}
static void default_policy() { syscall(SYS_set_mempolicy, 0, nullptr, 0); }
// which node did the kernel actually give us? (get_mempolicy MPOL_F_NODE|MPOL_F_ADDR = 3); -1 when unknown
static int node_of(void *addr) {
    int node = -1;
    if (syscall(SYS_get_mempolicy, &node, nullptr, 0, addr, 3) != 0) return -1;
    return node;
}
// The driver pins pages where the CALLING CPU sits (measured: with only a memory policy the pool still landed on
// the far socket, 35-50 GB/s instead of 88), so the thread is moved onto the GPU's node for the duration of the slab
// allocation and then put back. Fails harmlessly inside a cpuset that excludes those CPUs.
struct NodeAffinity {
    cpu_set_t saved;
    bool moved = false;
    explicit NodeAffinity(int node) {
        if (node < 0) return;
        char path[96];
        std::snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        FILE *f = std::fopen(path, "r");
        if (!f) return;
        char buf[1024] = {0};
        if (!std::fgets(buf, sizeof buf, f)) { std::fclose(f); return; }
        std::fclose(f);
        cpu_set_t want;
        CPU_ZERO(&want);
        for (char *tok = std::strtok(buf, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            if (std::sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && c < CPU_SETSIZE; c++) CPU_SET(c, &want); }
            else if (std::sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) CPU_SET(a, &want);
        }
        if (sched_getaffinity(0, sizeof saved, &saved) != 0) return;
        cpu_set_t both;
        CPU_AND(&both, &want, &saved);                 // stay inside whatever cpuset the container was given
        if (CPU_COUNT(&both) == 0) return;
        moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    ~NodeAffinity() { if (moved) sched_setaffinity(0, sizeof saved, &saved); }
};

SwapEngine *SwapEngine::create(int dev, const SwapConfig &cfg) {
    SwapEngine *e = new SwapEngine();
    if (!e->init(dev, cfg)) {
        delete e;
        return nullptr;
    }
    return e;
}

bool SwapEngine::init(int dev, const SwapConfig &cfg) {
    const DriverTable &d = drv();
    dev_ = dev;
    cfg_ = cfg;
    profile_.store(cfg.profile);
    trace_want_ = (uint32_t)env_u64("VGPU_SWAP_TRACE", 0);
    trace_skip_ = (uint32_t)env_u64("VGPU_SWAP_TRACE_SKIP", 1200);
    k_ = kernels_for_current_ctx();
    if (!k_) return false;
    numa_node_ = std::getenv("VGPU_SWAP_NO_NUMA") ? -1 : gpu_numa_node(dev);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev_;
    if (!d.cuMemGetAllocationGranularity || d.cuMemGetAllocationGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS) {
        LOG_ERROR("VMM unavailable (cuMemGetAllocationGranularity)");
        return false;
    }
    cfg_.chunk_bytes = round_up(cfg_.chunk_bytes, gran_);
    if (cfg_.resident_cap == 0) {
        size_t fr = 0, tot = 0;
        d.cuMemGetInfo_v2(&fr, &tot);
        uint64_t overhead = 2ull * cfg_.ring_slots * cfg_.chunk_bytes + (512ull << 20);
        cfg_.resident_cap = fr > overhead ? fr - overhead : fr / 2;
    }
    quota_cap_ = cfg_.resident_cap;
    uint64_t want = round_up(cfg_.arena_bytes, gran_);
    CUresult r = CUDA_ERROR_UNKNOWN;
    // base alignment: buffers sit at multiples of their own (power-of-two-ish) sizes from the arena base, so a base
    // aligned to the upper page-directory span keeps a buffer from straddling page-table pages
    uint64_t align = env_u64("VGPU_SWAP_ARENA_ALIGN_GB", 4) << 30;
    while (want >= (8ull << 30)) {
        r = d.cuMemAddressReserve(&arena_, want, align, 0, 0);
        if (r != CUDA_SUCCESS && align) r = d.cuMemAddressReserve(&arena_, want, 0, 0, 0);
        if (r == CUDA_SUCCESS) break;
        want >>= 1;
    }
    if (r != CUDA_SUCCESS) { LOG_ERROR("cuMemAddressReserve failed: %d %s", (int)r, cu_err(r)); return false; }
    cfg_.arena_bytes = want;
    if (cfg_.host_backed) {
        CUmemAllocationProp hp = {};
        hp.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        hp.location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA;
        hp.location.id = numa_node_ >= 0 ? numa_node_ : 0;
        size_t hg = 0;
        if (d.cuMemGetAllocationGranularity(&hg, &hp, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || hg > gran_ ||
            d.cuMemAddressReserve(&harena_, want, align, 0, 0) != CUDA_SUCCESS) {
            LOG_ERROR("VGPU_SWAP_HOST_BACKED: this driver has no host-located VMM memory (CU_MEM_LOCATION_TYPE_HOST_NUMA); evicted ranges stay unmapped");
            cfg_.host_backed = false;
            harena_ = 0;
        }
    }
    va_free_[0] = want;
    owner_.assign(want / gran_, -1);

    // highest priority: when the application saturates the SMs, the block scheduler hands freed slots to the swap
    // kernels first (they are short and the link is waiting on them)
    int prio_lo = 0, prio_hi = 0;
    if (d.cuCtxGetStreamPriorityRange) d.cuCtxGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (CUstream *s : {&s_scan_, &s_pack_, &s_unpack_, &s_out_, &s_in_}) {
        CUresult sr = (d.cuStreamCreateWithPriority && env_u64("VGPU_SWAP_STREAM_PRIO", 1)) ? d.cuStreamCreateWithPriority(s, CU_STREAM_NON_BLOCKING, prio_hi)
                                                   : d.cuStreamCreate(s, CU_STREAM_NON_BLOCKING);
        if (sr != CUDA_SUCCESS) { LOG_ERROR("side stream creation failed"); return false; }
    }
    auto mkring = [&](std::vector<Slot> &ring) {
        ring.resize(cfg_.ring_slots);
        for (auto &s : ring) {
            if (d.cuMemAlloc_v2(&s.buf, cfg_.chunk_bytes) != CUDA_SUCCESS) return false;
            if (d.cuEventCreate(&s.busy, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) return false;
        }
        return true;
    };
    if (!mkring(ring_out_) || !mkring(ring_in_)) { LOG_ERROR("staging ring allocation failed (%d x %zu MiB x 2)", cfg_.ring_slots, cfg_.chunk_bytes >> 20); return false; }

    tbl_cap_ = 4096;
    if (d.cuMemAlloc_v2(&d_tbl_, (size_t)tbl_cap_ * sizeof(VgpuEntry)) != CUDA_SUCCESS) return false;
    if (d.cuMemHostAlloc((void **)&h_tbl_stage_, (size_t)tbl_cap_ * sizeof(VgpuEntry), CU_MEMHOSTALLOC_DEVICEMAP) != CUDA_SUCCESS) return false;
    if (d.cuMemHostGetDevicePointer_v2(&dh_tbl_stage_, h_tbl_stage_, 0) != CUDA_SUCCESS) return false;
    scanner_.reset(new VictimScanner());
    if (scanner_->init(k_, tbl_cap_) != CUDA_SUCCESS) return false;
    own_events_.resize(1024);
    for (auto &e : own_events_) if (d.cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) return false;
    use_ring_.assign(1024, nullptr);
    use_stream_.assign(1024, nullptr);
    use_ctx_.assign(1024, nullptr);
    d.cuCtxGetCurrent(&ctx_);
    pager_ = std::thread([this] { pager_main(); });
    LOG_INFO("swap engine dev %d (numa %d): resident cap %lu MiB, virtual cap %lu MiB, staging %zu MiB x %d, copies %zu MiB, prefetch %lu MiB, arena %lu GiB, gran %zu",
             dev_, numa_node_, (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(cfg_.virtual_cap >> 20), cfg_.chunk_bytes >> 20,
             cfg_.ring_slots, cfg_.copy_bytes >> 20, (unsigned long)(cfg_.prefetch_bytes >> 20), (unsigned long)(cfg_.arena_bytes >> 30), gran_);
    return true;
}

void SwapEngine::stop_pager() {
    if (!pager_.joinable()) return;
    { std::lock_guard<std::mutex> g(mu_); drop_prefetch_queue_locked(); stop_ = true; kick_ = true; }
    cv_pager_.notify_all();
    cv_admit_.notify_all();
    pager_.join();
    { std::lock_guard<std::mutex> g(host_mu_); grow_stop_ = true; }
    host_cv_.notify_all();
    if (grower_.joinable()) grower_.join();
}

SwapEngine::~SwapEngine() {
    const DriverTable &d = drv();
    if (pager_.joinable()) {
        if (d.loaded) drain();
        stop_pager();
    }
    if (!d.loaded) return;
    for (CUstream s : {s_scan_, s_pack_, s_unpack_, s_out_, s_in_}) if (s) d.cuStreamSynchronize(s);
    for (size_t i = 0; i < rows_.size(); i++) {
        if (side_[i].has_handle) { d.cuMemUnmap(rows_[i].base, side_[i].mapped); d.cuMemRelease(side_[i].handle); }
        if (side_[i].has_hh) drop_host_handle(rows_[i].base, side_[i].mapped, side_[i].va_off, side_[i].hhandle, side_[i].hosted);
        if (side_[i].ready) d.cuEventDestroy_v2(side_[i].ready);
        if (side_[i].evict_done) d.cuEventDestroy_v2(side_[i].evict_done);
    }
    for (auto &p : phys_pool_) d.cuMemRelease(p.second);
    for (auto &s : ring_out_) { if (s.buf) d.cuMemFree_v2(s.buf); if (s.busy) d.cuEventDestroy_v2(s.busy); }
    for (auto &s : ring_in_) { if (s.buf) d.cuMemFree_v2(s.buf); if (s.busy) d.cuEventDestroy_v2(s.busy); }
    for (auto &sl : slabs_) if (sl.host) d.cuMemFreeHost(sl.host);
    for (auto &e : own_events_) if (e) d.cuEventDestroy_v2(e);
    for (auto &c : ctx_events_) for (auto &e : c.ev) if (e) d.cuEventDestroy_v2(e);
    for (auto &e : ev_pool_) d.cuEventDestroy_v2(e);
    for (auto &p : prof_) { d.cuEventDestroy_v2(p.a); d.cuEventDestroy_v2(p.b); }
    if (d_span_) d.cuMemFree_v2(d_span_);
    scanner_.reset();
    if (d_tbl_) d.cuMemFree_v2(d_tbl_);
    if (h_tbl_stage_) d.cuMemFreeHost(h_tbl_stage_);
    if (arena_) d.cuMemAddressFree(arena_, cfg_.arena_bytes);
    if (harena_) d.cuMemAddressFree(harena_, cfg_.arena_bytes);
    for (CUstream s : {s_scan_, s_pack_, s_unpack_, s_out_, s_in_}) if (s) d.cuStreamDestroy_v2(s);
}

// ---------------------------------------------------------------------------------------------- small allocators
static bool map_alloc(std::map<uint64_t, uint64_t> &fl, uint64_t bytes, uint64_t *off) {
    for (auto it = fl.begin(); it != fl.end(); ++it) {
        if (it->second >= bytes) {
            *off = it->first;
            uint64_t rest = it->second - bytes, at = it->first + bytes;
            fl.erase(it);
            if (rest) fl[at] = rest;
            return true;
        }
    }
    return false;
}
static void map_free(std::map<uint64_t, uint64_t> &fl, uint64_t off, uint64_t bytes) {
    auto nx = fl.lower_bound(off);
    if (nx != fl.begin()) {
        auto pv = std::prev(nx);
        if (pv->first + pv->second == off) { off = pv->first; bytes += pv->second; fl.erase(pv); }
    }
    if (nx != fl.end() && off + bytes == nx->first) { bytes += nx->second; fl.erase(nx); }
    fl[off] = bytes;
}

bool SwapEngine::va_alloc(size_t bytes, uint64_t *off) { return map_alloc(va_free_, bytes, off); }
void SwapEngine::va_free(uint64_t off, size_t bytes) { map_free(va_free_, off, bytes); }

// Pins one slab (on the GPU's NUMA node) WITHOUT holding host_mu_: a 1 GiB cuMemHostAlloc takes 50-150 ms.
bool SwapEngine::pin_slab(size_t sb, Slab *out, bool *local) {
    const DriverTable &d = drv();
    CUresult r;
    {
        NodeAffinity on_node(numa_node_);
        if (numa_node_ >= 0) prefer_node(numa_node_);
        r = d.cuMemHostAlloc((void **)&out->host, sb, CU_MEMHOSTALLOC_PORTABLE);
        if (numa_node_ >= 0) default_policy();
    }
    if (r != CUDA_SUCCESS) { LOG_ERROR("pinned slab of %zu MiB failed: %d %s", sb >> 20, (int)r, cu_err(r)); return false; }
    *local = false;
    if (numa_node_ >= 0) {
        int got = node_of(out->host + sb / 2);
        *local = got == numa_node_;
        if (!*local) LOG_WARN("pinned slab landed on NUMA node %d, GPU is on node %d: page traffic will cross the socket link", got, numa_node_);
    }
    out->bytes = sb;
    out->free.clear();
    out->free[0] = sb;
    return true;
}

// Every live row will need a pinned block sooner or later (a block stays with its row, so that a clean row is evicted
// without a copy): the pool is grown TOWARDS the live bytes by a helper thread as soon as they are allocated, so that the
// pinning happens while the application populates its buffers and not, slab by slab, inside the pager's evictions.
void SwapEngine::grow_main() {
    drv().cuCtxSetCurrent(ctx_);
    for (;;) {
        size_t sb;
        {
            std::lock_guard<std::mutex> g(host_mu_);
            uint64_t want = host_want_.load();
            if (grow_stop_ || host_total_ >= want) { growing_ = false; host_cv_.notify_all(); return; }
            sb = cfg_.slab_bytes;
            if (cfg_.host_pool_cap && host_total_ + sb > cfg_.host_pool_cap) {
                if (host_total_ >= cfg_.host_pool_cap) { growing_ = false; host_cv_.notify_all(); return; }
                sb = cfg_.host_pool_cap - host_total_;
            }
        }
        Slab s;
        bool local = false;
        bool ok = pin_slab(sb, &s, &local);
        std::lock_guard<std::mutex> g(host_mu_);
        if (!ok) { growing_ = false; host_cv_.notify_all(); return; }
        host_total_ += sb;
        host_slabs_++; if (local) host_slabs_local_++;
        slabs_.push_back(std::move(s));
        host_cv_.notify_all();
    }
}
void SwapEngine::want_host_pool(uint64_t live_bytes) {
    host_want_.store(live_bytes);
    std::lock_guard<std::mutex> g(host_mu_);
    if (growing_ || grow_stop_ || host_total_ >= live_bytes) return;
    if (cfg_.host_pool_cap && host_total_ >= cfg_.host_pool_cap) return;
    if (grower_.joinable()) grower_.join();
    growing_ = true;
    grower_ = std::thread([this] { grow_main(); });
}

// pinned pool: global offset = slab_index << 44 | offset-in-slab. Called by the pager (page-out) and, for releases, by
// application threads (free of a paged-out buffer): own leaf mutex.
bool SwapEngine::host_alloc(size_t bytes, uint64_t *off) {
    bytes = round_up(bytes, 256);
    std::unique_lock<std::mutex> g(host_mu_);
    for (;;) {
        for (size_t i = 0; i < slabs_.size(); i++) {
            uint64_t o;
            if (map_alloc(slabs_[i].free, bytes, &o)) {
                *off = ((uint64_t)i << 44) | o;
                host_used_ += bytes;
                return true;
            }
        }
        if (growing_ && bytes <= cfg_.slab_bytes) { host_cv_.wait(g); continue; }   // a slab is on its way
        break;
    }
    size_t sb = std::max<size_t>(cfg_.slab_bytes, bytes);
    if (cfg_.host_pool_cap && host_total_ + sb > cfg_.host_pool_cap) {
        if (host_total_ + bytes > cfg_.host_pool_cap) return false;
        sb = bytes;
    }
    g.unlock();
    Slab s;
    bool local = false;
    if (!pin_slab(sb, &s, &local)) return false;
    g.lock();
    host_total_ += sb;
    host_slabs_++; if (local) host_slabs_local_++;
    uint64_t o = 0;
    map_alloc(s.free, bytes, &o);
    slabs_.push_back(std::move(s));
    *off = ((uint64_t)(slabs_.size() - 1) << 44) | o;
    host_used_ += bytes;
    return true;
}
unsigned char *SwapEngine::host_ptr(uint64_t off) {
    std::lock_guard<std::mutex> g(host_mu_);
    return slabs_[off >> 44].host + (off & ((1ull << 44) - 1));
}
void SwapEngine::release_host_range(uint64_t off, uint64_t len) {
    std::lock_guard<std::mutex> g(host_mu_);
    len = round_up(len, 256);
    map_free(slabs_[off >> 44].free, off & ((1ull << 44) - 1), len);
    host_used_ -= std::min<uint64_t>(len, host_used_.load());
}
// Pinned pool exhausted: resident rows keep their blocks so that a clean eviction needs no copy — those blocks are the
// first thing to give up (dirty rows' blocks hold stale bytes anyway).
bool SwapEngine::reclaim_host_blocks(Lock &lk, uint64_t bytes) {
    (void)lk;
    uint64_t got = 0;
    for (int pass = 0; pass < 2 && got < bytes; pass++) {
        for (size_t i = 0; i < rows_.size() && got < bytes; i++) {
            Side &s = side_[i];
            if (!(rows_[i].state & VGPU_ST_RESIDENT) || s.phase != PH_IDLE || !s.has_host || s.ready) continue;
            if (pass == 0 && !s.dirty) continue;
            release_host_range(s.host_off, round_up(rows_[i].size, 256));
            s.has_host = false;
            s.dirty = true;          // the HBM copy is the only one now
            got += round_up(rows_[i].size, 256);
        }
    }
    return got >= bytes;
}

// ---------------------------------------------------------------------------------------------- table
void SwapEngine::relock(Lock &lk) { uint64_t t0 = mono_ns(); lk.lock(); pst_.pager_lock_ns += mono_ns() - t0; }
int SwapEngine::new_row() {
    if (!free_rows_.empty()) { int r = free_rows_.back(); free_rows_.pop_back(); return r; }
    rows_.push_back(VgpuEntry{});
    side_.push_back(Side{});
    succ_.push_back(-1);
    return (int)rows_.size() - 1;
}
void SwapEngine::mark_dirty(int row) {
    dirty_lo_ = std::min<uint32_t>(dirty_lo_, (uint32_t)row);
    dirty_hi_ = std::max<uint32_t>(dirty_hi_, (uint32_t)row + 1);
}
// pager thread, mu_ held on entry and exit (released while the scan stream is synchronised)
CUresult SwapEngine::sync_table(Lock &lk, CUstream s) {
    const DriverTable &d = drv();
    uint32_t n = (uint32_t)rows_.size();
    if (n > tbl_cap_) {
        uint32_t nc = tbl_cap_;
        while (nc < n) nc *= 2;
        lk.unlock();
        CUresult r = d.cuStreamSynchronize(s);
        if (r == CUDA_SUCCESS) {
            d.cuMemFree_v2(d_tbl_);
            d.cuMemFreeHost(h_tbl_stage_);
            d_tbl_ = 0; h_tbl_stage_ = nullptr;
            r = d.cuMemAlloc_v2(&d_tbl_, (size_t)nc * sizeof(VgpuEntry));
            if (r == CUDA_SUCCESS) r = d.cuMemHostAlloc((void **)&h_tbl_stage_, (size_t)nc * sizeof(VgpuEntry), CU_MEMHOSTALLOC_DEVICEMAP);
            if (r == CUDA_SUCCESS) r = d.cuMemHostGetDevicePointer_v2(&dh_tbl_stage_, h_tbl_stage_, 0);
            if (r == CUDA_SUCCESS) { scanner_.reset(new VictimScanner()); r = scanner_->init(k_, nc); }
        }
        relock(lk);
        if (r != CUDA_SUCCESS) return r;
        tbl_cap_ = nc;
        dirty_lo_ = 0; dirty_hi_ = (uint32_t)rows_.size();
    }
    if (dirty_lo_ >= dirty_hi_) return CUDA_SUCCESS;
    // the pinned staging copy must not be overwritten while a previous upload is in flight (scans synchronise the stream,
    // so this is immediate in practice)
    lk.unlock();
    CUresult r = d.cuStreamSynchronize(s);
    relock(lk);
    if (r != CUDA_SUCCESS) return r;
    n = (uint32_t)rows_.size();
    if (n > tbl_cap_) return sync_table(lk, s);          // grew meanwhile
    uint32_t lo = dirty_lo_, hi = std::min(dirty_hi_, n);
    if (lo >= hi) return CUDA_SUCCESS;
    std::memcpy(h_tbl_stage_ + lo, rows_.data() + lo, (size_t)(hi - lo) * sizeof(VgpuEntry));
    dirty_lo_ = UINT32_MAX; dirty_hi_ = 0;
    // by kernel, not by a copy engine: a 30 KiB cuMemcpyAsync would queue behind all the page traffic (see vgpu_copy16)
    CU_TRY(launch_copy16(k_, d_tbl_ + (size_t)lo * sizeof(VgpuEntry), dh_tbl_stage_ + (size_t)lo * sizeof(VgpuEntry), (size_t)(hi - lo) * sizeof(VgpuEntry), s));
    pst_.scan_launches++;                                 // the upload is a kernel launch of ours like the scan itself
    return CUDA_SUCCESS;
}

int SwapEngine::lookup(CUdeviceptr p) const {
    if (!owns(p)) return -1;
    std::lock_guard<std::mutex> g(mu_);
    int r = owner_[(p - arena_) / gran_];
    if (r < 0) return -1;
    if (p >= rows_[r].base + side_[r].mapped) return -1;
    return r;
}

bool SwapEngine::range_of(CUdeviceptr p, CUdeviceptr *base, size_t *size) const {
    if (!owns(p)) return false;
    std::lock_guard<std::mutex> g(mu_);
    int r = owner_[(p - arena_) / gran_];
    if (r < 0 || rows_[r].state == VGPU_ST_FREE || p >= rows_[r].base + rows_[r].size) return false;
    if (base) *base = rows_[r].base;
    if (size) *size = rows_[r].size;
    return true;
}

bool SwapEngine::fits_together(const int *rows, int n) const {
    std::lock_guard<std::mutex> g(mu_);
    if (cfg_.host_backed) return true;           // what does not fit is used in place
    uint64_t sum = 0;
    for (int i = 0; i < n; i++) if (row_live(rows[i])) sum += side_[rows[i]].mapped;
    return sum <= (budget_fn_ && sibling_engines_ > 1 ? std::max(cfg_.resident_cap, fair_share_) : cfg_.resident_cap);
}

void SwapEngine::collect_rows(const void *param, size_t bytes, std::vector<int> *rows) const {
    if (bytes < 8) return;
    const unsigned char *b = static_cast<const unsigned char *>(param);
    std::lock_guard<std::mutex> g(mu_);
    for (size_t o = 0; o + 8 <= bytes; o += 8) {
        uint64_t v;
        std::memcpy(&v, b + o, 8);
        if (v < arena_ || v >= arena_ + cfg_.arena_bytes) continue;
        int r = owner_[(v - arena_) / gran_];
        if (r < 0) continue;
        if (std::find(rows->begin(), rows->end(), r) == rows->end()) rows->push_back(r);
    }
}

void SwapEngine::retire_row_locked(int row) {
    Side &s = side_[row];
    va_free(s.va_off, s.mapped);
    uint32_t gen = s.gen + 1;
    CUevent keep_done = s.evict_done;
    if (s.ready) put_event(s.ready);
    s = Side{};
    s.gen = gen;
    if (keep_done) put_event(keep_done);
    rows_[row] = VgpuEntry{};
    succ_[row] = -1;
    mark_dirty(row);
    free_rows_.push_back(row);
}

// ---------------------------------------------------------------------------------------------- events / rings
// the last-use event of ring slot `slot` that belongs to context `cur` (an event is recorded on a stream of its own context)
CUevent SwapEngine::use_slot_event(CUcontext cur, size_t slot) {
    if (cur == ctx_) return own_events_[slot];
    for (CtxEvents &c : ctx_events_) if (c.ctx == cur) return c.ev[slot];
    CtxEvents c;
    c.ctx = cur;
    c.ev.assign(own_events_.size(), nullptr);
    for (auto &e : c.ev) if (drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) { e = nullptr; LOG_ERROR("cuEventCreate in a second context failed"); }
    ctx_events_.push_back(std::move(c));
    return ctx_events_.back().ev[slot];
}
CUevent SwapEngine::use_event(uint64_t seq) {
    if (seq == 0) return nullptr;
    if (seq + use_ring_.size() <= use_seq_) return nullptr;  // slot was recycled: that use is known complete (note_use)
    return use_ring_[seq % use_ring_.size()];
}
// ONE pool of (timing-disabled) events for page-in completions, eviction completions and pack markers, shared by the
// pager and the application threads behind its own leaf mutex: cuEventCreate takes the same driver lock the VMM calls
// stall on, so events are recycled, never created in steady state.
CUevent SwapEngine::get_event() {
    {
        std::lock_guard<std::mutex> g(ev_mu_);
        if (!ev_pool_.empty()) { CUevent e = ev_pool_.back(); ev_pool_.pop_back(); return e; }
    }
    CUevent e = nullptr;
    if (drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) return nullptr;
    return e;
}
void SwapEngine::put_event(CUevent e) { if (e) { std::lock_guard<std::mutex> g(ev_mu_); ev_pool_.push_back(e); } }

SwapEngine::Slot &SwapEngine::acquire_slot(std::vector<Slot> &ring, int *cursor) {
    Slot &s = ring[*cursor];
    *cursor = (*cursor + 1) % (int)ring.size();
    if (s.used) { ScopedNs t(&pst_.pager_ring_ns); drv().cuEventSynchronize(s.busy); }  // back-pressure of the staged path
    s.used = true;
    s.seq++;
    return s;
}
CUdeviceptr SwapEngine::next_span(bool unpack) {
    if (!profile_.load(std::memory_order_relaxed)) return 0;
    const DriverTable &d = drv();
    if (!d_span_) {
        span_cap_ = 1u << 16;
        if (d.cuMemAlloc_v2(&d_span_, (size_t)span_cap_ * 16) != CUDA_SUCCESS) { d_span_ = 0; return 0; }
        std::vector<uint64_t> init((size_t)span_cap_ * 2);
        for (uint32_t i = 0; i < span_cap_; i++) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
        d.cuMemcpyHtoD_v2(d_span_, init.data(), init.size() * 8);
        span_unpack_.assign(span_cap_, 0);
    }
    if (span_next_ >= span_cap_) return 0;   // enough samples
    span_unpack_[span_next_] = unpack;
    return d_span_ + (size_t)(span_next_++) * 16;
}
void SwapEngine::harvest_spans() {
    if (!d_span_ || span_read_ >= span_next_) return;
    const DriverTable &d = drv();
    std::vector<uint64_t> v((size_t)(span_next_ - span_read_) * 2);
    if (d.cuMemcpyDtoH_v2(v.data(), d_span_ + (size_t)span_read_ * 16, v.size() * 8) != CUDA_SUCCESS) return;
    for (uint32_t i = 0; i < span_next_ - span_read_; i++) {
        uint64_t a = v[2 * i], b = v[2 * i + 1];
        if (a == ~0ull || b <= a) continue;
        (span_unpack_[span_read_ + i] ? pst_.unpack_span_ms : pst_.pack_span_ms) += (double)(b - a) / 1e6;
    }
    span_read_ = span_next_;
}
void SwapEngine::prof_begin(CUstream s, CUevent *a) {
    *a = nullptr;
    if (!profile_.load(std::memory_order_relaxed)) return;
    const DriverTable &d = drv();
    if (d.cuEventCreate(a, CU_EVENT_DEFAULT) != CUDA_SUCCESS) { *a = nullptr; return; }
    d.cuEventRecord(*a, s);
}
void SwapEngine::prof_end(CUstream s, CUevent a, bool unpack, uint64_t bytes) {
    if (!a) return;
    const DriverTable &d = drv();
    CUevent b;
    if (d.cuEventCreate(&b, CU_EVENT_DEFAULT) != CUDA_SUCCESS) { d.cuEventDestroy_v2(a); return; }
    d.cuEventRecord(b, s);
    prof_.push_back(Prof{a, b, unpack, bytes});
    if (prof_.size() > 4096) harvest_prof(false);
}
void SwapEngine::harvest_prof(bool wait) {
    const DriverTable &d = drv();
    size_t keep = 0;
    for (size_t i = 0; i < prof_.size(); i++) {
        Prof &p = prof_[i];
        if (wait) d.cuEventSynchronize(p.b);
        if (d.cuEventQuery(p.b) == CUDA_SUCCESS) {
            float ms = 0;
            if (d.cuEventElapsedTime(&ms, p.a, p.b) == CUDA_SUCCESS) {
                if (p.unpack) { pst_.unpack_ms += ms; pst_.unpack_bytes += p.bytes; }
                else { pst_.pack_ms += ms; pst_.pack_bytes += p.bytes; }
            }
            d.cuEventDestroy_v2(p.a);
            d.cuEventDestroy_v2(p.b);
        } else {
            prof_[keep++] = p;
        }
    }
    prof_.resize(keep);
}
void SwapEngine::set_profile(bool on) { profile_.store(on); }

bool SwapEngine::trace_begin(int dir, int row, CUstream s) {
    if (!trace_want_ || trace_.size() >= 2u * trace_want_) return false;
    if (trace_skip_) { trace_skip_--; return false; }
    const DriverTable &d = drv();
    if (!trace_base_) { d.cuEventCreate(&trace_base_, CU_EVENT_DEFAULT); d.cuEventRecord(trace_base_, s); d.cuEventSynchronize(trace_base_); trace_base_ns_ = mono_ns(); trace_.reserve(2u * trace_want_); }
    TraceRec r{dir, row, mono_ns(), nullptr, nullptr};
    d.cuEventCreate(&r.a, CU_EVENT_DEFAULT); d.cuEventCreate(&r.b, CU_EVENT_DEFAULT);
    d.cuEventRecord(r.a, s);
    trace_.push_back(r);
    return true;
}
void SwapEngine::trace_end(CUstream s) { drv().cuEventRecord(trace_.back().b, s); }
void SwapEngine::dump_trace(FILE *f) {
    const DriverTable &d = drv();
    if (trace_.empty()) return;
    for (const TraceRec &r : trace_) {
        float a = 0, b = 0;
        d.cuEventSynchronize(r.b);
        d.cuEventElapsedTime(&a, trace_base_, r.a); d.cuEventElapsedTime(&b, trace_base_, r.b);
        std::fprintf(f, "[vgpu-b200 trace] %s row %d issued %.0f us, ran %.0f..%.0f us (%.0f)\n", r.dir ? "H2D" : "D2H", r.row, (double)(r.host_ns - trace_base_ns_) / 1e3, a * 1e3, b * 1e3, (b - a) * 1e3);
    }
}

void SwapEngine::flush_pager_stats_locked() {
    SwapStats &p = pst_;
    st_.page_out_bytes += p.page_out_bytes; st_.page_in_bytes += p.page_in_bytes; st_.evictions += p.evictions;
    st_.pack_launches += p.pack_launches; st_.unpack_launches += p.unpack_launches; st_.scan_launches += p.scan_launches;
    st_.scans += p.scans; st_.scan_cache_hits += p.scan_cache_hits; st_.phys_creates += p.phys_creates; st_.phys_reuses += p.phys_reuses;
    st_.pack_ms += p.pack_ms; st_.unpack_ms += p.unpack_ms; st_.pack_span_ms += p.pack_span_ms; st_.unpack_span_ms += p.unpack_span_ms;
    st_.pager_vmm_ns += p.pager_vmm_ns; st_.pager_scan_ns += p.pager_scan_ns; st_.pager_packsync_ns += p.pager_packsync_ns;
    st_.pager_ring_ns += p.pager_ring_ns; st_.pager_busy_ns += p.pager_busy_ns; st_.vmm_calls += p.vmm_calls;
    st_.vmm_slow_calls += p.vmm_slow_calls; st_.vmm_slow_ns += p.vmm_slow_ns; st_.vmm_max_ns = std::max(st_.vmm_max_ns, p.vmm_max_ns);
    st_.pager_issue_ns += p.pager_issue_ns; st_.pager_poll_ns += p.pager_poll_ns; st_.pager_lock_ns += p.pager_lock_ns;
    for (int i = 0; i < 5; i++) st_.pager_step_ns[i] += p.pager_step_ns[i];
    st_.pager_unmap_ns += p.pager_unmap_ns; st_.pager_setaccess_ns += p.pager_setaccess_ns; st_.pager_map_ns += p.pager_map_ns; st_.pager_create_ns += p.pager_create_ns;
    st_.pack_bytes += p.pack_bytes; st_.unpack_bytes += p.unpack_bytes; st_.direct_out_bytes += p.direct_out_bytes; st_.direct_in_bytes += p.direct_in_bytes;
    st_.prefetch_issued += p.prefetch_issued; st_.prefetch_wasted += p.prefetch_wasted; st_.clean_evictions += p.clean_evictions;
    p = SwapStats{};
}

// ---------------------------------------------------------------------------------------------- physical memory (pager)
void SwapEngine::pool_phys(size_t mapped, CUmemGenericAllocationHandle h) {
    phys_pool_.emplace(mapped, h);
    phys_pool_bytes_ += mapped;
}
void SwapEngine::trim_phys_pool(uint64_t keep) {
    const DriverTable &d = drv();
    while (!phys_pool_.empty() && phys_pool_bytes_ > keep) {
        auto it = std::prev(phys_pool_.end());
        d.cuMemRelease(it->second);
        phys_pool_bytes_ -= it->first;
        phys_pool_.erase(it);
    }
}
// A handle of exactly `mapped` bytes: recycled from an evicted row of the same size when there is one (steady state:
// no cuMemCreate / cuMemRelease at all), else created — after releasing pooled handles of other sizes when the
// quota would otherwise be exceeded or the device is full.
CUresult SwapEngine::obtain_phys(size_t mapped, CUmemGenericAllocationHandle *h, bool *pressure) {
    const DriverTable &d = drv();
    *pressure = false;
    auto it = phys_pool_.find(mapped);
    if (it != phys_pool_.end()) {
        *h = it->second;
        phys_pool_bytes_ -= mapped;
        phys_pool_.erase(it);
        pst_.phys_reuses++;
        return CUDA_SUCCESS;
    }
    // the caller has already reserved `mapped` in resident_mapped_; the pool may keep what is left under the cap
    uint64_t held, cap;
    { std::lock_guard<std::mutex> g(mu_); held = resident_mapped_ + evicting_mapped_; cap = cfg_.resident_cap; }
    trim_phys_pool(cap > held ? cap - held : 0);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev_;
    CUresult r;
    { ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_create_ns); r = d.cuMemCreate(h, mapped, &prop, 0); }
    if (r == CUDA_ERROR_OUT_OF_MEMORY && !phys_pool_.empty()) {
        trim_phys_pool(0);
        ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_create_ns);
        r = d.cuMemCreate(h, mapped, &prop, 0);
    }
    if (r == CUDA_SUCCESS) pst_.phys_creates++;
    else if (r == CUDA_ERROR_OUT_OF_MEMORY) *pressure = true;
    return r;
}

// Host-backed mode: the row's backing store is a host-located VMM handle, mapped for good at the alias address (what the
// copy engines target). Costs one cuMemCreate + cuMemSetAccess of host memory (~50 us per MiB: the GPU maps system
// memory in small pages) the first time a row is evicted.
CUresult SwapEngine::make_host_handle(size_t mapped, uint64_t va_off, CUmemGenericAllocationHandle *h) {
    const DriverTable &d = drv();
    ScopedNs t(&pst_.pager_vmm_ns);
    CUmemAllocationProp hp = {};
    hp.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    hp.location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA;
    hp.location.id = numa_node_ >= 0 ? numa_node_ : 0;
    CUresult r = d.cuMemCreate(h, mapped, &hp, 0);
    if (r != CUDA_SUCCESS) { LOG_ERROR("host-located cuMemCreate of %zu MiB failed: %d %s", mapped >> 20, (int)r, cu_err(r)); return r; }
    r = d.cuMemMap(harena_ + va_off, mapped, 0, *h, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = dev_;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.cuMemSetAccess(harena_ + va_off, mapped, &acc, 1);
        if (r != CUDA_SUCCESS) d.cuMemUnmap(harena_ + va_off, mapped);
    }
    if (r != CUDA_SUCCESS) { LOG_ERROR("mapping the host backing failed: %d %s", (int)r, cu_err(r)); d.cuMemRelease(*h); }
    return r;
}
void SwapEngine::drop_host_handle(CUdeviceptr base, size_t mapped, uint64_t va_off, CUmemGenericAllocationHandle h, bool hosted) {
    const DriverTable &d = drv();
    if (hosted) d.cuMemUnmap(base, mapped);
    d.cuMemUnmap(harena_ + va_off, mapped);
    d.cuMemRelease(h);
}

void SwapEngine::note_vmm_call(const char *what, uint64_t ns) {
    pst_.vmm_calls++;
    if (ns > pst_.vmm_max_ns) pst_.vmm_max_ns = ns;
    if (ns > 2000000ull) {
        pst_.vmm_slow_calls++; pst_.vmm_slow_ns += ns;
        if (trace_want_) std::fprintf(stderr, "[vgpu-b200 trace] slow %s: %.1f ms at %.0f us\n", what, ns / 1e6, trace_base_ns_ ? (double)(mono_ns() - trace_base_ns_) / 1e3 : 0.0);
    }
}
static void merge_runs(std::vector<std::pair<CUdeviceptr, size_t>> &v) {
    std::sort(v.begin(), v.end());
    size_t k = 0;
    for (size_t i = 0; i < v.size(); i++) {
        if (k && v[k - 1].first + v[k - 1].second == v[i].first) v[k - 1].second += v[i].second;
        else v[k++] = v[i];
    }
    v.resize(k);
}
// One cuMemUnmap per run of adjacent ranges. Every VMM call on B200 / driver 580 waits for the copy in flight
// (profiles/r02_hostvmm_probe.md), so the number of calls matters more than their size.
void SwapEngine::unmap_batch(std::vector<std::pair<CUdeviceptr, size_t>> &ranges) {
    const DriverTable &d = drv();
    if (ranges.empty()) return;
    ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_unmap_ns);
    std::vector<std::pair<CUdeviceptr, size_t>> runs = ranges;
    if (unmap_runs_ok_) {
        merge_runs(runs);
        bool ok = true;
        size_t done = 0;
        for (; done < runs.size(); done++) {
            uint64_t t0 = mono_ns();
            CUresult ur = d.cuMemUnmap(runs[done].first, runs[done].second);
            note_vmm_call("cuMemUnmap", mono_ns() - t0);
            if (ur != CUDA_SUCCESS) { ok = false; break; }
        }
        if (ok) return;
        // this driver wants one call per mapping: finish the rest range by range (a failed call unmapped nothing)
        unmap_runs_ok_ = false;
        CUdeviceptr from = runs[done].first;
        for (auto &r : ranges) if (r.first >= from) { uint64_t t0 = mono_ns(); d.cuMemUnmap(r.first, r.second); note_vmm_call("cuMemUnmap", mono_ns() - t0); }
        return;
    }
    for (auto &r : ranges) { uint64_t t0 = mono_ns(); d.cuMemUnmap(r.first, r.second); note_vmm_call("cuMemUnmap", mono_ns() - t0); }
}
CUresult SwapEngine::set_access_batch(std::vector<std::pair<CUdeviceptr, size_t>> &ranges) {
    const DriverTable &d = drv();
    if (ranges.empty()) return CUDA_SUCCESS;
    ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_setaccess_ns);
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    std::vector<std::pair<CUdeviceptr, size_t>> runs = ranges;
    merge_runs(runs);
    CUresult rc = CUDA_SUCCESS;
    for (auto &r : runs) {
        uint64_t t0 = mono_ns();
        CUresult x = d.cuMemSetAccess(r.first, r.second, &acc, 1);
        note_vmm_call("cuMemSetAccess", mono_ns() - t0);
        if (x != CUDA_SUCCESS && runs.size() != ranges.size()) {
            // fall back to one call per mapping inside this run
            for (auto &q : ranges)
                if (q.first >= r.first && q.first < r.first + r.second) { pst_.vmm_calls++; x = d.cuMemSetAccess(q.first, q.second, &acc, 1); if (x != CUDA_SUCCESS) rc = x; }
        } else if (x != CUDA_SUCCESS) rc = x;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------- predictor (mu_)
// For every row: which row was touched right after it the last time. Training steps, sweeps and inference loops repeat
// their access sequence, so following these links from the row just touched names the next misses. Every touch scores
// the link it arrives by; the pager only prefetches while most recent predictions were right, which keeps random
// access patterns (no repeating order) from generating page traffic.
void SwapEngine::observe_touch(int row) {
    if (row == last_row_) return;
    if (last_row_ >= 0 && (size_t)last_row_ < succ_.size()) {
        int32_t predicted = succ_[last_row_];
        if (predicted >= 0) {
            pred_hist_ = (pred_hist_ << 1) | (predicted == row ? 1u : 0u);
            if (pred_count_ < 32) pred_count_++;
        }
        succ_[last_row_] = row;
    }
    last_row_ = row;
}
bool SwapEngine::predictor_confident() const {
    if (pred_count_ < 8) return false;
    uint32_t n = pred_count_ < 16 ? pred_count_ : 16;
    uint32_t mask = n >= 32 ? ~0u : ((1u << n) - 1u);
    return (uint32_t)__builtin_popcount(pred_hist_ & mask) * 4 >= n * 3;    // >= 75 % of the last n
}
void SwapEngine::hint_prefetch(int row) {
    std::lock_guard<std::mutex> g(mu_);
    if (!row_live(row)) return;
    Side &s = side_[row];
    if (rows_[row].state != VGPU_ST_PAGED_OUT || s.phase != PH_IDLE || s.inplace > 0) return;
    s.phase = PH_QUEUED;
    s.demand = false;
    prefetch_q_.push_back(QEntry{row, s.gen});
    queued_prefetch_bytes_ += s.mapped;
    kick_pager_locked();
}
void SwapEngine::hint_evict(int row) {
    std::lock_guard<std::mutex> g(mu_);
    if (!row_live(row) || rows_[row].state != VGPU_ST_RESIDENT || side_[row].phase != PH_IDLE) return;   // pinned, paged out or on its way: nothing to say
    rows_[row].last_touch = 0;                   // oldest on the LRU clock: the next scan takes it first
    victim_cache_.clear();                       // ... and the next eviction does scan (the cached surplus predates the hint)
    if (side_[row].prefetched) { side_[row].prefetched = false; prefetched_bytes_ -= side_[row].mapped; st_.prefetch_wasted++; }
    mark_dirty(row);
}
void SwapEngine::schedule_prefetch() {
    if (!cfg_.prefetch_bytes || last_row_ < 0 || resident_mapped_ >= live_mapped_ || !predictor_confident()) return;
    uint64_t window = std::min<uint64_t>(cfg_.prefetch_bytes, cfg_.resident_cap / 4);
    uint64_t ahead = queued_prefetch_bytes_ + prefetched_bytes_;
    // top the window up in halves, not row by row: rows wished for together are mapped together (one cuMemSetAccess per
    // run of adjacent ranges) and their victims are unmapped together — VMM calls are where the driver stalls
    if (ahead > window / 2) return;
    int cur = last_row_;
    bool queued = false;
    for (int steps = 0; steps < 64 && ahead < window; steps++) {
        int nxt = (size_t)cur < succ_.size() ? succ_[cur] : -1;
        if (nxt < 0 || nxt == last_row_ || !row_live(nxt)) break;
        cur = nxt;
        Side &s = side_[cur];
        if (rows_[cur].state != VGPU_ST_PAGED_OUT || s.phase != PH_IDLE || s.inplace > 0) continue;   // resident, loading, already queued, or in use where it is
        if (s.mapped > window) break;
        s.phase = PH_QUEUED;
        s.demand = false;
        prefetch_q_.push_back(QEntry{cur, s.gen});
        queued_prefetch_bytes_ += s.mapped;
        ahead += s.mapped;
        queued = true;
    }
    if (queued) kick_pager_locked();
}

// ---------------------------------------------------------------------------------------------- pager: victims
void SwapEngine::collect_waits_locked(int row, std::vector<CUevent> *out) {
    Side &s = side_[row];
    for (int i = 0; i < s.nuses; i++) if (CUevent e = use_event(s.uses[i])) out->push_back(e);
    if (s.ready) out->push_back(s.ready);
}

// Exact-LRU victims for `shortage` bytes: the surplus of the previous GPU scan while it is still valid, else a new scan
// (asking for scan_lookahead times the need). mu_ is released while the scan runs; what it returns is re-validated
// against the table afterwards. *evictable < shortage means even all candidates do not suffice.
CUresult SwapEngine::choose_victims(Lock &lk, uint64_t shortage, std::vector<uint32_t> *victims, uint64_t *evictable) {
    victims->clear();
    *evictable = 0;
    auto valid = [&](const Cand &c) {
        return c.row < rows_.size() && rows_[c.row].state == VGPU_ST_RESIDENT && rows_[c.row].last_touch == c.touch &&
               rows_[c.row].base == c.base && side_[c.row].phase == PH_IDLE;
    };
    uint64_t freed = 0;
    {
        std::deque<Cand> keep = victim_cache_;
        std::vector<uint32_t> got;
        uint64_t f = 0;
        while (f < shortage && !keep.empty()) {
            Cand cnd = keep.front();
            keep.pop_front();
            if (!valid(cnd)) continue;
            got.push_back(cnd.row);
            f += side_[cnd.row].mapped;
        }
        if (f >= shortage) { *victims = got; victim_cache_.swap(keep); pst_.scan_cache_hits++; *evictable = f; return CUDA_SUCCESS; }
        victim_cache_.clear();   // not enough left: start over from a fresh scan (nothing was consumed)
    }
    for (int attempt = 0; attempt < 3; attempt++) {
        ScopedNs t(&pst_.pager_scan_ns);
        CUresult r = sync_table(lk, s_scan_);
        if (r != CUDA_SUCCESS) return r;
        std::vector<uint32_t> found;
        uint64_t found_bytes = 0;
        bool insufficient = false;
        int launches = 0;
        // Look ahead in units of what is being brought in, not of the shortage: the cap is rarely a multiple of the granule
        // (the container's small allocations come out of it byte by byte), so a shortage can be a few bytes — and the
        // next admissions will each need a whole buffer's worth again.
        uint64_t ask = std::max<uint64_t>(shortage, scan_unit_) * (uint64_t)(cfg_.scan_lookahead ? cfg_.scan_lookahead : 1);
        uint32_t n = (uint32_t)rows_.size();
        uint64_t tick = tick_;
        CUdeviceptr tbl = d_tbl_;
        lk.unlock();
        r = scanner_->scan(tbl, n, ask, tick, s_scan_, &found, &found_bytes, &insufficient, &launches);
        relock(lk);
        pst_.scan_launches += launches;
        pst_.scans++;
        if (r != CUDA_SUCCESS) { LOG_ERROR("victim scan failed: %d %s", (int)r, cu_err(r)); return r; }
        // the kernel returns the prefix SET in index order; LRU order within it comes from the host mirror. Rows that
        // changed while the lock was released (touched, pinned, freed) are dropped.
        std::vector<Cand> cands;
        for (uint32_t v : found) {
            if (v >= rows_.size()) continue;
            Cand c{v, rows_[v].last_touch, rows_[v].base};
            if (rows_[v].state == VGPU_ST_RESIDENT && side_[v].phase == PH_IDLE) cands.push_back(c);
        }
        std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.touch != b.touch ? a.touch < b.touch : a.row < b.row; });
        freed = 0;
        victims->clear();
        size_t k = 0;
        while (k < cands.size() && freed < shortage) { victims->push_back(cands[k].row); freed += side_[cands[k].row].mapped; k++; }
        if (freed >= shortage) {
            for (; k < cands.size(); k++) victim_cache_.push_back(cands[k]);
            *evictable = freed;
            return CUDA_SUCCESS;
        }
        *evictable = freed;
        if (insufficient || found.size() == cands.size()) break;    // nothing was lost to races: that is all there is
    }
    return CUDA_SUCCESS;
}

// mu_ held: the victims leave the resident set (the scan must not pick them again, an application thread that touches
// one queues a demand for it) but keep their physical memory until the pager has unmapped them.
void SwapEngine::begin_evict_locked(const std::vector<uint32_t> &victims, std::vector<OutItem> *items) {
    for (uint32_t v : victims) {
        Side &s = side_[v];
        OutItem it;
        it.row = v; it.base = rows_[v].base; it.len = round_up(rows_[v].size, 256); it.mapped = s.mapped;
        it.host_off = s.host_off; it.has_host = s.has_host;
        it.va_off = s.va_off; it.hh = s.hhandle; it.has_hh = s.has_hh;
        it.copy = s.dirty;                        // clean (block still valid) or never written: nothing to copy
        // host-backed mode exists for accesses the hook cannot see — it cannot see their writes either: only a range the
        // application declared read-mostly is trusted to be clean
        if (cfg_.host_backed && !(s.read_mostly && s.has_host && !s.dirty)) it.copy = true;
        collect_waits_locked((int)v, &it.wait);
        if (s.evict_done) it.wait.push_back(s.evict_done);       // an earlier page-out of this row may still be writing its block
        rows_[v].state = VGPU_ST_PAGED_OUT;
        mark_dirty((int)v);
        s.phase = PH_EVICTING;
        resident_mapped_ -= s.mapped;
        evicting_mapped_ += s.mapped;
        if (s.prefetched) { s.prefetched = false; prefetched_bytes_ -= s.mapped; pst_.prefetch_wasted++; }
        items->push_back(std::move(it));
    }
}

// Direct page-out: the copy engine reads the victims' own ranges (no pack kernel, no staging, no extra HBM traffic). Their
// physical memory comes back when the copy is done — the pager evicts ahead of need, so nobody waits for that.
CUresult SwapEngine::evict_direct(Lock &lk, const std::vector<uint32_t> &victims) {
    const DriverTable &d = drv();
    if (victims.empty()) return CUDA_SUCCESS;
    std::vector<OutItem> items;
    begin_evict_locked(victims, &items);
    lk.unlock();
    CUresult rc = CUDA_SUCCESS;
    for (OutItem &it : items) {
        it.done = get_event();
        if (cfg_.host_backed) {
            if (!it.has_hh) {
                // every evicted row gets its host backing, copied or not: its own range will be re-mapped onto it
                if (make_host_handle(it.mapped, it.va_off, &it.hh) != CUDA_SUCCESS) { rc = CUDA_ERROR_OUT_OF_MEMORY; it.failed = true; put_event(it.done); it.done = nullptr; continue; }
                it.has_hh = true;
            }
            if (it.copy) it.has_host = true;
        } else if (it.copy && !it.has_host) {
            if (!host_alloc(it.len, &it.host_off)) {
                relock(lk);
                bool ok = reclaim_host_blocks(lk, it.len);
                lk.unlock();
                if (!ok || !host_alloc(it.len, &it.host_off)) { rc = CUDA_ERROR_OUT_OF_MEMORY; it.failed = true; put_event(it.done); it.done = nullptr; continue; }
            }
            it.has_host = true;
            it.new_host = true;
        }
        CUstream s = it.copy ? s_out_ : s_scan_;
        ScopedNs t_issue(&pst_.pager_issue_ns);
        for (CUevent e : it.wait) d.cuStreamWaitEvent(s, e, 0);
        if (it.copy) {
            unsigned char *hp = cfg_.host_backed ? nullptr : host_ptr(it.host_off);
            bool tr = trace_begin(0, (int)it.row, s);
            for (uint64_t o = 0; o < it.len; o += cfg_.copy_bytes) {
                uint64_t n = std::min<uint64_t>(cfg_.copy_bytes, it.len - o);
                CUresult r = cfg_.host_backed ? d.cuMemcpyDtoDAsync_v2(harena_ + it.va_off + o, it.base + o, n, s)     // the alias range IS host memory
                                              : d.cuMemcpyDtoHAsync_v2(hp + o, it.base + o, n, s);
                if (r != CUDA_SUCCESS) { LOG_ERROR("page-out copy failed: %d %s", (int)r, cu_err(r)); rc = r; it.failed = true; break; }
            }
            if (tr) trace_end(s);
            if (it.failed) {
                // the block would hold part of the buffer: the victim stays resident (and dirty), a block taken for it goes back
                if (it.new_host) { release_host_range(it.host_off, it.len); it.has_host = false; }
                put_event(it.done); it.done = nullptr;
                continue;
            }
            pst_.page_out_bytes += it.len;
            pst_.direct_out_bytes += it.len;
        } else {
            pst_.clean_evictions++;
        }
        d.cuEventRecord(it.done, s);
    }
    relock(lk);
    for (OutItem &it : items) {
        Side &s = side_[it.row];
        if (it.failed) {
            // no block for it (the pinned pool is exhausted) or the copy could not be issued: this victim stays resident
            rows_[it.row].state = VGPU_ST_RESIDENT;
            mark_dirty((int)it.row);
            s.phase = PH_IDLE;
            evicting_mapped_ -= s.mapped;
            resident_mapped_ += s.mapped;
            if (cfg_.host_backed) { s.hhandle = it.hh; s.has_hh = it.has_hh; }      // a backing made for it is kept for the next time
            LOG_ERROR("cannot page out row %u (pinned host pool: %lu MiB in use)", it.row, (unsigned long)(host_used_.load() >> 20));
            continue;
        }
        s.host_off = it.host_off;
        s.has_host = it.has_host;
        s.hhandle = it.hh; s.has_hh = it.has_hh;
        if (it.copy) s.dirty = false;              // once the copy is done the block equals the HBM content
        s.nuses = 0;                               // the page-out is ordered behind every outstanding use (it.wait)
        if (s.ready) { put_event(s.ready); s.ready = nullptr; }   // its waiters are enqueued; the record they refer to is fixed
        if (s.evict_done) put_event(s.evict_done);
        s.evict_done = it.done;
        s.out_slot = -1;
        rows_[it.row].host_slot = (uint32_t)(s.host_off >> 12);
        evicting_.push_back(it.row);
    }
    flush_pager_stats_locked();
    return rc;
}

// ---------------------------------------------------------------------------------------------- pager: page-in
void SwapEngine::begin_load_locked(int row, bool prefetch, InItem *it) {
    Side &s = side_[row];
    it->row = row; it->base = rows_[row].base; it->len = round_up(rows_[row].size, 256); it->mapped = s.mapped;
    it->host_off = s.host_off; it->has_host = s.has_host; it->prefetch = prefetch;
    it->va_off = s.va_off; it->hosted = s.hosted;
    it->after = nullptr;
    it->wait.clear();
    if (s.hosted) collect_waits_locked(row, &it->wait);      // uses of the row where it was (in-place operands of oversized launches)
    s.phase = PH_LOADING;
    resident_mapped_ += s.mapped;
    it->ready = get_event();
}
void SwapEngine::commit_load_locked(InItem &it) {
    Side &s = side_[it.row];
    s.handle = it.h;
    s.has_handle = true;
    s.hosted = false;
    s.phase = PH_IDLE;
    s.fail = CUDA_SUCCESS;
    s.dirty = false;
    s.demand = false;
    if (s.ready) put_event(s.ready);
    s.ready = it.ready;
    rows_[it.row].state = VGPU_ST_RESIDENT | ((s.pins > 0 || s.locked) ? VGPU_ST_PINNED : 0u);
    if (it.prefetch && s.pins == 0) {
        // newest on the LRU clock, like a touch: the exact-LRU scan must not take what is about to be used
        rows_[it.row].last_touch = tick_;
        s.prefetched = true;
        prefetched_bytes_ += s.mapped;
        pst_.prefetch_issued++;
    }
    mark_dirty(it.row);
}
void SwapEngine::fail_load_locked(InItem &it, CUresult rc) {
    Side &s = side_[it.row];
    resident_mapped_ -= s.mapped;
    if (it.ready) put_event(it.ready);
    it.ready = nullptr;
    s.hosted = it.hosted;                        // false when its host mapping was already taken down for this attempt
    s.phase = PH_IDLE;
    s.fail = rc;
    s.demand = false;
}

// The device cannot give what the quota promises: on an overcommitted GPU other containers hold the rest (the reference
// leaves this case to UVM, which pages between processes). Live within what we have: lower the working cap to what is
// held plus what the device still has, and queue the row again — the pager then makes the room out of our own LRU rows.
bool SwapEngine::requeue_under_pressure_locked(int row, bool was_demand, size_t free_dev) {
    Side &s = side_[row];
    uint64_t held = resident_mapped_ + evicting_mapped_ + (uint64_t)free_dev / gran_ * gran_;
    if (held < s.mapped || ++s.retries > 8) return false;          // nothing of ours left to give up
    if (!pressure_ || cfg_.resident_cap > held) {
        if (!pressure_) LOG_WARN("device %d: physical memory exhausted below the quota; resident cap %lu -> %lu MiB", dev_,
                                 (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(held >> 20));
        else LOG_INFO("device %d: upward probe failed; resident cap back to %lu MiB", dev_, (unsigned long)(held >> 20));
        cfg_.resident_cap = std::min(cfg_.resident_cap, held);
        pressure_ = true;
    }
    pressure_events_++;
    s.fail = CUDA_SUCCESS;
    s.phase = PH_QUEUED;
    s.demand = was_demand;
    if (was_demand) demand_q_.push_front(QEntry{row, s.gen});
    else { prefetch_q_.push_front(QEntry{row, s.gen}); queued_prefetch_bytes_ += s.mapped; }
    return true;
}

// Direct page-in of a batch: map every row (cuMemMap is cheap), ONE cuMemSetAccess per run of adjacent ranges, then the
// copy engine writes each row's own range from its pinned block and the row's `ready` event is recorded behind it.
CUresult SwapEngine::load_direct(Lock &lk, std::vector<InItem> &items) {
    const DriverTable &d = drv();
    if (items.empty()) return CUDA_SUCCESS;
    lk.unlock();
    std::vector<std::pair<CUdeviceptr, size_t>> ranges;
    bool pressure_seen = false;
    for (InItem &it : items) {
        bool pressure = false;
        it.rc = obtain_phys(it.mapped, &it.h, &pressure);
        pressure_seen |= pressure;
        if (it.rc != CUDA_SUCCESS) continue;
        if (it.hosted) {                             // host-backed mode: the range maps the host backing right now
            for (CUevent e : it.wait) d.cuEventSynchronize(e);    // ... and work that was told to use it there must be through with it
            std::vector<std::pair<CUdeviceptr, size_t>> one{{it.base, it.mapped}};
            unmap_batch(one);
            it.hosted = false;
        }
        {
            ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_map_ns);
            it.rc = d.cuMemMap(it.base, it.mapped, 0, it.h, 0);
        }
        if (it.rc != CUDA_SUCCESS) { LOG_ERROR("cuMemMap failed: %d %s", (int)it.rc, cu_err(it.rc)); pool_phys(it.mapped, it.h); continue; }
        ranges.emplace_back(it.base, it.mapped);
    }
    CUresult arc = set_access_batch(ranges);
    for (InItem &it : items) {
        if (it.rc != CUDA_SUCCESS) continue;
        if (arc != CUDA_SUCCESS) {
            LOG_ERROR("cuMemSetAccess failed: %d %s", (int)arc, cu_err(arc));
            d.cuMemUnmap(it.base, it.mapped); pool_phys(it.mapped, it.h); it.rc = arc;
            continue;
        }
        if (!it.has_host) { if (it.ready) { put_event(it.ready); it.ready = nullptr; } continue; }   // never written: nothing to load
        if (!it.ready) it.ready = get_event();
        ScopedNs t_issue(&pst_.pager_issue_ns);
        if (it.after) d.cuStreamWaitEvent(s_in_, it.after, 0);
        unsigned char *hp = cfg_.host_backed ? nullptr : host_ptr(it.host_off);
        bool tr = trace_begin(1, it.row, s_in_);
        for (uint64_t o = 0; o < it.len; o += cfg_.copy_bytes) {
            uint64_t n = std::min<uint64_t>(cfg_.copy_bytes, it.len - o);
            CUresult r = cfg_.host_backed ? d.cuMemcpyDtoDAsync_v2(it.base + o, harena_ + it.va_off + o, n, s_in_)
                                          : d.cuMemcpyHtoDAsync_v2(it.base + o, hp + o, n, s_in_);
            if (r != CUDA_SUCCESS) { LOG_ERROR("page-in copy failed: %d %s", (int)r, cu_err(r)); it.rc = r; break; }
        }
        if (tr) trace_end(s_in_);
        if (it.rc != CUDA_SUCCESS) {
            // part of the row may be on its way in: let the stream run dry before the range goes away again
            d.cuStreamSynchronize(s_in_);
            std::vector<std::pair<CUdeviceptr, size_t>> one{{it.base, it.mapped}};
            unmap_batch(one);
            pool_phys(it.mapped, it.h);
            continue;
        }
        d.cuEventRecord(it.ready, s_in_);
        pst_.page_in_bytes += it.len;
        pst_.direct_in_bytes += it.len;
    }
    size_t fr = 0, tot = 0;
    if (pressure_seen) d.cuMemGetInfo_v2(&fr, &tot);
    relock(lk);
    CUresult rc = CUDA_SUCCESS;
    for (InItem &it : items) {
        if (it.rc == CUDA_SUCCESS) { side_[it.row].retries = 0; commit_load_locked(it); continue; }
        Side &s = side_[it.row];
        bool was_demand = s.demand;
        fail_load_locked(it, it.rc);
        if (it.rc == CUDA_ERROR_OUT_OF_MEMORY && requeue_under_pressure_locked(it.row, was_demand, fr)) continue;
        rc = it.rc;
    }
    flush_pager_stats_locked();
    publish_locked();
    cv_admit_.notify_all();
    return rc;
}

// ---------------------------------------------------------------------------------------------- pager: staged path
// Pack the victims' dirty bytes into staging slots with the TMA kernel and drain every slot to the victims' pinned
// blocks. *packed fires when the LAST PACK has read the victims: from then on their physical memory may be recycled,
// one PCIe transfer earlier than with a direct copy.
CUresult SwapEngine::staged_out(std::vector<OutItem> &items, CUevent *packed) {
    const DriverTable &d = drv();
    *packed = nullptr;
    for (OutItem &it : items) {
        if (!it.copy) { pst_.clean_evictions++; for (CUevent e : it.wait) d.cuStreamWaitEvent(s_pack_, e, 0); continue; }
        if (!it.has_host) {
            if (!host_alloc(it.len, &it.host_off)) { LOG_ERROR("pinned host pool exhausted (%lu MiB in use)", (unsigned long)(host_used_.load() >> 20)); return CUDA_ERROR_OUT_OF_MEMORY; }
            it.has_host = true;
        }
        // order the pack behind the victim's last users and its own page-in
        for (CUevent e : it.wait) d.cuStreamWaitEvent(s_pack_, e, 0);
    }
    struct OutRun { unsigned char *dst; uint64_t pos, len; };
    std::vector<PackSegment> segs;
    std::vector<OutRun> runs;
    Slot *slot = nullptr;
    uint64_t pos = 0;
    auto flush = [&]() -> CUresult {
        if (!slot || segs.empty()) return CUDA_SUCCESS;
        int launches = 0;
        CUevent pa;
        prof_begin(s_pack_, &pa);
        CUresult r = launch_pack(k_, segs.data(), segs.size(), s_pack_, &launches, next_span(false));
        if (r != CUDA_SUCCESS) return r;
        prof_end(s_pack_, pa, false, pos);
        pst_.pack_launches += launches;
        // slot.busy doubles as "packed" marker for the copy stream, then is re-recorded as "drained"
        CU_TRY(d.cuEventRecord(slot->busy, s_pack_));
        CU_TRY(d.cuStreamWaitEvent(s_out_, slot->busy, 0));
        for (const OutRun &rn : runs) CU_TRY(d.cuMemcpyDtoHAsync_v2(rn.dst, slot->buf + rn.pos, rn.len, s_out_));
        CU_TRY(d.cuEventRecord(slot->busy, s_out_));
        pst_.page_out_bytes += pos;
        segs.clear(); runs.clear();
        slot = nullptr;
        return CUDA_SUCCESS;
    };
    for (OutItem &it : items) {
        if (!it.copy) continue;
        unsigned char *hp = host_ptr(it.host_off);
        uint64_t off = 0;
        while (off < it.len) {
            if (!slot) { slot = &acquire_slot(ring_out_, &cur_out_); pos = 0; }
            uint64_t piece = std::min<uint64_t>(it.len - off, cfg_.chunk_bytes - pos);
            segs.push_back(PackSegment{it.base + off, slot->buf + pos, piece});
            if (!runs.empty() && runs.back().dst + runs.back().len == hp + off && runs.back().pos + runs.back().len == pos) runs.back().len += piece;
            else runs.push_back(OutRun{hp + off, pos, piece});
            // the staging slot that carries this row's tail: a page-in of the same row orders its H2D behind THAT slot's
            // drain only (not behind the whole page-out queue, which would serialise the two link directions)
            it.out_slot = (int)(slot - ring_out_.data());
            it.out_seq = slot->seq;
            off += piece; pos += piece;
            if (pos == cfg_.chunk_bytes || segs.size() == VGPU_PACK_MAX_SEG) { CUresult r = flush(); if (r != CUDA_SUCCESS) return r; }
        }
    }
    { CUresult r = flush(); if (r != CUDA_SUCCESS) return r; }
    *packed = get_event();
    if (!*packed) return CUDA_ERROR_OUT_OF_MEMORY;
    CU_TRY(d.cuEventRecord(*packed, s_pack_));
    return CUDA_SUCCESS;
}

void SwapEngine::plan_staged_in(const InItem &it, std::vector<InJob> *jobs) {
    jobs->clear();
    if (!it.has_host) return;
    unsigned char *hp = host_ptr(it.host_off);
    InJob cur;
    uint64_t pos = 0, off = 0;
    while (off < it.len) {
        uint64_t piece = std::min<uint64_t>(it.len - off, cfg_.chunk_bytes - pos);
        if (!cur.runs.empty() && cur.runs.back().src + cur.runs.back().len == hp + off && cur.runs.back().pos + cur.runs.back().len == pos) cur.runs.back().len += piece;
        else cur.runs.push_back(InRun{hp + off, pos, piece});
        cur.segs.push_back(PackSegment{pos, it.base + off, piece});   // src = offset inside the slot, fixed up at issue time
        off += piece; pos += piece;
        if (pos == cfg_.chunk_bytes || cur.segs.size() == VGPU_PACK_MAX_SEG) { cur.bytes = pos; jobs->push_back(std::move(cur)); cur = InJob(); pos = 0; }
    }
    if (!cur.segs.empty()) { cur.bytes = pos; jobs->push_back(std::move(cur)); }
}

CUresult SwapEngine::in_issue_copies(InJob &j) {
    const DriverTable &d = drv();
    for (const InRun &r : j.runs) {
        CU_TRY(d.cuMemcpyHtoDAsync_v2(j.slot->buf + r.pos, r.src, r.len, s_in_));
        pst_.page_in_bytes += r.len;
    }
    CU_TRY(d.cuEventRecord(j.slot->busy, s_in_));     // "loaded"; re-recorded as "unpacked" after the unpack launch
    return CUDA_SUCCESS;
}

// The latency path of a demand miss that finds no free physical memory and no eviction in flight:
//   packs of the victims (short) -> H2D of the incoming row into staging (needs no physical memory yet) -> wait for the
//   last pack, unmap the victims -> map the incoming row with a recycled handle -> unpack behind the H2D.
// The incoming transfer starts before the victims' memory is free, and that memory is free after an HBM-speed pack
// instead of after a PCIe transfer.
CUresult SwapEngine::swap_staged(Lock &lk, int row, const std::vector<uint32_t> &victims) {
    const DriverTable &d = drv();
    std::vector<OutItem> outs;
    begin_evict_locked(victims, &outs);
    InItem in;
    {
        Side &s = side_[row];
        in.row = row; in.base = rows_[row].base; in.len = round_up(rows_[row].size, 256); in.mapped = s.mapped;
        in.host_off = s.host_off; in.has_host = s.has_host; in.prefetch = false;
        s.phase = PH_LOADING;                      // owned by the pager from here on (accounted once the victims are gone)
        in.ready = get_event();
    }
    int in_out_slot = side_[row].out_slot;
    uint64_t in_out_seq = side_[row].out_seq;
    lk.unlock();

    CUevent packed = nullptr;
    CUresult rc = staged_out(outs, &packed);
    std::vector<InJob> jobs;
    size_t staged = 0;
    if (rc == CUDA_SUCCESS) {
        // a row that was paged out moments ago: its D2H may still be in flight -> order the H2D behind that chunk only
        if (in_out_slot >= 0 && ring_out_[in_out_slot].seq == in_out_seq) d.cuStreamWaitEvent(s_in_, ring_out_[in_out_slot].busy, 0);
        plan_staged_in(in, &jobs);
        for (InJob &j : jobs) {
            // A slot staged here holds data until its unpack has been launched below; its `busy` event meanwhile only says
            // "loaded". So never wrap around onto a slot this admission has already filled.
            if (staged == ring_in_.size()) break;
            Slot &s = ring_in_[cur_in_];
            if (s.used && d.cuEventQuery(s.busy) != CUDA_SUCCESS) break;   // never block here
            cur_in_ = (cur_in_ + 1) % (int)ring_in_.size();
            s.used = true;
            s.seq++;
            j.slot = &s;
            staged++;
            if ((rc = in_issue_copies(j)) != CUDA_SUCCESS) break;
        }
    }
    // victims: wait for the last pack, then their ranges can go
    std::vector<std::pair<CUdeviceptr, size_t>> ranges;
    if (packed) { ScopedNs t(&pst_.pager_packsync_ns); d.cuEventSynchronize(packed); put_event(packed); }
    else if (rc != CUDA_SUCCESS) d.cuStreamSynchronize(s_pack_);
    bool evicted = rc == CUDA_SUCCESS;
    if (evicted) {
        for (OutItem &it : outs) ranges.emplace_back(it.base, it.mapped);
        unmap_batch(ranges);
    }
    relock(lk);
    for (OutItem &it : outs) {
        Side &s = side_[it.row];
        if (!evicted) {                              // nothing was unmapped: the victims stay where they are
            s.host_off = it.host_off; s.has_host = it.has_host;   // a block taken for it stays its block
            rows_[it.row].state = VGPU_ST_RESIDENT;
            s.phase = PH_IDLE;
            evicting_mapped_ -= s.mapped;
            resident_mapped_ += s.mapped;
            mark_dirty((int)it.row);
            continue;
        }
        pool_phys(s.mapped, s.handle);
        s.has_handle = false;
        s.nuses = 0;
        s.host_off = it.host_off; s.has_host = it.has_host;
        if (it.copy) { s.dirty = false; s.out_slot = it.out_slot; s.out_seq = it.out_seq; }
        if (s.ready) { put_event(s.ready); s.ready = nullptr; }
        rows_[it.row].host_slot = (uint32_t)(s.host_off >> 12);
        evicting_mapped_ -= s.mapped;
        pst_.evictions++;
        if (s.demand) { s.phase = PH_QUEUED; demand_q_.push_back(QEntry{(int)it.row, s.gen}); }
        else s.phase = PH_IDLE;
        mark_dirty((int)it.row);
    }
    if (rc != CUDA_SUCCESS) {
        side_[row].phase = PH_IDLE;
        side_[row].fail = rc;
        side_[row].demand = false;
        if (in.ready) put_event(in.ready);
        flush_pager_stats_locked();
        cv_admit_.notify_all();
        return rc;
    }
    resident_mapped_ += in.mapped;
    lk.unlock();

    bool pressure = false;
    rc = obtain_phys(in.mapped, &in.h, &pressure);
    if (rc == CUDA_SUCCESS) {
        { ScopedNs t(&pst_.pager_vmm_ns), t2(&pst_.pager_map_ns); rc = d.cuMemMap(in.base, in.mapped, 0, in.h, 0); }
        if (rc != CUDA_SUCCESS) pool_phys(in.mapped, in.h);
    }
    if (rc == CUDA_SUCCESS) {
        std::vector<std::pair<CUdeviceptr, size_t>> r1{{in.base, in.mapped}};
        rc = set_access_batch(r1);
        if (rc != CUDA_SUCCESS) { d.cuMemUnmap(in.base, in.mapped); pool_phys(in.mapped, in.h); }
    }
    if (rc == CUDA_SUCCESS && !jobs.empty()) {
        if (!in.ready) in.ready = get_event();
        for (InJob &j : jobs) {
            if (!j.slot) {
                j.slot = &acquire_slot(ring_in_, &cur_in_);
                if ((rc = in_issue_copies(j)) != CUDA_SUCCESS) break;
            }
            for (PackSegment &sg : j.segs) sg.src += j.slot->buf;
            d.cuStreamWaitEvent(s_unpack_, j.slot->busy, 0);
            int launches = 0;
            CUevent pa;
            prof_begin(s_unpack_, &pa);
            rc = launch_pack(k_, j.segs.data(), j.segs.size(), s_unpack_, &launches, next_span(true));
            if (rc != CUDA_SUCCESS) break;
            prof_end(s_unpack_, pa, true, j.bytes);
            pst_.unpack_launches += launches;
            d.cuEventRecord(j.slot->busy, s_unpack_);
        }
        if (rc == CUDA_SUCCESS) d.cuEventRecord(in.ready, s_unpack_);
    } else if (rc == CUDA_SUCCESS && in.ready) { put_event(in.ready); in.ready = nullptr; }
    size_t fr = 0, tot = 0;
    if (pressure) d.cuMemGetInfo_v2(&fr, &tot);
    relock(lk);
    in.rc = rc;
    if (rc == CUDA_SUCCESS) { side_[row].retries = 0; commit_load_locked(in); }
    else {
        // staged copies that were issued for nothing only wrote staging slots; the row stays paged out
        fail_load_locked(in, rc);
        if (rc == CUDA_ERROR_OUT_OF_MEMORY && requeue_under_pressure_locked(row, true, fr)) rc = CUDA_SUCCESS;
        else LOG_ERROR("page-in of row %d failed: %d %s", row, (int)rc, cu_err(rc));
    }
    flush_pager_stats_locked();
    publish_locked();
    cv_admit_.notify_all();
    return rc;
}

// Shared budget (several processes, one quota): refresh this engine's cap from the container's region and reserve the
// growth there. Without a budget function the cap is local and the growth is granted when it fits.
bool SwapEngine::reserve_locked(uint64_t extra) {
    if (!budget_fn_) return (int64_t)extra <= free_phys_locked();
    bool granted = false;
    int engines = 1;
    uint64_t share = 0;
    uint64_t cap = budget_fn_(resident_mapped_ + evicting_mapped_ + extra, live_bytes_.load(), &granted, &engines, &share);
    fair_share_ = share;
    budget_checked_ns_ = mono_ns();
    sibling_engines_ = engines;
    if (cap != ~0ull) {
        quota_cap_ = cap;
        cfg_.resident_cap = pressure_ ? std::min(cap, cfg_.resident_cap) : cap;
    }
    return granted && (int64_t)extra <= free_phys_locked();
}

// ---------------------------------------------------------------------------------------------- pager: main loop
bool SwapEngine::step_zombies(Lock &lk) {
    const DriverTable &d = drv();
    if (zombies_.empty()) return false;
    std::vector<uint32_t> ready_rows;
    for (auto it = zombies_.begin(); it != zombies_.end();) {
        Side &s = side_[*it];
        bool busy = false;
        for (int i = 0; i < s.nuses && !busy; i++) if (CUevent e = use_event(s.uses[i])) busy = d.cuEventQuery(e) != CUDA_SUCCESS;
        if (!busy && s.ready) busy = d.cuEventQuery(s.ready) != CUDA_SUCCESS;
        if (!busy && s.evict_done) busy = d.cuEventQuery(s.evict_done) != CUDA_SUCCESS;
        if (busy) { ++it; continue; }
        ready_rows.push_back(*it);
        it = zombies_.erase(it);
    }
    if (ready_rows.empty()) return false;
    std::vector<std::pair<CUdeviceptr, size_t>> ranges;
    struct DropHost { CUdeviceptr base; size_t mapped; uint64_t va_off; CUmemGenericAllocationHandle hh; bool hosted; };
    std::vector<DropHost> drops;
    for (uint32_t r : ready_rows) {
        if (side_[r].has_handle) ranges.emplace_back(rows_[r].base, side_[r].mapped);
        if (side_[r].has_hh) drops.push_back(DropHost{rows_[r].base, side_[r].mapped, side_[r].va_off, side_[r].hhandle, side_[r].hosted});
    }
    lk.unlock();
    unmap_batch(ranges);
    for (DropHost &x : drops) drop_host_handle(x.base, x.mapped, x.va_off, x.hh, x.hosted);
    relock(lk);
    for (uint32_t r : ready_rows) {
        Side &s = side_[r];
        s.has_hh = false; s.hosted = false;
        if (!s.has_handle) { if (s.has_host && !cfg_.host_backed) release_host_range(s.host_off, s.zombie_len); s.has_host = false; retire_row_locked((int)r); continue; }
        pool_phys(s.mapped, s.handle);
        s.has_handle = false;
        resident_mapped_ -= s.mapped;
        if (s.has_host && !cfg_.host_backed) release_host_range(s.host_off, s.zombie_len);
        s.has_host = false;
        retire_row_locked((int)r);
    }
    flush_pager_stats_locked();
    publish_locked();
    cv_admit_.notify_all();
    return true;
}

// Evictions whose copy is done: unmap (one call per run of adjacent ranges), recycle the handles.
bool SwapEngine::step_reap(Lock &lk) {
    const DriverTable &d = drv();
    if (evicting_.empty()) return false;
    std::vector<uint32_t> done;
    uint64_t t_poll = mono_ns();
    for (auto it = evicting_.begin(); it != evicting_.end();) {
        Side &s = side_[*it];
        if (s.phase != PH_EVICTING) { it = evicting_.erase(it); continue; }
        if (s.evict_done && d.cuEventQuery(s.evict_done) != CUDA_SUCCESS) { ++it; continue; }
        done.push_back(*it);
        it = evicting_.erase(it);
        if (done.size() >= 4u * cfg_.batch_rows) break;
    }
    pst_.pager_poll_ns += mono_ns() - t_poll;
    if (done.empty()) return false;
    // Unmap in batches (adjacent victims go in one call) unless somebody is waiting for the memory: the evictions were
    // issued ahead of need, so their frames are not urgent, and fewer VMM calls means fewer chances to stall in the driver.
    bool urgent = !demand_q_.empty() || (!prefetch_q_.empty() && free_phys_locked() <= 0) || stop_;
    if (!urgent && done.size() < cfg_.batch_rows && !evicting_.empty() && reap_defer_ < 64) {
        reap_defer_++;
        for (auto it = done.rbegin(); it != done.rend(); ++it) evicting_.push_front(*it);
        return false;
    }
    reap_defer_ = 0;
    std::vector<std::pair<CUdeviceptr, size_t>> ranges;
    struct Rehost { CUdeviceptr base; size_t mapped; CUmemGenericAllocationHandle hh; bool ok; };
    std::vector<Rehost> rehost;
    for (uint32_t r : done) {
        ranges.emplace_back(rows_[r].base, side_[r].mapped);
        if (cfg_.host_backed && side_[r].has_hh) rehost.push_back(Rehost{rows_[r].base, side_[r].mapped, side_[r].hhandle, false});
    }
    lk.unlock();
    unmap_batch(ranges);
    if (!rehost.empty()) {
        // host-backed mode: the range does not stay a hole — it is mapped onto the row's host backing, so that an access the
        // hook never saw is served over PCIe instead of killing the context (between the two calls it IS a hole: ~1 ms)
        ScopedNs t(&pst_.pager_vmm_ns);
        std::vector<std::pair<CUdeviceptr, size_t>> hr;
        for (Rehost &x : rehost) {
            x.ok = d.cuMemMap(x.base, x.mapped, 0, x.hh, 0) == CUDA_SUCCESS;
            if (x.ok) hr.emplace_back(x.base, x.mapped);
        }
        if (set_access_batch(hr) != CUDA_SUCCESS) for (Rehost &x : rehost) if (x.ok) { d.cuMemUnmap(x.base, x.mapped); x.ok = false; }
    }
    relock(lk);
    size_t rh = 0;
    for (uint32_t r : done) {
        Side &s = side_[r];
        if (cfg_.host_backed && s.has_hh) s.hosted = rehost[rh++].ok;
        pool_phys(s.mapped, s.handle);
        s.has_handle = false;
        evicting_mapped_ -= s.mapped;
        pst_.evictions++;
        put_event(s.evict_done);
        s.evict_done = nullptr;
        if (s.demand) { s.phase = PH_QUEUED; demand_q_.push_front(QEntry{(int)r, s.gen}); }   // touched while it was on its way out
        else s.phase = PH_IDLE;
    }
    flush_pager_stats_locked();
    publish_locked();
    cv_admit_.notify_all();
    return true;
}

bool SwapEngine::step_demand(Lock &lk) {
    while (!demand_q_.empty()) {
        QEntry e = demand_q_.front();
        if ((size_t)e.row < side_.size() && side_[e.row].gen == e.gen && side_[e.row].phase == PH_QUEUED && side_[e.row].demand) break;
        demand_q_.pop_front();
    }
    if (demand_q_.empty()) return false;
    if (side_[demand_q_.front().row].inplace > 0) {
        // somebody is using this row where it is (host-mapped, an oversized launch): it cannot move before that use is over
        QEntry e = demand_q_.front();
        demand_q_.pop_front();
        demand_q_.push_back(e);
        return false;
    }
    const int row = demand_q_.front().row;
    const uint64_t need = side_[row].mapped;
    scan_unit_ = need;
    if (pressure_ && cfg_.resident_cap < quota_cap_ && (++pressure_probe_ & 31u) == 0) {
        // probe upwards: the other tenants may have let go; a failed cuMemCreate simply lowers the cap again
        cfg_.resident_cap = std::min(quota_cap_, cfg_.resident_cap + need);
        if (cfg_.resident_cap == quota_cap_) pressure_ = false;
    }
    auto fail = [&](CUresult rc) {
        demand_row_ = -1;
        demand_q_.pop_front();
        side_[row].phase = PH_IDLE;
        side_[row].fail = rc;
        side_[row].demand = false;
        cv_admit_.notify_all();
        return true;
    };
    // the siblings may have grown or let go: refresh the cap (a semaphore round trip on the container's region: at most
    // every 200 us — this step runs every 40 us while a demand waits)
    if (budget_fn_ && mono_ns() - budget_checked_ns_ > 200000ull) reserve_locked(0);
    if (need > cfg_.resident_cap && sibling_engines_ <= 1) return fail(CUDA_ERROR_OUT_OF_MEMORY);
    int64_t free_now = free_phys_locked();
    if (free_now >= (int64_t)need && reserve_locked(need)) {
        // room is there (freed by the pager ahead of need, or never used): map + direct copy, together with whatever
        // other demanded rows fit
        std::vector<InItem> items;
        demand_row_ = -1;                                // (the waiting clocks below restart with the next demand)
        while (!demand_q_.empty() && items.size() < cfg_.batch_rows) {
            QEntry e = demand_q_.front();
            Side &s = side_[e.row];
            bool ok = s.gen == e.gen && s.phase == PH_QUEUED && s.demand;
            if (ok && s.inplace > 0) break;
            if (ok && (free_phys_locked() < (int64_t)s.mapped || (budget_fn_ && !items.empty() && !reserve_locked(s.mapped)))) break;
            demand_q_.pop_front();
            if (!ok) continue;
            items.emplace_back();
            begin_load_locked(e.row, false, &items.back());
            Side &sr = side_[e.row];
            if (sr.out_slot >= 0 && ring_out_[sr.out_slot].seq == sr.out_seq) items.back().after = ring_out_[sr.out_slot].busy;
            sr.out_slot = -1;
        }
        load_direct(lk, items);
        return true;
    }
    if (free_now + (int64_t)evicting_mapped_ >= (int64_t)need) return false;     // on its way: the reap will free it
    const uint64_t shortage = (uint64_t)((int64_t)need - free_now - (int64_t)evicting_mapped_);
    std::vector<uint32_t> victims;
    uint64_t evictable = 0;
    const uint64_t epoch_before = release_epoch_;
    CUresult r = choose_victims(lk, shortage, &victims, &evictable);
    if (r != CUDA_SUCCESS) return fail(r);
    // the lock was released during the scan: the world may have moved on
    if (demand_q_.empty() || demand_q_.front().row != row || side_[row].phase != PH_QUEUED) return true;
    if (evictable < shortage) {
        if (demand_row_ != row) { demand_row_ = row; demand_since_ns_ = mono_ns(); }
        // pins were released while the scan ran: it saw them, look again (bounded like the waits below: another thread that
        // launches on resident buffers in a tight loop releases pins all the time without ever making room)
        if (release_epoch_ != epoch_before && mono_ns() - demand_since_ns_ < 2000000000ull) return true;
    }
    if (evictable < shortage) {
        bool in_flight = evicting_mapped_ > 0 || !zombies_.empty();
        if (in_flight) return false;
        if (open_admissions_ > 0) {
            // what is in the way is pinned by admissions of other threads whose launch is being issued right now: their pins go
            // with their note_use. Evict what can be evicted and look again — for a bounded time (an application that nests
            // acquisitions on one thread would wait for itself)
            if (demand_row_ != row) { demand_row_ = row; demand_since_ns_ = mono_ns(); }
            if (mono_ns() - demand_since_ns_ < 2000000000ull) {
                if (!victims.empty()) evict_direct(lk, victims);
                return false;
            }
        }
        if (sibling_engines_ > 1 && need <= fair_share_) {
            // the room is held by a sibling process of the container: its pager gives it up as soon as it sees our live bytes
            // (fair share of the common quota); evict what we can meanwhile and keep waiting — but not for ever (a sibling
            // that is stopped in a debugger never shrinks)
            if (demand_row_ != row) { demand_row_ = row; demand_since_ns_ = mono_ns(); }
            if (mono_ns() - demand_since_ns_ < 20000000000ull) {
                if (!victims.empty()) evict_direct(lk, victims);
                return false;
            }
        }
        if (cfg_.host_backed) LOG_INFO("no room for row %d (need %lu MiB more, evictable %lu MiB): it is used where it is (host-mapped)", row,
                                       (unsigned long)(shortage >> 20), (unsigned long)(evictable >> 20));
        else LOG_ERROR("resident quota %lu MiB cannot hold the working set of this launch (need %lu MiB more, evictable %lu MiB; %d admissions in flight, waited %lu ms)",
                  (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(shortage >> 20), (unsigned long)(evictable >> 20), open_admissions_,
                  (unsigned long)(demand_row_ == row ? (mono_ns() - demand_since_ns_) / 1000000 : 0));
        if (std::getenv("VGPU_SWAP_DEBUG_DUMP"))
            for (size_t i = 0; i < rows_.size(); i++)
                std::fprintf(stderr, "  row %zu state=%u pins=%d inplace=%d phase=%d locked=%d demand=%d hosted=%d mapped=%zu%s\n", i, rows_[i].state, side_[i].pins,
                             side_[i].inplace, (int)side_[i].phase, (int)side_[i].locked, (int)side_[i].demand, (int)side_[i].hosted, side_[i].mapped, (int)i == row ? "  <- demanded" : "");
        return fail(CUDA_ERROR_OUT_OF_MEMORY);
    }
    demand_row_ = -1;
    if (evicting_mapped_ == 0 && !cfg_.host_backed) {
        demand_q_.pop_front();
        swap_staged(lk, row, victims);           // latency path
    } else if (evict_direct(lk, victims) != CUDA_SUCCESS) {   // the pipeline is running: add to it and wait for the reap
        fail_demands_locked(CUDA_ERROR_OUT_OF_MEMORY);
    }
    return true;
}

bool SwapEngine::step_prefetch(Lock &lk) {
    std::vector<InItem> items;
    while (!prefetch_q_.empty() && items.size() < cfg_.batch_rows) {
        QEntry e = prefetch_q_.front();
        bool ok = (size_t)e.row < side_.size() && side_[e.row].gen == e.gen && side_[e.row].phase == PH_QUEUED && !side_[e.row].demand;
        if (ok && side_[e.row].inplace > 0) { side_[e.row].phase = PH_IDLE; queued_prefetch_bytes_ -= side_[e.row].mapped; ok = false; }   // in use where it is
        if (!ok) { prefetch_q_.pop_front(); continue; }
        Side &s = side_[e.row];
        if (free_phys_locked() < (int64_t)s.mapped || (budget_fn_ && !reserve_locked(s.mapped))) break;
        prefetch_q_.pop_front();
        queued_prefetch_bytes_ -= s.mapped;
        items.emplace_back();
        begin_load_locked(e.row, true, &items.back());
        if (s.out_slot >= 0 && ring_out_[s.out_slot].seq == s.out_seq) items.back().after = ring_out_[s.out_slot].busy;
        s.out_slot = -1;
    }
    if (items.empty()) return false;
    load_direct(lk, items);
    return true;
}

// Everything that is queued (demanded rows waiting for room, rows the predictor wants) must fit next to what is
// resident: evict the LRU rows ahead of need so that both DMA queues always have work.
bool SwapEngine::step_evict_ahead(Lock &lk) {
    uint64_t wanted = queued_prefetch_bytes_;
    for (const QEntry &e : demand_q_)
        if ((size_t)e.row < side_.size() && side_[e.row].gen == e.gen && side_[e.row].phase == PH_QUEUED && side_[e.row].demand) wanted += side_[e.row].mapped;
    // While the prefetch pipeline runs, keep some physical memory free AHEAD of the page-in queue: a page-in can then be
    // mapped and put on the wire the moment it is wished for, instead of one eviction + one unmap later — the H2D queue
    // stays as deep as the D2H queue, and a slow VMM call stalls neither.
    if (queued_prefetch_bytes_ + prefetched_bytes_ > 0) {
        uint64_t window = std::min<uint64_t>(cfg_.prefetch_bytes, cfg_.resident_cap / 4);
        wanted += cfg_.headroom_bytes == ~0ull ? window : std::min<uint64_t>(cfg_.headroom_bytes, cfg_.resident_cap / 4);
    }
    if (resident_mapped_ + wanted <= cfg_.resident_cap) return false;
    uint64_t shortage = resident_mapped_ + wanted - cfg_.resident_cap;
    std::vector<uint32_t> victims;
    uint64_t evictable = 0;
    CUresult crc = choose_victims(lk, shortage, &victims, &evictable);
    if (crc != CUDA_SUCCESS || victims.empty()) {
        // nothing can be evicted (everything resident is in use): wishes are dropped, demands keep waiting for a release
        bool had = !prefetch_q_.empty();
        drop_prefetch_queue_locked();
        return had;
    }
    if (evictable < shortage && wanted == queued_prefetch_bytes_ && wanted > 0) {
        // prefetch must never squeeze out what is in use: drop the tail of the wish list instead
        while (!prefetch_q_.empty() && evictable < shortage) {
            QEntry e = prefetch_q_.back();
            prefetch_q_.pop_back();
            if ((size_t)e.row >= side_.size() || side_[e.row].gen != e.gen || side_[e.row].phase != PH_QUEUED || side_[e.row].demand) continue;
            side_[e.row].phase = PH_IDLE;
            queued_prefetch_bytes_ -= side_[e.row].mapped;
            shortage = shortage > side_[e.row].mapped ? shortage - side_[e.row].mapped : 0;
        }
        if (shortage == 0) { victim_cache_.clear(); return true; }
    }
    if (evict_direct(lk, victims) != CUDA_SUCCESS) {
        // the pinned pool is exhausted: nothing more can be paged out
        drop_prefetch_queue_locked();
        fail_demands_locked(CUDA_ERROR_OUT_OF_MEMORY);
    }
    return true;
}

void SwapEngine::drop_prefetch_queue_locked() {
    for (const QEntry &e : prefetch_q_) {
        if ((size_t)e.row >= side_.size() || side_[e.row].gen != e.gen || side_[e.row].phase != PH_QUEUED || side_[e.row].demand) continue;
        side_[e.row].phase = PH_IDLE;
    }
    prefetch_q_.clear();
    queued_prefetch_bytes_ = 0;
}
void SwapEngine::fail_demands_locked(CUresult rc) {
    for (const QEntry &e : demand_q_) {
        if ((size_t)e.row >= side_.size() || side_[e.row].gen != e.gen || side_[e.row].phase != PH_QUEUED || !side_[e.row].demand) continue;
        side_[e.row].phase = PH_IDLE;
        side_[e.row].fail = rc;
        side_[e.row].demand = false;
    }
    demand_q_.clear();
    cv_admit_.notify_all();
}

void SwapEngine::pager_main() {
    const DriverTable &d = drv();
    d.cuCtxSetCurrent(ctx_);
    Lock lk(mu_);
    const auto poll = std::chrono::microseconds(env_u64("VGPU_SWAP_POLL_US", 40));
    bool trace_dumped = false;
    for (;;) {
        if (stop_) break;
        if (trace_want_ && !trace_dumped && trace_.size() >= 2u * trace_want_) {   // the driver is gone by the time atexit handlers run
            trace_dumped = true;
            lk.unlock(); dump_trace(stderr); lk.lock();
        }
        uint64_t t0 = mono_ns();
        bool progress = false;
        uint64_t ts = t0, te;
        bool p0 = step_zombies(lk);     te = mono_ns(); if (p0) pst_.pager_step_ns[0] += te - ts; ts = te;
        bool p1 = step_reap(lk);        te = mono_ns(); if (p1) pst_.pager_step_ns[1] += te - ts; ts = te;
        bool p2 = step_demand(lk);      te = mono_ns(); if (p2) pst_.pager_step_ns[2] += te - ts; ts = te;
        bool p3 = step_prefetch(lk);    te = mono_ns(); if (p3) pst_.pager_step_ns[3] += te - ts; ts = te;
        bool p4 = step_evict_ahead(lk); te = mono_ns(); if (p4) pst_.pager_step_ns[4] += te - ts;
        progress = p0 || p1 || p2 || p3 || p4;
        if (progress) { pst_.pager_busy_ns += te - t0; continue; }
        // a sibling process may have started, grown or be waiting for its share of the common quota: look at the region about
        // once a millisecond (one semaphore round trip)
        if (budget_fn_ && mono_ns() - budget_checked_ns_ > 1000000ull) reserve_locked(0);
        bool outstanding = !evicting_.empty() || !zombies_.empty() || !demand_q_.empty();
        if (!outstanding) {
            // idle: fold the profiling samples in while nobody waits for the link
            if (!prof_.empty() || span_read_ < span_next_) { lk.unlock(); harvest_prof(false); relock(lk); }
            flush_pager_stats_locked();
            pager_idle_ = true;
            cv_admit_.notify_all();
            if (budget_fn_) cv_pager_.wait_for(lk, std::chrono::milliseconds(sibling_engines_ > 1 ? 2 : 50), [&] { return stop_ || kick_; });
            else cv_pager_.wait(lk, [&] { return stop_ || kick_; });
            kick_ = false;
            pager_idle_ = false;
        } else {
            // GPU work outstanding (copies to reap, users to wait for): poll at a period far below one transfer
            cv_pager_.wait_for(lk, poll);
        }
    }
    flush_pager_stats_locked();
    pager_idle_ = true;
    cv_admit_.notify_all();
}

// ---------------------------------------------------------------------------------------------- public operations
void SwapEngine::set_resident_cap(uint64_t cap) {
    std::lock_guard<std::mutex> g(mu_);
    if (budget_fn_) { budget_checked_ns_ = 0; kick_pager_locked(); return; }   // the cap comes from the container's shared budget: just look again
    quota_cap_ = cap;
    // under physical pressure (see load_direct) the working cap stays at what the device could actually give
    cfg_.resident_cap = pressure_ ? std::min(cap, cfg_.resident_cap) : cap;
    kick_pager_locked();
}

CUresult SwapEngine::alloc(CUdeviceptr *dptr, size_t bytes) {
    if (!dptr || bytes == 0) return CUDA_ERROR_INVALID_VALUE;
    std::lock_guard<std::mutex> gate(gate_mu_);
    Lock lk(mu_);
    ScopedNs t_admit(&st_.host_admit_ns);
    if (cfg_.virtual_cap && live_bytes_.load() + bytes > cfg_.virtual_cap) {
        LOG_ERROR("Device %d OOM (virtual) %lu / %lu", dev_, (unsigned long)(live_bytes_.load() + bytes), (unsigned long)cfg_.virtual_cap);
        return CUDA_ERROR_OUT_OF_MEMORY;
    }
    uint64_t mapped = round_up(bytes, gran_);
    if (cfg_.resident_cap && mapped > cfg_.resident_cap && !budget_fn_) {      // (with sibling engines the cap moves: the pager decides)
        // a buffer is resident as a whole while a kernel uses it: one that exceeds the resident cap can never be admitted
        LOG_ERROR("Device %d OOM: a single %lu-byte buffer exceeds the resident cap of %lu bytes", dev_, (unsigned long)bytes, (unsigned long)cfg_.resident_cap);
        return CUDA_ERROR_OUT_OF_MEMORY;
    }
    uint64_t off;
    if (!va_alloc(mapped, &off)) { LOG_ERROR("swap arena exhausted"); return CUDA_ERROR_OUT_OF_MEMORY; }
    int row = new_row();
    uint32_t gen = side_[row].gen;
    rows_[row] = VgpuEntry{arena_ + off, bytes, ++tick_, VGPU_ST_PAGED_OUT, 0};
    side_[row] = Side{};
    Side &s = side_[row];
    s.gen = gen;
    s.mapped = mapped;
    s.va_off = off;
    s.pins = 1;                                  // until this call returns
    s.phase = PH_QUEUED;
    s.demand = true;
    mark_dirty(row);
    // live from now on: sibling engines of the container (other processes, same device) give up room for what is LIVE here
    live_bytes_ += bytes;
    live_mapped_ += mapped;
    host_need_ += round_up(bytes, 256);
    demand_q_.push_back(QEntry{row, s.gen});
    kick_pager_locked();
    {
        ScopedNs t_wait(&st_.host_wait_ns);
        cv_admit_.wait(lk, [&] { return side_[row].phase == PH_IDLE && !side_[row].demand; });
    }
    Side &s2 = side_[row];
    if (!(rows_[row].state & VGPU_ST_RESIDENT)) {
        CUresult rc = s2.fail != CUDA_SUCCESS ? s2.fail : CUDA_ERROR_OUT_OF_MEMORY;
        s2.pins = 0;
        live_bytes_ -= bytes;
        live_mapped_ -= mapped;
        host_need_ -= std::min<uint64_t>(host_need_, round_up(bytes, 256));
        retire_row_locked(row);
        return rc;
    }
    s2.pins = 0;
    rows_[row].state = VGPU_ST_RESIDENT | (s2.locked ? VGPU_ST_PINNED : 0u);
    for (uint64_t gidx = off / gran_; gidx < (off + mapped) / gran_; gidx++) owner_[gidx] = row;
    mark_dirty(row);
    uint64_t need_now = host_need_;
    *dptr = arena_ + off;
    publish_locked();
    lk.unlock();
    // only what cannot stay resident will ever be paged out: grow the pinned pool once the live bytes pass the cap
    if (need_now > cfg_.resident_cap) want_host_pool(need_now);
    return CUDA_SUCCESS;
}

CUresult SwapEngine::free(CUdeviceptr dptr) {
    if (!owns(dptr)) return CUDA_ERROR_INVALID_VALUE;
    Lock lk(mu_);
    int row = owner_[(dptr - arena_) / gran_];
    if (row < 0 || rows_[row].base != dptr) return CUDA_ERROR_INVALID_VALUE;
    // the pager may be working on it with the lock released
    cv_admit_.wait(lk, [&] { return side_[row].phase != PH_LOADING && side_[row].phase != PH_EVICTING; });
    Side &s = side_[row];
    if (s.phase == PH_QUEUED) {                  // stale queue entries are recognised by their phase
        if (!s.demand) queued_prefetch_bytes_ -= s.mapped;
        else { s.demand = false; s.fail = CUDA_ERROR_INVALID_VALUE; cv_admit_.notify_all(); }   // freed under a waiting admission
        s.phase = PH_IDLE;
    }
    if (s.prefetched) { s.prefetched = false; prefetched_bytes_ -= s.mapped; st_.prefetch_wasted++; }
    for (uint64_t gidx = s.va_off / gran_; gidx < (s.va_off + s.mapped) / gran_; gidx++) owner_[gidx] = -1;
    live_bytes_ -= rows_[row].size;
    live_mapped_ -= s.mapped;
    host_need_ -= std::min<uint64_t>(host_need_, round_up(rows_[row].size, 256));
    host_want_.store(host_need_);
    s.pins = 0;                                  // rows pinned for a stream capture are never unpinned by note_use
    s.locked = false;
    s.zombie_len = round_up(rows_[row].size, 256);
    if (last_row_ == row) last_row_ = -1;
    if (rows_[row].state & VGPU_ST_RESIDENT) {
        // still mapped, maybe still in use by queued work: the pager unmaps it once its last users are done and only then
        // hands the address range out again
        rows_[row].state = VGPU_ST_FREE;
        rows_[row].size = 0;
        mark_dirty(row);
        s.phase = PH_ZOMBIE;
        zombies_.push_back((uint32_t)row);
        kick_pager_locked();
    } else if (s.has_hh) {
        // host-backed mode: the range (and the alias range) still map the host backing — VMM calls are the pager's
        rows_[row].state = VGPU_ST_FREE;
        rows_[row].size = 0;
        mark_dirty(row);
        s.phase = PH_ZOMBIE;
        zombies_.push_back((uint32_t)row);
        kick_pager_locked();
    } else {
        if (s.has_host) release_host_range(s.host_off, round_up(rows_[row].size, 256));
        s.has_host = false;
        retire_row_locked(row);                    // a staged D2H into its block may still be queued: harmless, the block's
                                                   // next writer queues behind it on the same stream
    }
    publish_locked();
    return CUDA_SUCCESS;
}

CUresult SwapEngine::ensure_resident(const int *rows, int n, CUstream stream) {
    const DriverTable &d = drv();
    Lock lk(mu_);
    ScopedNs t_admit(&st_.host_admit_ns);
    // an admission that has to wait for the pager takes the gate first: two of them could otherwise each hold (pin) part
    // of what the other needs
    bool gated = false;
    auto any_missing = [&] {
        for (int i = 0; i < n; i++) if (!(rows_[rows[i]].state & VGPU_ST_RESIDENT)) return true;
        return false;
    };
    if (any_missing()) {
        lk.unlock();
        gate_mu_.lock();
        gated = true;
        lk.lock();
    }
    st_.admissions++;
    tick_++;
    std::vector<int> missing;
    for (int i = 0; i < n; i++) {
        int r = rows[i];
        Side &s = side_[r];
        observe_touch(r);
        rows_[r].last_touch = tick_;
        s.pins++;
        if (rows_[r].state & VGPU_ST_RESIDENT) {
            rows_[r].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED;
            if (s.prefetched) { s.prefetched = false; prefetched_bytes_ -= s.mapped; st_.prefetch_hits++; st_.faults++; }
        } else {
            missing.push_back(r);
        }
        mark_dirty(r);
    }
    CUresult rc = CUDA_SUCCESS;
    if (!missing.empty()) {
        st_.faults += missing.size();
        st_.demand_waits++;
        if (!ctx_warned_) {
            // The engine's module, side streams and page-in events live in the context that was current when the first
            // swappable allocation created it; another context of the same device can use the buffers all the same (events
            // are waited on across contexts, last-use events are kept per context, see note_use). Said once, for the log.
            CUcontext cur = nullptr;
            if (d.cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur && cur != ctx_) {
                ctx_warned_ = true;
                LOG_INFO("device %d: swappable buffers are used from context %p, the swap engine lives in %p", dev_, (void *)cur, (void *)ctx_);
            }
        }
        uint64_t need = 0, pinned = 0;
        for (int i = 0; i < n; i++) if (rows_[rows[i]].state & VGPU_ST_RESIDENT) pinned += side_[rows[i]].mapped;
        // with sibling engines the cap moves (they give room up as they see our demand): judge against the fair share then
        const uint64_t room = budget_fn_ && sibling_engines_ > 1 ? std::max(cfg_.resident_cap, fair_share_) : cfg_.resident_cap;
        std::vector<int> inplace;
        if (cfg_.host_backed) {
            // Host-backed mode: a launch whose operands do not fit the quota together still runs (under UVM, the reference's
            // swap, it would thrash but work): as many operands as fit are paged in, in argument order; the others are used
            // WHERE THEY ARE — a paged-out row's own range maps its host backing — at PCIe speed. The pager leaves such a row
            // alone until this use has been recorded and has completed (Side::inplace, load_direct). A row that somebody is
            // using in place right now (or for good: a captured graph's operand) cannot move: it is used in place here too.
            uint64_t take = pinned;
            std::vector<int> fit;
            for (int r : missing) {
                if (side_[r].inplace == 0 && take + side_[r].mapped <= room) { take += side_[r].mapped; fit.push_back(r); }
                else inplace.push_back(r);
            }
            missing.swap(fit);
            for (int r : inplace) {
                Side &s = side_[r];
                s.inplace++;
                if (s.phase == PH_QUEUED && !s.demand) { s.phase = PH_IDLE; queued_prefetch_bytes_ -= s.mapped; }    // a wish for it: dropped
            }
            st_.inplace_uses += inplace.size();
            st_.faults -= inplace.size();
        }
        for (int r : missing) need += side_[r].mapped;
        if (need + pinned > room) {
            LOG_ERROR("resident quota %lu MiB cannot hold the working set of this launch (%lu MiB)", (unsigned long)(cfg_.resident_cap >> 20),
                      (unsigned long)((need + pinned) >> 20));
            rc = CUDA_ERROR_OUT_OF_MEMORY;
        } else {
            for (int r : missing) {
                Side &s = side_[r];
                s.fail = CUDA_SUCCESS;
                if (s.phase == PH_IDLE) { s.phase = PH_QUEUED; s.demand = true; demand_q_.push_back(QEntry{r, s.gen}); }
                else if (s.phase == PH_QUEUED && !s.demand) { s.demand = true; queued_prefetch_bytes_ -= s.mapped; demand_q_.push_back(QEntry{r, s.gen}); }
                else s.demand = true;            // loading or on its way out: the pager brings it (back) in
            }
            kick_pager_locked();
            {
                ScopedNs t_wait(&st_.host_wait_ns);
                cv_admit_.wait(lk, [&] {
                    for (int r : missing) {
                        const Side &s = side_[r];
                        bool settled = s.phase == PH_IDLE && !s.demand;
                        if (!settled) return false;
                    }
                    return true;
                });
            }
            for (int r : missing) {
                Side &s = side_[r];
                if (rows_[r].state & VGPU_ST_RESIDENT) rows_[r].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED;
                else if (cfg_.host_backed && s.fail == CUDA_ERROR_OUT_OF_MEMORY && s.hosted && s.phase == PH_IDLE) {
                    // no room could be made for it (what is resident is pinned — other launches in flight, captured graphs'
                    // operands): host-backed mode uses it where it is instead of failing
                    s.inplace++;
                    st_.inplace_uses++;
                    st_.faults--;
                    inplace.push_back(r);
                }
                else if (rc == CUDA_SUCCESS) rc = s.fail != CUDA_SUCCESS ? s.fail : CUDA_ERROR_OUT_OF_MEMORY;
                mark_dirty(r);
            }
            if (!inplace.empty()) {
                // a row on its way out or in is a hole for a moment: wait until the pager is done with it
                cv_admit_.wait(lk, [&] {
                    for (int r : inplace) if (side_[r].phase == PH_EVICTING || side_[r].phase == PH_LOADING) return false;
                    return true;
                });
                for (int r : inplace) {
                    Side &s = side_[r];
                    if (rows_[r].state & VGPU_ST_RESIDENT) { rows_[r].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED; mark_dirty(r); s.inplace--; }   // paged in meanwhile: an ordinary pinned operand
                    else if (!s.hosted && rc == CUDA_SUCCESS) rc = CUDA_ERROR_OUT_OF_MEMORY;   // its host mapping could not be made (reap)
                }
            }
        }
        if (rc != CUDA_SUCCESS) for (int r : inplace) if (!(rows_[r].state & VGPU_ST_RESIDENT)) side_[r].inplace--;
        if (rc != CUDA_SUCCESS) {
            for (int i = 0; i < n; i++) {
                Side &s = side_[rows[i]];
                if (--s.pins == 0 && !s.locked && (rows_[rows[i]].state & VGPU_ST_RESIDENT)) rows_[rows[i]].state = VGPU_ST_RESIDENT;
                mark_dirty(rows[i]);
            }
            if (gated) gate_mu_.unlock();
            return rc;
        }
    }
    schedule_prefetch();
    std::vector<CUevent> host_wait;
    for (int i = 0; i < n; i++) {
        Side &s = side_[rows[i]];
        if (!s.ready) continue;
        if (stream == kHostWait) { host_wait.push_back(s.ready); continue; }
        if (d.cuEventQuery(s.ready) == CUDA_SUCCESS) { put_event(s.ready); s.ready = nullptr; }
        else d.cuStreamWaitEvent(stream, s.ready, 0);
    }
    if (!missing.empty()) publish_locked();
    if (stream != kHostWait) open_admissions_++;     // (host-wait admissions — capture, pin, graph nodes — keep their pins for good)
    if (!host_wait.empty()) {
        // the rows are pinned, so their `ready` events stay theirs while the lock is released
        lk.unlock();
        for (CUevent e : host_wait) d.cuEventSynchronize(e);
        lk.lock();
    }
    if (gated) gate_mu_.unlock();
    return CUDA_SUCCESS;
}

void SwapEngine::note_use(const int *rows, int n, CUstream stream, bool writes, bool closes_admission) {
    const DriverTable &d = drv();
    std::lock_guard<std::mutex> g(mu_);
    if (closes_admission && open_admissions_ > 0) open_admissions_--;
    release_epoch_++;
    uint64_t seq = ++use_seq_;
    size_t ring = use_ring_.size();
    CUcontext cur = nullptr;
    d.cuCtxGetCurrent(&cur);
    if (!cur) cur = ctx_;
    // the slot's previous owner (seq - ring size) is only forgotten once it is known complete, see use_event()
    if (CUevent prev = use_ring_[seq % ring]) d.cuEventSynchronize(prev);
    CUevent ev = use_slot_event(cur, seq % ring);
    // A row remembers one outstanding use per stream (up to kMaxUses): a later eviction or free waits for ALL of them — work
    // queued on stream A must not lose its operand because stream B used it afterwards. A row used from more streams than
    // that gets its oldest use chained in front of this one.
    for (int i = 0; i < n; i++) {
        Side &s = side_[rows[i]];
        int k = 0;
        for (int j = 0; j < s.nuses; j++) {
            uint64_t q = s.uses[j];
            if (q + ring <= seq) continue;                                 // recycled: complete
            if (use_stream_[q % ring] == stream && use_ctx_[q % ring] == cur) continue;   // same stream: this use supersedes it
            s.uses[k++] = q;
        }
        s.nuses = k;
        if (s.nuses == kMaxUses) {
            if (CUevent old = use_event(s.uses[0])) d.cuStreamWaitEvent(stream, old, 0);
            for (int j = 1; j < s.nuses; j++) s.uses[j - 1] = s.uses[j];
            s.nuses--;
        }
    }
    if (ev) d.cuEventRecord(ev, stream);
    use_ring_[seq % ring] = ev;
    use_stream_[seq % ring] = stream;
    use_ctx_[seq % ring] = cur;
    for (int i = 0; i < n; i++) {
        Side &s = side_[rows[i]];
        s.uses[s.nuses++] = seq;
        if (writes && !s.read_mostly) s.dirty = true;
        if (s.inplace > 0 && !(rows_[rows[i]].state & VGPU_ST_RESIDENT)) s.inplace--;     // used where it was (host-mapped): the pager may move it again once this use is done
        if (s.pins > 0 && --s.pins == 0 && !s.locked && (rows_[rows[i]].state & VGPU_ST_RESIDENT)) {
            rows_[rows[i]].state = VGPU_ST_RESIDENT;
            mark_dirty(rows[i]);
        }
    }
}

void SwapEngine::advise_read_mostly(int row, bool on) {
    std::lock_guard<std::mutex> g(mu_);
    if (row_live(row)) side_[row].read_mostly = on;
}

CUresult SwapEngine::pin_resident(int row, bool on) {
    if (on) {
        CUresult rc = ensure_resident(&row, 1, kHostWait);
        if (rc != CUDA_SUCCESS) return rc;
        std::lock_guard<std::mutex> g(mu_);
        Side &s = side_[row];
        if (!(rows_[row].state & VGPU_ST_RESIDENT)) {            // (host-backed mode served it in place: it does not fit the quota at all)
            if (s.inplace > 0) s.inplace--;
            if (s.pins > 0) s.pins--;
            return CUDA_ERROR_OUT_OF_MEMORY;
        }
        s.locked = true;
        if (s.pins > 0) s.pins--;                 // the lock keeps it; the admission's pin is not needed
        s.dirty = true;                           // whoever reaches it behind the hook's back may write
        rows_[row].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED;
        mark_dirty(row);
        return CUDA_SUCCESS;
    }
    std::lock_guard<std::mutex> g(mu_);
    if (!row_live(row)) return CUDA_ERROR_INVALID_VALUE;
    Side &s = side_[row];
    s.locked = false;
    if (s.pins == 0 && (rows_[row].state & VGPU_ST_RESIDENT)) { rows_[row].state = VGPU_ST_RESIDENT; mark_dirty(row); }
    return CUDA_SUCCESS;
}

void SwapEngine::publish_locked() {
    resident_pub_.store(resident_mapped_ + evicting_mapped_, std::memory_order_relaxed);
    vgpu_swap_record_t *r = shared_;
    if (!r) return;
    __atomic_store_n(&r->page_out_bytes, st_.page_out_bytes, __ATOMIC_RELAXED);
    __atomic_store_n(&r->page_in_bytes, st_.page_in_bytes, __ATOMIC_RELAXED);
    __atomic_store_n(&r->evictions, st_.evictions, __ATOMIC_RELAXED);
    __atomic_store_n(&r->faults, st_.faults, __ATOMIC_RELAXED);
    __atomic_store_n(&r->resident_bytes, (uint64_t)(resident_mapped_ + evicting_mapped_), __ATOMIC_RELAXED);
    __atomic_store_n(&r->live_bytes, (uint64_t)live_bytes_.load(), __ATOMIC_RELAXED);
    __atomic_store_n(&r->host_bytes, (uint64_t)host_used_.load(), __ATOMIC_RELAXED);
}

CUresult SwapEngine::drain() {
    const DriverTable &d = drv();
    Lock lk(mu_);
    drop_prefetch_queue_locked();                // wishes, not work
    kick_pager_locked();
    cv_admit_.wait(lk, [&] { return stop_ || (pager_idle_ && !kick_ && demand_q_.empty() && evicting_.empty() && zombies_.empty()); });
    CUresult r = CUDA_SUCCESS, t;
    // the pager is parked on its condition variable: its private state may be touched from here
    for (CUstream s : {s_scan_, s_pack_, s_unpack_, s_out_, s_in_})
        if (s && (t = d.cuStreamSynchronize(s)) != CUDA_SUCCESS) r = t;
    harvest_spans();
    harvest_prof(true);
    flush_pager_stats_locked();
    return r;
}

SwapStats SwapEngine::stats() {
    std::lock_guard<std::mutex> g(mu_);
    SwapStats s = st_;
    s.resident_bytes = resident_mapped_ + evicting_mapped_;
    s.live_bytes = live_bytes_.load();
    s.host_bytes = host_used_.load();
    { std::lock_guard<std::mutex> h(host_mu_); s.host_slabs = host_slabs_; s.host_slabs_local = host_slabs_local_; }
    s.entries = rows_.size() - free_rows_.size() - zombies_.size();
    return s;
}

std::vector<VgpuEntry> SwapEngine::snapshot_table() {
    std::lock_guard<std::mutex> g(mu_);
    return rows_;
}

}  // namespace vgpu
