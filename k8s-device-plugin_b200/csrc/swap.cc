// swap.cc — see swap.h.
#include "swap.h"

#include <sched.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
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
    c.scan_lookahead = (uint32_t)env_u64("VGPU_SWAP_SCAN_LOOKAHEAD", 8);
    c.async_unmap = env_u64("VGPU_SWAP_ASYNC_UNMAP", 0) != 0;
    c.trace = (uint32_t)env_u64("VGPU_SWAP_TRACE", 0);
    c.spare_bytes = env_u64("VGPU_SWAP_SPARE_MB", 128) << 20;
    if (c.ring_slots < 2) c.ring_slots = 2;
    if (c.chunk_bytes < (1u << 20)) c.chunk_bytes = 1u << 20;
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
    scan_lookahead_ = cfg.scan_lookahead;
    trace_want_ = cfg.trace;
    trace_skip_ = 300;   // let the pipeline reach steady state first
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
    va_free_[0] = want;
    owner_.assign(want / gran_, -1);

    // highest priority: when the application saturates the SMs, the block scheduler hands freed slots to the swap
    // kernels first (they are short and the link is waiting on them)
    int prio_lo = 0, prio_hi = 0;
    if (d.cuCtxGetStreamPriorityRange) d.cuCtxGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (CUstream *s : {&s_scan_, &s_pack_, &s_unpack_, &s_out_, &s_in_}) {
        CUresult sr = d.cuStreamCreateWithPriority ? d.cuStreamCreateWithPriority(s, CU_STREAM_NON_BLOCKING, prio_hi)
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
    if (d.cuMemHostAlloc((void **)&h_tbl_stage_, (size_t)tbl_cap_ * sizeof(VgpuEntry), 0) != CUDA_SUCCESS) return false;
    scanner_.reset(new VictimScanner());
    if (scanner_->init(k_, tbl_cap_) != CUDA_SUCCESS) return false;
    use_ring_.resize(1024);
    for (auto &e : use_ring_) if (d.cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) return false;
    d.cuCtxGetCurrent(&ctx_);
    if (cfg_.async_unmap) reaper_ = std::thread([this] { reaper_main(); });
    LOG_INFO("swap engine dev %d (numa %d): resident cap %lu MiB, virtual cap %lu MiB, chunk %zu MiB x %d, arena %lu GiB, gran %zu",
             dev_, numa_node_, (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(cfg_.virtual_cap >> 20), cfg_.chunk_bytes >> 20,
             cfg_.ring_slots, (unsigned long)(cfg_.arena_bytes >> 30), gran_);
    return true;
}

SwapEngine::~SwapEngine() {
    const DriverTable &d = drv();
    if (!d.loaded) return;
    drain();
    if (reaper_.joinable()) {
        { std::lock_guard<std::mutex> g(rq_mu_); reaper_stop_ = true; }
        rq_cv_.notify_all();
        reaper_.join();
    }
    for (size_t i = 0; i < rows_.size(); i++) {
        if (rows_[i].state == VGPU_ST_FREE) continue;
        if (rows_[i].state & VGPU_ST_RESIDENT) { d.cuMemUnmap(rows_[i].base, side_[i].mapped); if (side_[i].has_handle) d.cuMemRelease(side_[i].handle); }
    }
    for (auto &p : phys_pool_) d.cuMemRelease(p.second);
    for (auto &s : ring_out_) { if (s.buf) d.cuMemFree_v2(s.buf); if (s.busy) d.cuEventDestroy_v2(s.busy); }
    for (auto &s : ring_in_) { if (s.buf) d.cuMemFree_v2(s.buf); if (s.busy) d.cuEventDestroy_v2(s.busy); }
    for (auto &sl : slabs_) if (sl.host) d.cuMemFreeHost(sl.host);
    for (auto &e : use_ring_) if (e) d.cuEventDestroy_v2(e);
    for (auto &e : ready_free_) d.cuEventDestroy_v2(e);
    if (d_tbl_) d.cuMemFree_v2(d_tbl_);
    if (h_tbl_stage_) d.cuMemFreeHost(h_tbl_stage_);
    if (arena_) d.cuMemAddressFree(arena_, cfg_.arena_bytes);
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

// pinned pool: global offset = slab_index << 44 | offset-in-slab
bool SwapEngine::host_alloc(size_t bytes, uint64_t *off) {
    const DriverTable &d = drv();
    bytes = round_up(bytes, 256);
    for (int attempt = 0; attempt < 2; attempt++) {
        // second try: wait for parked ranges (their H2D reads finish within ~1 ms) rather than pin a new slab (~100 ms) —
        // but only when enough bytes are parked to make the wait worthwhile
        if (attempt == 1) {
            uint64_t parked = 0;
            for (auto &p : pending_host_) parked += p.len;
            if (parked < bytes) break;
        }
        reap_pending_host(attempt == 1);
        for (size_t i = 0; i < slabs_.size(); i++) {
            uint64_t o;
            if (map_alloc(slabs_[i].free, bytes, &o)) {
                *off = ((uint64_t)i << 44) | o;
                host_used_ += bytes;
                return true;
            }
        }
        if (pending_host_.empty()) break;
    }
    size_t sb = std::max<size_t>(cfg_.slab_bytes, bytes);
    uint64_t have = 0;
    for (auto &s : slabs_) have += s.bytes;
    if (cfg_.host_pool_cap && have + sb > cfg_.host_pool_cap) {
        if (have + bytes > cfg_.host_pool_cap) return false;
        sb = bytes;
    }
    Slab s;
    CUresult r;
    {
        NodeAffinity on_node(numa_node_);
        if (numa_node_ >= 0) prefer_node(numa_node_);
        r = d.cuMemHostAlloc((void **)&s.host, sb, CU_MEMHOSTALLOC_PORTABLE);
        if (numa_node_ >= 0) default_policy();
    }
    if (r != CUDA_SUCCESS) { LOG_ERROR("pinned slab of %zu MiB failed: %d %s", sb >> 20, (int)r, cu_err(r)); return false; }
    st_.host_slabs++;
    if (numa_node_ >= 0) {
        int got = node_of(s.host + sb / 2);
        if (got == numa_node_) st_.host_slabs_local++;
        else LOG_WARN("pinned slab landed on NUMA node %d, GPU is on node %d: page traffic will cross the socket link", got, numa_node_);
    }
    s.bytes = sb;
    if (sb > bytes) s.free[bytes] = sb - bytes;
    slabs_.push_back(std::move(s));
    *off = ((uint64_t)(slabs_.size() - 1) << 44);
    host_used_ += bytes;
    return true;
}
unsigned char *SwapEngine::host_ptr(uint64_t off) { return slabs_[off >> 44].host + (off & ((1ull << 44) - 1)); }

// ---------------------------------------------------------------------------------------------- table
int SwapEngine::new_row() {
    if (!free_rows_.empty()) { int r = free_rows_.back(); free_rows_.pop_back(); return r; }
    rows_.push_back(VgpuEntry{});
    side_.push_back(Side{});
    return (int)rows_.size() - 1;
}
void SwapEngine::mark_dirty(int row) {
    dirty_lo_ = std::min<uint32_t>(dirty_lo_, (uint32_t)row);
    dirty_hi_ = std::max<uint32_t>(dirty_hi_, (uint32_t)row + 1);
}
CUresult SwapEngine::sync_table(CUstream s) {
    const DriverTable &d = drv();
    uint32_t n = (uint32_t)rows_.size();
    if (n > tbl_cap_) {
        uint32_t nc = tbl_cap_;
        while (nc < n) nc *= 2;
        CU_TRY(d.cuStreamSynchronize(s));
        d.cuMemFree_v2(d_tbl_);
        d.cuMemFreeHost(h_tbl_stage_);
        CU_TRY(d.cuMemAlloc_v2(&d_tbl_, (size_t)nc * sizeof(VgpuEntry)));
        CU_TRY(d.cuMemHostAlloc((void **)&h_tbl_stage_, (size_t)nc * sizeof(VgpuEntry), 0));
        scanner_.reset(new VictimScanner());
        CU_TRY(scanner_->init(k_, nc));
        tbl_cap_ = nc;
        dirty_lo_ = 0; dirty_hi_ = n;
    }
    if (dirty_lo_ >= dirty_hi_) return CUDA_SUCCESS;
    uint32_t lo = dirty_lo_, hi = std::min(dirty_hi_, n);
    // the pinned staging copy must not be overwritten while a previous upload is in flight
    CU_TRY(d.cuStreamSynchronize(s));
    std::memcpy(h_tbl_stage_ + lo, rows_.data() + lo, (size_t)(hi - lo) * sizeof(VgpuEntry));
    CU_TRY(d.cuMemcpyHtoDAsync_v2(d_tbl_ + (size_t)lo * sizeof(VgpuEntry), h_tbl_stage_ + lo, (size_t)(hi - lo) * sizeof(VgpuEntry), s));
    dirty_lo_ = UINT32_MAX; dirty_hi_ = 0;
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

// ---------------------------------------------------------------------------------------------- physical memory
void SwapEngine::trim_phys_pool(uint64_t need) {
    const DriverTable &d = drv();
    const uint64_t allowed = cfg_.resident_cap + (cfg_.async_unmap ? cfg_.spare_bytes : 0);
    while (!phys_pool_.empty() && resident_mapped_ + evicting_mapped_ + phys_pool_bytes_ + need > allowed) {
        auto it = std::prev(phys_pool_.end());
        d.cuMemRelease(it->second);
        phys_pool_bytes_ -= it->first;
        phys_pool_.erase(it);
    }
}
CUresult SwapEngine::get_phys(size_t mapped, CUmemGenericAllocationHandle *h) {
    const DriverTable &d = drv();
    for (;;) {
        auto it = phys_pool_.find(mapped);
        if (it != phys_pool_.end()) {
            *h = it->second;
            phys_pool_bytes_ -= mapped;
            phys_pool_.erase(it);
            st_.phys_reuses++;
            return CUDA_SUCCESS;
        }
        // physical memory held = resident + victims awaiting their unmap + pooled handles; it may exceed the quota
        // by at most spare_bytes (a documented overhead like the staging rings) so that mapping the incoming row
        // does not have to wait for the reaper
        uint64_t held = resident_mapped_ + evicting_mapped_ + phys_pool_bytes_;
        uint64_t allowed = cfg_.resident_cap + (cfg_.async_unmap ? cfg_.spare_bytes : 0);
        if (held + mapped <= allowed) break;
        if (!phys_pool_.empty()) { trim_phys_pool(mapped); held = resident_mapped_ + evicting_mapped_ + phys_pool_bytes_; if (held + mapped <= allowed) break; }
        if (evicting_mapped_ == 0) break;            // nothing more will come back: create and let the driver decide
        reap_cv_.wait(mu_);                          // a reaper batch will return handles
    }
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev_;
    CUresult r = d.cuMemCreate(h, mapped, &prop, 0);
    if (r == CUDA_ERROR_OUT_OF_MEMORY && !phys_pool_.empty()) {
        for (auto &p : phys_pool_) d.cuMemRelease(p.second);
        phys_pool_.clear();
        phys_pool_bytes_ = 0;
        r = d.cuMemCreate(h, mapped, &prop, 0);
    }
    if (r == CUDA_SUCCESS) st_.phys_creates++;
    return r;
}
CUresult SwapEngine::map_row(int row) {
    const DriverTable &d = drv();
    wait_not_evicting(row);                      // its own previous mapping may still be queued at the reaper
    CUmemGenericAllocationHandle h;
    CUresult r = get_phys(side_[row].mapped, &h);   // may wait for the reaper with mu_ released: take references after
    for (int attempt = 0; r == CUDA_ERROR_OUT_OF_MEMORY && attempt < 4; attempt++) {
        // The device cannot give what the quota promises: on an overcommitted GPU other containers hold the rest (the
        // reference leaves this case to UVM, which pages between processes). Live within what we have: lower the working
        // cap to what is mapped now and make the room out of our own least recently used rows.
        const uint64_t need = side_[row].mapped;
        uint64_t held = resident_mapped_ + evicting_mapped_;
        if (held < need) break;                        // nothing of ours left to give up
        // what the device can still give on top of what we hold (the pool was already released by get_phys)
        size_t fr = 0, tot = 0;
        if (d.cuMemGetInfo_v2(&fr, &tot) == CUDA_SUCCESS) held += (uint64_t)fr / gran_ * gran_;
        if (!pressure_ || cfg_.resident_cap > held) {
            if (!pressure_) {
                quota_cap_ = std::max(quota_cap_, cfg_.resident_cap);
                LOG_WARN("device %d: physical memory exhausted below the quota; resident cap %lu -> %lu MiB", dev_,
                         (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(held >> 20));
            } else {
                LOG_INFO("device %d: upward probe failed; resident cap back to %lu MiB", dev_, (unsigned long)(held >> 20));
            }
            cfg_.resident_cap = held;
            pressure_ = true;
        }
        pressure_events_++;
        if (make_room(need, true) != CUDA_SUCCESS) break;
        r = get_phys(need, &h);
    }
    if (r != CUDA_SUCCESS) return r;
    ScopedNs t(&st_.host_vmm_ns);
    Side &s = side_[row];
    r = d.cuMemMap(rows_[row].base, s.mapped, 0, h, 0);
    if (r != CUDA_SUCCESS) { d.cuMemRelease(h); LOG_ERROR("cuMemMap failed: %d %s", (int)r, cu_err(r)); return r; }
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = d.cuMemSetAccess(rows_[row].base, s.mapped, &acc, 1);
    if (r != CUDA_SUCCESS) { d.cuMemUnmap(rows_[row].base, s.mapped); d.cuMemRelease(h); LOG_ERROR("cuMemSetAccess failed: %d %s", (int)r, cu_err(r)); return r; }
    s.handle = h;
    s.has_handle = true;
    resident_mapped_ += s.mapped;
    return CUDA_SUCCESS;
}
void SwapEngine::unmap_row(int row) {
    const DriverTable &d = drv();
    ScopedNs t(&st_.host_vmm_ns);
    Side &s = side_[row];
    d.cuMemUnmap(rows_[row].base, s.mapped);
    phys_pool_.emplace(s.mapped, s.handle);
    phys_pool_bytes_ += s.mapped;
    s.has_handle = false;
    resident_mapped_ -= s.mapped;
}

// ---------------------------------------------------------------------------------------------- events / rings
CUevent SwapEngine::use_event(uint64_t seq) {
    if (seq == 0) return nullptr;
    if (seq + use_ring_.size() <= use_seq_) return nullptr;  // slot was recycled: that use is known complete (note_use)
    return use_ring_[seq % use_ring_.size()];
}
SwapEngine::Slot &SwapEngine::acquire_slot(std::vector<Slot> &ring, int *cursor) {
    Slot &s = ring[*cursor];
    *cursor = (*cursor + 1) % (int)ring.size();
    if (s.used) { ScopedNs t(&st_.host_ring_ns); drv().cuEventSynchronize(s.busy); }  // back-pressure: the only place the host waits for the link
    s.used = true;
    s.seq++;
    return s;
}
CUdeviceptr SwapEngine::next_span(bool unpack) {
    if (!cfg_.profile) return 0;
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
        (span_unpack_[span_read_ + i] ? st_.unpack_span_ms : st_.pack_span_ms) += (double)(b - a) / 1e6;
    }
    span_read_ = span_next_;
}

void SwapEngine::prof_begin(CUstream s, CUevent *a) {
    *a = nullptr;
    if (!cfg_.profile) return;
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
                if (p.unpack) { st_.unpack_ms += ms; st_.unpack_bytes += p.bytes; }
                else { st_.pack_ms += ms; st_.pack_bytes += p.bytes; }
            }
            d.cuEventDestroy_v2(p.a);
            d.cuEventDestroy_v2(p.b);
        } else {
            prof_[keep++] = p;
        }
    }
    prof_.resize(keep);
}

// ---------------------------------------------------------------------------------------------- page-out / page-in
CUresult SwapEngine::page_out(const std::vector<uint32_t> &victims, bool finish) {
    const DriverTable &d = drv();
    if (victims.empty()) return CUDA_SUCCESS;
    // one contiguous pinned block for the whole batch so the staging layout equals the host layout (one DMA per
    // chunk even when the victims are many small buffers); fall back to one block per victim when fragmented
    std::vector<uint64_t> len(victims.size());
    uint64_t total = 0;
    for (size_t i = 0; i < victims.size(); i++) { len[i] = round_up(rows_[victims[i]].size, 256); total += len[i]; }
    uint64_t block = 0;
    bool contiguous = host_alloc(total, &block);
    if (contiguous) {
        uint64_t o = 0;
        for (size_t i = 0; i < victims.size(); i++) { side_[victims[i]].host_off = block + o; o += len[i]; }
    } else {
        if (victims.size() == 1) { LOG_ERROR("pinned host pool exhausted (%lu MiB in use)", (unsigned long)(host_used_ >> 20)); return CUDA_ERROR_OUT_OF_MEMORY; }
        for (uint32_t v : victims) { CUresult r = page_out(std::vector<uint32_t>{v}, false); if (r != CUDA_SUCCESS) return r; }
        return finish ? page_out_finish() : CUDA_SUCCESS;
    }
    // order the pack behind the victims' last users
    for (uint32_t v : victims) {
        if (CUevent e = use_event(side_[v].use_seq)) d.cuStreamWaitEvent(s_pack_, e, 0);
        if (side_[v].ready) d.cuStreamWaitEvent(s_pack_, side_[v].ready, 0);   // its own page-in may still be in flight
    }

    std::vector<PackSegment> segs;
    Slot *slot = nullptr;
    uint64_t pos = 0, host_pos = 0, slot_host_start = 0;
    auto flush = [&]() -> CUresult {
        if (!slot || segs.empty()) return CUDA_SUCCESS;
        int launches = 0;
        CUevent pa;
        prof_begin(s_pack_, &pa);
        CUresult r = launch_pack(k_, segs.data(), segs.size(), s_pack_, &launches, next_span(false));
        if (r != CUDA_SUCCESS) return r;
        prof_end(s_pack_, pa, false, pos);
        st_.pack_launches += launches;
        // slot.busy doubles as "packed" marker for the copy stream, then is re-recorded as "drained"
        CU_TRY(d.cuEventRecord(slot->busy, s_pack_));
        CU_TRY(d.cuStreamWaitEvent(s_out_, slot->busy, 0));
        if (tr_) trace_mark(&tr_->d2h, s_out_);
        CU_TRY(d.cuMemcpyDtoHAsync_v2(host_ptr(block) + slot_host_start, slot->buf, pos, s_out_));
        if (tr_) trace_mark(&tr_->d2h, s_out_);
        CU_TRY(d.cuEventRecord(slot->busy, s_out_));
        st_.page_out_bytes += pos;
        segs.clear();
        slot = nullptr;
        return CUDA_SUCCESS;
    };
    for (size_t i = 0; i < victims.size(); i++) {
        uint64_t off = 0;
        while (off < len[i]) {
            if (!slot) { slot = &acquire_slot(ring_out_, &cur_out_); pos = 0; slot_host_start = host_pos; }
            uint64_t piece = std::min<uint64_t>(len[i] - off, cfg_.chunk_bytes - pos);
            segs.push_back(PackSegment{rows_[victims[i]].base + off, slot->buf + pos, piece});
            side_[victims[i]].out_slot = (int)(slot - ring_out_.data());
            side_[victims[i]].out_seq = slot->seq;
            off += piece; pos += piece; host_pos += piece;
            if (pos == cfg_.chunk_bytes || segs.size() == VGPU_PACK_MAX_SEG) { CUresult r = flush(); if (r != CUDA_SUCCESS) return r; }
        }
    }
    { CUresult r = flush(); if (r != CUDA_SUCCESS) return r; }
    // the victims' physical pages may be recycled as soon as the LAST PACK has read them — not when the DMA is done.
    // The wait for that pack is deferred to page_out_finish() so the caller can put the page-in's H2D copies on
    // the wire first (they only need staging slots, not the physical memory being freed here).
    for (uint32_t v : victims) out_pending_.push_back(v);
    return finish ? page_out_finish() : CUDA_SUCCESS;
}

void SwapEngine::reaper_main() {
    const DriverTable &d = drv();
    d.cuCtxSetCurrent(ctx_);
    for (;;) {
        ReapJob job;
        {
            std::unique_lock<std::mutex> lk(rq_mu_);
            rq_cv_.wait(lk, [&] { return reaper_stop_ || !rq_.empty(); });
            if (rq_.empty()) return;
            job = std::move(rq_.front());
            rq_.pop_front();
            reaper_busy_ = true;
        }
        d.cuEventSynchronize(job.packed);                       // the last pack of the batch has read the victims
        for (size_t i = 0; i < job.rows.size(); i++) d.cuMemUnmap(job.bases[i], job.mapped[i]);   // no engine lock held
        {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t i = 0; i < job.rows.size(); i++) {
                Side &s = side_[job.rows[i]];
                phys_pool_.emplace(s.mapped, s.handle);
                phys_pool_bytes_ += s.mapped;
                s.has_handle = false;
                s.evicting = false;
                evicting_mapped_ -= s.mapped;
            }
            ready_free_.push_back(job.packed);
        }
        { std::lock_guard<std::mutex> lk(rq_mu_); reaper_busy_ = false; }
        reap_cv_.notify_all();
    }
}

void SwapEngine::wait_not_evicting(int row) {
    while (side_[row].evicting) reap_cv_.wait(mu_);            // mu_ is held by the caller; released while waiting
}

CUresult SwapEngine::page_out_finish() {
    const DriverTable &d = drv();
    if (out_pending_.empty()) return CUDA_SUCCESS;
    CUevent packed = get_event();
    if (!packed) return CUDA_ERROR_OUT_OF_MEMORY;
    CU_TRY(d.cuEventRecord(packed, s_pack_));
    if (cfg_.async_unmap) {
        ReapJob job;
        job.packed = packed;
        for (uint32_t v : out_pending_) {
            Side &sd = side_[v];
            job.rows.push_back(v); job.bases.push_back(rows_[v].base); job.mapped.push_back(sd.mapped);
            sd.evicting = true;
            sd.has_host = true;
            resident_mapped_ -= sd.mapped;
            evicting_mapped_ += sd.mapped;
            rows_[v].state = VGPU_ST_PAGED_OUT;
            rows_[v].host_slot = (uint32_t)(sd.host_off >> 12);
            if (sd.ready) { ready_free_.push_back(sd.ready); sd.ready = nullptr; }
            mark_dirty((int)v);
            st_.evictions++;
        }
        out_pending_.clear();
        { std::lock_guard<std::mutex> lk(rq_mu_); rq_.push_back(std::move(job)); }
        rq_cv_.notify_one();
        return CUDA_SUCCESS;
    }
    { ScopedNs t(&st_.host_packsync_ns); CU_TRY(d.cuEventSynchronize(packed)); }
    ready_free_.push_back(packed);
    for (uint32_t v : out_pending_) {
        unmap_row((int)v);
        side_[v].has_host = true;
        rows_[v].state = VGPU_ST_PAGED_OUT;
        rows_[v].host_slot = (uint32_t)(side_[v].host_off >> 12);
        if (side_[v].ready) { ready_free_.push_back(side_[v].ready); side_[v].ready = nullptr; }
        mark_dirty((int)v);
        st_.evictions++;
    }
    out_pending_.clear();
    return CUDA_SUCCESS;
}

// Page-in is planned first (which pieces of which rows travel in which staging-slot job), then runs in two phases:
//   stage  : host -> staging H2D copies for as many jobs as there are free staging slots, WITHOUT blocking. Needs
//            neither the rows' physical memory nor their mappings, so it is issued before the victims are unmapped
//            and overlaps the VMM calls (0.3-0.8 ms per remap on B200, profiles/README.md) with the transfer;
//   finish : map the rows, then for every job (staged or not) launch the unpack kernel behind its H2D.
CUresult SwapEngine::page_in_plan(const std::vector<int> &rows) {
    in_jobs_.clear();
    InJob cur;
    uint64_t pos = 0;
    auto close = [&]() { if (!cur.segs.empty()) { cur.bytes = pos; in_jobs_.push_back(std::move(cur)); cur = InJob(); pos = 0; } };
    for (int r : rows) {
        uint64_t len = round_up(rows_[r].size, 256), off = 0;
        while (off < len) {
            uint64_t piece = std::min<uint64_t>(len - off, cfg_.chunk_bytes - pos);
            unsigned char *src = host_ptr(side_[r].host_off) + off;
            if (!cur.runs.empty() && cur.runs.back().src + cur.runs.back().len == src && cur.runs.back().pos + cur.runs.back().len == pos) cur.runs.back().len += piece;
            else cur.runs.push_back(InRun{src, pos, piece});
            cur.segs.push_back(PackSegment{pos, rows_[r].base + off, piece});   // src = offset inside the slot, fixed up at issue time
            off += piece; pos += piece;
            if (off == len) cur.done_rows.push_back(r);
            if (pos == cfg_.chunk_bytes || cur.segs.size() == VGPU_PACK_MAX_SEG) close();
        }
    }
    close();
    return CUDA_SUCCESS;
}

CUresult SwapEngine::in_issue_copies(InJob &j) {
    const DriverTable &d = drv();
    if (tr_) trace_mark(&tr_->h2d, s_in_);
    for (const InRun &r : j.runs) {
        CU_TRY(d.cuMemcpyHtoDAsync_v2(j.slot->buf + r.pos, r.src, r.len, s_in_));
        st_.page_in_bytes += r.len;
    }
    if (tr_) trace_mark(&tr_->h2d, s_in_);
    CU_TRY(d.cuEventRecord(j.slot->busy, s_in_));     // "loaded"; re-recorded as "unpacked" in finish
    return CUDA_SUCCESS;
}

CUresult SwapEngine::page_in_stage(const std::vector<int> &rows) {
    const DriverTable &d = drv();
    for (int r : rows) {
        // a row that was paged out moments ago: its D2H may still be in flight -> order the H2D behind that chunk
        // only (not behind the whole page-out queue, which would serialise the two link directions)
        Side &s = side_[r];
        if (s.out_slot >= 0 && ring_out_[s.out_slot].seq == s.out_seq) CU_TRY(d.cuStreamWaitEvent(s_in_, ring_out_[s.out_slot].busy, 0));
        s.out_slot = -1;
    }
    CU_TRY(page_in_plan(rows));
    size_t staged = 0;
    for (InJob &j : in_jobs_) {
        // A slot staged here holds data until page_in_finish() has launched its unpack; its `busy` event meanwhile only
        // says "loaded". So never wrap around onto a slot this admission has already filled — an H2D that completed
        // quickly would otherwise let job N+ring overwrite job N's staging before it is unpacked.
        if (staged == ring_in_.size()) break;
        Slot &s = ring_in_[cur_in_];
        if (s.used && d.cuEventQuery(s.busy) != CUDA_SUCCESS) break;   // never block here
        cur_in_ = (cur_in_ + 1) % (int)ring_in_.size();
        s.used = true;
        s.seq++;
        j.slot = &s;
        staged++;
        CU_TRY(in_issue_copies(j));
    }
    return CUDA_SUCCESS;
}

CUresult SwapEngine::page_in_finish(const std::vector<int> &rows) {
    const DriverTable &d = drv();
    for (int r : rows) {
        CUresult rc = map_row(r);
        if (rc != CUDA_SUCCESS) return rc;
    }
    std::vector<PendingHost> fresh;
    for (InJob &j : in_jobs_) {
        if (!j.slot) {
            j.slot = &acquire_slot(ring_in_, &cur_in_);
            CU_TRY(in_issue_copies(j));
        }
        for (PackSegment &sg : j.segs) sg.src += j.slot->buf;
        CU_TRY(d.cuStreamWaitEvent(s_unpack_, j.slot->busy, 0));
        int launches = 0;
        CUevent pa;
        prof_begin(s_unpack_, &pa);
        CUresult r = launch_pack(k_, j.segs.data(), j.segs.size(), s_unpack_, &launches, next_span(true));
        if (r != CUDA_SUCCESS) return r;
        prof_end(s_unpack_, pa, true, j.bytes);
        st_.unpack_launches += launches;
        CU_TRY(d.cuEventRecord(j.slot->busy, s_unpack_));
        for (int row : j.done_rows) {
            CUevent ev = get_event();
            if (!ev) return CUDA_ERROR_OUT_OF_MEMORY;
            CU_TRY(d.cuEventRecord(ev, s_unpack_));
            side_[row].ready = ev;
            fresh.push_back(PendingHost{side_[row].host_off, round_up(rows_[row].size, 256), nullptr});
            side_[row].has_host = false;
            rows_[row].state = VGPU_ST_RESIDENT;
            mark_dirty(row);
        }
    }
    in_jobs_.clear();
    if (!fresh.empty()) {
        // the pinned ranges go back to the pool only after the H2D copies above have READ them (a later page-out
        // would otherwise overwrite them): parked with one event, reaped by host_alloc() once it has fired
        CUevent ev = get_event();
        if (!ev) return CUDA_ERROR_OUT_OF_MEMORY;
        CU_TRY(d.cuEventRecord(ev, s_in_));
        for (size_t i = 0; i < fresh.size(); i++) { fresh[i].done = (i + 1 == fresh.size()) ? ev : nullptr; pending_host_.push_back(fresh[i]); }
    }
    return CUDA_SUCCESS;
}

CUevent SwapEngine::get_event() {
    if (!ready_free_.empty()) { CUevent e = ready_free_.back(); ready_free_.pop_back(); return e; }
    CUevent e = nullptr;
    if (drv().cuEventCreate(&e, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) return nullptr;
    return e;
}

void SwapEngine::reap_pending_host(bool wait) {
    const DriverTable &d = drv();
    // entries are in issue order; an entry without its own event completes with the next one that has one
    size_t done_upto = 0;
    for (size_t i = 0; i < pending_host_.size(); i++) {
        if (!pending_host_[i].done) continue;
        if (wait) d.cuEventSynchronize(pending_host_[i].done);
        if (d.cuEventQuery(pending_host_[i].done) != CUDA_SUCCESS) break;
        done_upto = i + 1;
    }
    for (size_t i = 0; i < done_upto; i++) {
        release_host_range(pending_host_[i].off, pending_host_[i].len);
        if (pending_host_[i].done) ready_free_.push_back(pending_host_[i].done);
    }
    pending_host_.erase(pending_host_.begin(), pending_host_.begin() + done_upto);
}

void SwapEngine::release_host_range(uint64_t off, uint64_t len) {
    // page_out carves one block per batch; every victim gives back exactly its own 256-byte-granular sub-range
    map_free(slabs_[off >> 44].free, off & ((1ull << 44) - 1), len);
    host_used_ = host_used_ > len ? host_used_ - len : 0;
}

CUresult SwapEngine::make_room(uint64_t need_mapped, bool finish) {
    if (need_mapped > cfg_.resident_cap) return CUDA_ERROR_OUT_OF_MEMORY;
    if (resident_mapped_ + need_mapped <= cfg_.resident_cap) return CUDA_SUCCESS;
    const uint64_t deficit = resident_mapped_ + need_mapped - cfg_.resident_cap;
    std::vector<uint32_t> victims;
    uint64_t freed = 0;
    auto valid = [&](const Cand &c) {
        return c.row < rows_.size() && rows_[c.row].state == VGPU_ST_RESIDENT && rows_[c.row].last_touch == c.touch && rows_[c.row].base == c.base;
    };
    // 1. leftovers of the previous scan (still the exact LRU prefix, see swap.h)
    {
        std::deque<Cand> keep = victim_cache_;
        std::vector<uint32_t> got;
        uint64_t f = 0;
        while (f < deficit && !keep.empty()) {
            Cand cnd = keep.front();
            keep.pop_front();
            if (!valid(cnd)) continue;
            got.push_back(cnd.row);
            f += rows_[cnd.row].size;
        }
        if (f >= deficit) { victims = got; freed = f; victim_cache_.swap(keep); st_.scan_cache_hits++; }
        else victim_cache_.clear();   // not enough left: start over from a fresh scan (nothing was consumed)
    }
    // 2. GPU scan, asking for `scan_lookahead_` times the deficit so the next evictions need no scan
    if (victims.empty()) {
        ScopedNs t(&st_.host_scan_ns);
        CUresult r = sync_table(s_scan_);
        if (r != CUDA_SUCCESS) return r;
        std::vector<uint32_t> found;
        uint64_t found_bytes = 0;
        bool insufficient = false;
        int launches = 0;
        uint64_t ask = deficit * (uint64_t)(scan_lookahead_ ? scan_lookahead_ : 1);
        r = scanner_->scan(d_tbl_, (uint32_t)rows_.size(), ask, tick_, s_scan_, &found, &found_bytes, &insufficient, &launches);
        st_.scan_launches += launches;
        st_.scans++;
        if (r != CUDA_SUCCESS) { LOG_ERROR("victim scan failed: %d %s", (int)r, cu_err(r)); return r; }
        if (found_bytes < deficit) {
            LOG_ERROR("resident quota %lu MiB cannot hold the working set of this launch (need %lu MiB more, evictable %lu MiB)",
                      (unsigned long)(cfg_.resident_cap >> 20), (unsigned long)(deficit >> 20), (unsigned long)(found_bytes >> 20));
            return CUDA_ERROR_OUT_OF_MEMORY;
        }
        // the kernel returns the prefix SET in index order; LRU order within it comes from the host mirror
        std::sort(found.begin(), found.end(), [&](uint32_t a, uint32_t b) {
            if (rows_[a].last_touch != rows_[b].last_touch) return rows_[a].last_touch < rows_[b].last_touch;
            return a < b;
        });
        size_t k = 0;
        while (k < found.size() && freed < deficit) { victims.push_back(found[k]); freed += rows_[found[k]].size; k++; }
        for (; k < found.size(); k++) victim_cache_.push_back(Cand{found[k], rows_[found[k]].last_touch, rows_[found[k]].base});
    }
    // the victims' physical handles are now pooled; get_phys() re-maps a same-size one under the incoming buffer
    // (no cuMemCreate/cuMemRelease in steady state) and only trims the pool when it has to create
    return page_out(victims, finish);
}

// ---------------------------------------------------------------------------------------------- public operations
CUresult SwapEngine::alloc(CUdeviceptr *dptr, size_t bytes) {
    if (!dptr || bytes == 0) return CUDA_ERROR_INVALID_VALUE;
    std::lock_guard<std::mutex> g(mu_);
    if (cfg_.virtual_cap && live_bytes_ + bytes > cfg_.virtual_cap) {
        LOG_ERROR("Device %d OOM (virtual) %lu / %lu", dev_, (unsigned long)(live_bytes_ + bytes), (unsigned long)cfg_.virtual_cap);
        return CUDA_ERROR_OUT_OF_MEMORY;
    }
    uint64_t mapped = round_up(bytes, gran_);
    if (cfg_.resident_cap && mapped > cfg_.resident_cap) {
        // a buffer is resident as a whole while a kernel uses it: one that exceeds the resident cap can never be admitted
        LOG_ERROR("Device %d OOM: a single %lu-byte buffer exceeds the resident cap of %lu bytes", dev_, (unsigned long)bytes, (unsigned long)cfg_.resident_cap);
        return CUDA_ERROR_OUT_OF_MEMORY;
    }
    uint64_t off;
    if (!va_alloc(mapped, &off)) { LOG_ERROR("swap arena exhausted"); return CUDA_ERROR_OUT_OF_MEMORY; }
    CUresult r = make_room(mapped);
    if (r != CUDA_SUCCESS) { va_free(off, mapped); return r; }
    int row = new_row();
    rows_[row] = VgpuEntry{arena_ + off, bytes, ++tick_, VGPU_ST_RESIDENT, 0};
    side_[row] = Side{};
    side_[row].mapped = mapped;
    side_[row].va_off = off;
    r = map_row(row);
    if (r != CUDA_SUCCESS) {
        rows_[row].state = VGPU_ST_FREE;
        free_rows_.push_back(row);
        va_free(off, mapped);
        return r;
    }
    for (uint64_t gidx = off / gran_; gidx < (off + mapped) / gran_; gidx++) owner_[gidx] = row;
    mark_dirty(row);
    live_bytes_ += bytes;
    *dptr = arena_ + off;
    publish_locked();
    return CUDA_SUCCESS;
}

CUresult SwapEngine::free(CUdeviceptr dptr) {
    const DriverTable &d = drv();
    if (!owns(dptr)) return CUDA_ERROR_INVALID_VALUE;
    std::lock_guard<std::mutex> g(mu_);
    int row = owner_[(dptr - arena_) / gran_];
    if (row < 0 || rows_[row].base != dptr) return CUDA_ERROR_INVALID_VALUE;
    wait_not_evicting(row);
    Side &s = side_[row];
    if (CUevent e = use_event(s.use_seq)) d.cuEventSynchronize(e);
    if (s.ready) { d.cuEventSynchronize(s.ready); ready_free_.push_back(s.ready); s.ready = nullptr; }
    if (rows_[row].state & VGPU_ST_RESIDENT) unmap_row(row);
    else if (s.has_host) release_host_range(s.host_off, round_up(rows_[row].size, 256));
    for (uint64_t gidx = s.va_off / gran_; gidx < (s.va_off + s.mapped) / gran_; gidx++) owner_[gidx] = -1;
    va_free(s.va_off, s.mapped);
    live_bytes_ -= rows_[row].size;
    s.pins = 0;                       // rows pinned for a stream capture are never unpinned by note_use
    rows_[row].state = VGPU_ST_FREE;
    rows_[row].size = 0;
    mark_dirty(row);
    free_rows_.push_back(row);
    publish_locked();
    return CUDA_SUCCESS;
}

CUresult SwapEngine::ensure_resident(const int *rows, int n, CUstream stream) {
    const DriverTable &d = drv();
    std::lock_guard<std::mutex> g(mu_);
    ScopedNs t_admit(&st_.host_admit_ns);
    st_.admissions++;
    tick_++;
    std::vector<int> missing;
    uint64_t need = 0;
    for (int i = 0; i < n; i++) {
        int r = rows[i];
        rows_[r].last_touch = tick_;
        side_[r].pins++;
        if (rows_[r].state & VGPU_ST_RESIDENT) rows_[r].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED;
        else { missing.push_back(r); need += side_[r].mapped; }
        mark_dirty(r);
    }
    if (!missing.empty()) {
        st_.faults += missing.size();
        if (!ctx_warned_) {
            // The engine's module, side streams and events live in the context that was current when the first swappable
            // allocation created it. An application that hops between SEVERAL contexts of one device is outside what swap
            // mode supports; say so once instead of failing obscurely inside the driver.
            CUcontext cur = nullptr;
            if (d.cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur != ctx_) {
                ctx_warned_ = true;
                LOG_ERROR("device %d: a paged-out buffer is used from context %p, the swap engine lives in %p — several contexts on one "
                          "device are not supported in swap mode", dev_, (void *)cur, (void *)ctx_);
            }
        }
        // order matters for overlap: packs of the victims first (short), then the H2D copies of the incoming rows
        // into staging, THEN the host-side wait for the packs and the VMM remaps, and finally the unpacks
        tr_ = nullptr;
        if (trace_want_ && trace_.size() == trace_want_ && !trace_dumped_) {
            // the driver is gone by the time atexit handlers run: print as soon as the window is full
            trace_dumped_ = true;
            mu_.unlock();
            dump_trace(stderr);
            mu_.lock();
        }
        if (trace_want_ && trace_.size() < trace_want_) {
            if (trace_skip_) trace_skip_--;
            else {
                if (!trace_base_) { d.cuEventCreate(&trace_base_, CU_EVENT_DEFAULT); d.cuEventRecord(trace_base_, s_out_); d.cuEventSynchronize(trace_base_); trace_base_ns_ = mono_ns(); trace_.reserve(trace_want_); }
                trace_.emplace_back();
                tr_ = &trace_.back();
                tr_->t_begin = mono_ns();
            }
        }
        if (pressure_ && cfg_.resident_cap < quota_cap_ && (++pressure_probe_ & 31u) == 0) {
            // probe upwards: the other tenants may have let go; a failed cuMemCreate simply lowers the cap again
            cfg_.resident_cap = std::min(quota_cap_, cfg_.resident_cap + need);
            if (cfg_.resident_cap == quota_cap_) pressure_ = false;
        }
        CUresult r = make_room(need, false);
        if (tr_) tr_->t_packs = mono_ns();
        if (r == CUDA_SUCCESS) r = page_in_stage(missing);
        if (tr_) tr_->t_staged = mono_ns();
        if (r == CUDA_SUCCESS) r = page_out_finish();
        if (tr_) tr_->t_unmapped = mono_ns();
        if (r == CUDA_SUCCESS) r = page_in_finish(missing);
        if (tr_) { tr_->t_end = mono_ns(); tr_ = nullptr; }
        if (r != CUDA_SUCCESS) { page_out_finish(); in_jobs_.clear(); }
        if (r != CUDA_SUCCESS) {
            for (int i = 0; i < n; i++) {
                Side &s = side_[rows[i]];
                if (--s.pins == 0 && (rows_[rows[i]].state & VGPU_ST_RESIDENT)) rows_[rows[i]].state = VGPU_ST_RESIDENT;
                mark_dirty(rows[i]);
            }
            return r;
        }
        for (int r2 : missing) rows_[r2].state = VGPU_ST_RESIDENT | VGPU_ST_PINNED;
    }
    for (int i = 0; i < n; i++) {
        Side &s = side_[rows[i]];
        if (!s.ready) continue;
        if (stream == kHostWait) d.cuEventSynchronize(s.ready);
        if (d.cuEventQuery(s.ready) == CUDA_SUCCESS) { ready_free_.push_back(s.ready); s.ready = nullptr; }
        else d.cuStreamWaitEvent(stream, s.ready, 0);
    }
    if (!missing.empty()) publish_locked();
    return CUDA_SUCCESS;
}

void SwapEngine::note_use(const int *rows, int n, CUstream stream) {
    const DriverTable &d = drv();
    std::lock_guard<std::mutex> g(mu_);
    uint64_t seq = ++use_seq_;
    CUevent ev = use_ring_[seq % use_ring_.size()];
    // the slot's previous owner (seq - ring size) is only forgotten once it is known complete, see use_event()
    if (seq > use_ring_.size()) d.cuEventSynchronize(ev);
    d.cuEventRecord(ev, stream);
    for (int i = 0; i < n; i++) {
        Side &s = side_[rows[i]];
        s.use_seq = seq;
        if (s.pins > 0 && --s.pins == 0 && (rows_[rows[i]].state & VGPU_ST_RESIDENT)) {
            rows_[rows[i]].state = VGPU_ST_RESIDENT;
            mark_dirty(rows[i]);
        }
    }
}

void SwapEngine::publish_locked() {
    vgpu_swap_record_t *r = shared_;
    if (!r) return;
    __atomic_store_n(&r->page_out_bytes, st_.page_out_bytes, __ATOMIC_RELAXED);
    __atomic_store_n(&r->page_in_bytes, st_.page_in_bytes, __ATOMIC_RELAXED);
    __atomic_store_n(&r->evictions, st_.evictions, __ATOMIC_RELAXED);
    __atomic_store_n(&r->faults, st_.faults, __ATOMIC_RELAXED);
    __atomic_store_n(&r->resident_bytes, (uint64_t)resident_mapped_, __ATOMIC_RELAXED);
    __atomic_store_n(&r->live_bytes, (uint64_t)live_bytes_, __ATOMIC_RELAXED);
    __atomic_store_n(&r->host_bytes, (uint64_t)host_used_, __ATOMIC_RELAXED);
}

CUresult SwapEngine::drain() {
    const DriverTable &d = drv();
    std::lock_guard<std::mutex> g(mu_);
    while (evicting_mapped_ != 0) reap_cv_.wait(mu_);
    CUresult r = CUDA_SUCCESS, t;
    for (CUstream s : {s_scan_, s_pack_, s_unpack_, s_out_, s_in_})
        if (s && (t = d.cuStreamSynchronize(s)) != CUDA_SUCCESS) r = t;
    reap_pending_host(true);
    harvest_spans();
    harvest_prof(true);
    return r;
}

SwapStats SwapEngine::stats() {
    std::lock_guard<std::mutex> g(mu_);
    harvest_prof(false);
    harvest_spans();
    SwapStats s = st_;
    s.resident_bytes = resident_mapped_;
    s.live_bytes = live_bytes_;
    s.host_bytes = host_used_;
    s.entries = rows_.size() - free_rows_.size();
    return s;
}

void SwapEngine::trace_mark(std::vector<CUevent> *v, CUstream s) {
    CUevent e = nullptr;
    if (drv().cuEventCreate(&e, CU_EVENT_DEFAULT) != CUDA_SUCCESS) return;
    drv().cuEventRecord(e, s);
    v->push_back(e);
}

void SwapEngine::dump_trace(FILE *f) {
    const DriverTable &d = drv();
    std::lock_guard<std::mutex> g(mu_);
    if (trace_.empty() || !trace_base_) return;
    for (CUstream s : {s_out_, s_in_}) d.cuStreamSynchronize(s);
    auto rel = [&](uint64_t ns) { return (double)(ns - trace_base_ns_) / 1e3; };
    for (size_t i = 0; i < trace_.size(); i++) {
        TraceRec &t = trace_[i];
        std::fprintf(f, "[vgpu-b200 trace] {\"i\": %zu, \"host_us\": {\"begin\": %.0f, \"packs\": %.0f, \"staged\": %.0f, \"unmapped\": %.0f, \"end\": %.0f}",
                     i, rel(t.t_begin), rel(t.t_packs), rel(t.t_staged), rel(t.t_unmapped), rel(t.t_end));
        for (int dir = 0; dir < 2; dir++) {
            std::vector<CUevent> &v = dir ? t.h2d : t.d2h;
            std::fprintf(f, ", \"%s_us\": [", dir ? "h2d" : "d2h");
            for (size_t k = 0; k + 1 < v.size(); k += 2) {
                float a = 0, b = 0;
                d.cuEventElapsedTime(&a, trace_base_, v[k]);
                d.cuEventElapsedTime(&b, trace_base_, v[k + 1]);
                std::fprintf(f, "%s[%.0f, %.0f]", k ? ", " : "", a * 1e3, b * 1e3);
            }
            std::fprintf(f, "]");
        }
        std::fprintf(f, "}\n");
    }
}

std::vector<VgpuEntry> SwapEngine::snapshot_table() {
    std::lock_guard<std::mutex> g(mu_);
    return rows_;
}

}  // namespace vgpu
