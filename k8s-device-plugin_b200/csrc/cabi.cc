// cabi.cc — implementation of include/vgpu.h on top of the core classes. Plain pointers and sizes only.
#include "vgpu.h"

#include <array>
#include <climits>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <mutex>
#include <memory>

#include "driver.h"
#include "kmod.h"
#include "limiter.h"
#include "log.h"
#include "region.h"
#include "runtime.h"
#include "swap.h"

using namespace vgpu;

#define VGPU_API extern "C" __attribute__((visibility("default")))

struct vgpu_region_handle { Region *r; };
struct vgpu_swap { SwapEngine *e; };
struct vgpu_limiter { Limiter *l; };

VGPU_API const char *vgpu_version(void) { return "k8s-device-plugin_b200 0.1 (sm_100a)"; }
VGPU_API uint64_t vgpu_parse_limit(const char *value) { return parse_limit(value); }

// ------------------------------------------------------------------------------------------------ region
VGPU_API int vgpu_region_open(const char *path, int create, const uint64_t *mem_limits, const uint64_t *sm_limits, int priority,
                              vgpu_region_handle_t **out) {
    if (!path || !out) return 1;
    std::string err;
    Region *r = Region::open(path, create != 0, mem_limits, sm_limits, priority, nullptr, 0, &err);
    if (!r) { LOG_WARN("vgpu_region_open(%s): %s", path, err.c_str()); return 1; }
    *out = new vgpu_region_handle{r};
    return 0;
}
VGPU_API void vgpu_region_close(vgpu_region_handle_t *h) { if (h) { delete h->r; delete h; } }
VGPU_API int vgpu_region_snapshot(vgpu_region_handle_t *h, vgpu_region_snapshot_t *o) {
    if (!h || !o) return 1;
    const vgpu_shared_region_t *r = h->r->raw();
    std::memset(o, 0, sizeof *o);
    o->initialized = r->initialized_flag == VGPU_REGION_MAGIC;
    int n = r->proc_num;
    if (n < 0) n = 0;
    if (n > VGPU_MAX_PROCS) n = VGPU_MAX_PROCS;
    o->proc_num = n;
    o->utilization_switch = r->utilization_switch;
    o->recent_kernel = r->recent_kernel;
    o->priority = r->priority;
    o->device_num = r->device_num;
    std::memcpy(o->limit, r->limit, sizeof o->limit);
    std::memcpy(o->sm_limit, r->sm_limit, sizeof o->sm_limit);
    std::memcpy(o->uuids, r->uuids, sizeof o->uuids);
    for (int i = 0; i < n; i++)
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) o->usage_total[d] += r->procs[i].used[d].total;
    return 0;
}
VGPU_API int vgpu_region_proc(vgpu_region_handle_t *h, int index, vgpu_proc_usage_t *o) {
    if (!h || !o || index < 0 || index >= h->r->raw()->proc_num) return 1;
    const vgpu_proc_slot_t &s = h->r->raw()->procs[index];
    o->pid = s.pid; o->hostpid = s.hostpid; o->status = s.status; o->_pad = 0;
    std::memcpy(o->used, s.used, sizeof o->used);
    return 0;
}
VGPU_API int vgpu_region_claim(vgpu_region_handle_t *h, int32_t pid) { return h ? h->r->claim_slot(pid) : -1; }
VGPU_API void vgpu_region_release(vgpu_region_handle_t *h, int32_t pid) { if (h) h->r->release_slot(pid); }
VGPU_API int vgpu_region_try_add(vgpu_region_handle_t *h, int32_t pid, int dev, uint64_t bytes, int type, int enforce) {
    if (!h || dev < 0 || dev >= VGPU_MAX_DEVICES) return 0;
    return h->r->try_add(pid, dev, bytes, type, enforce != 0) ? 1 : 0;
}
VGPU_API void vgpu_region_sub(vgpu_region_handle_t *h, int32_t pid, int dev, uint64_t bytes, int type) {
    if (h && dev >= 0 && dev < VGPU_MAX_DEVICES) h->r->sub(pid, dev, bytes, type);
}
VGPU_API uint64_t vgpu_region_usage(vgpu_region_handle_t *h, int dev) {
    return (h && dev >= 0 && dev < VGPU_MAX_DEVICES) ? h->r->usage(dev) : 0;
}
VGPU_API int vgpu_region_set_feedback(vgpu_region_handle_t *h, int32_t recent_kernel, int32_t utilization_switch) {
    if (!h) return 1;
    // the monitor writes these words without the semaphore (feedback.go:207-251); so do we
    if (recent_kernel != INT32_MIN) h->r->raw()->recent_kernel = recent_kernel;
    if (utilization_switch != INT32_MIN) h->r->raw()->utilization_switch = utilization_switch;
    return 0;
}
VGPU_API int vgpu_region_set_hostpid(vgpu_region_handle_t *h, int32_t pid, int32_t hostpid) {
    if (!h) return 1;
    h->r->set_hostpid(pid, hostpid);
    return 0;
}
VGPU_API int vgpu_region_set_uuid(vgpu_region_handle_t *h, int dev, const char *uuid) {
    if (!h || !uuid || dev < 0 || dev >= VGPU_MAX_DEVICES) return 1;
    vgpu_shared_region_t *r = h->r->raw();
    std::memset(r->uuids[dev], 0, VGPU_UUID_LEN);
    std::strncpy(r->uuids[dev], uuid, VGPU_UUID_LEN - 1);
    if (r->device_num < (uint64_t)dev + 1) r->device_num = (uint64_t)dev + 1;
    return 0;
}
VGPU_API int vgpu_region_swap_counters(vgpu_region_handle_t *h, int dev, vgpu_swap_record_t *out) {
    if (!h || !out || dev < 0 || dev >= VGPU_MAX_DEVICES) return -1;
    return h->r->swap_counters(dev, out) ? 0 : -1;
}
VGPU_API void *vgpu_region_raw(vgpu_region_handle_t *h) { return h ? h->r->raw() : nullptr; }

// ------------------------------------------------------------------------------------------------ monitor feedback
VGPU_API int vgpu_monitor_observe(vgpu_region_handle_t **regions, int n) {
    // utSwitchOn: GPU uuid -> active task count per priority (index 0 = high, 1 = low; feedback.go:214-217)
    std::map<std::string, std::array<int, 2>> active;
    auto key = [](const char *u) { return std::string(u, VGPU_UUID_LEN); };
    auto prio = [](const vgpu_shared_region_t *r) { return r->priority < 0 ? 0 : (r->priority > 1 ? 1 : r->priority); };
    for (int i = 0; i < n; i++) {
        if (!regions[i]) continue;
        vgpu_shared_region_t *r = regions[i]->r->raw();
        if (r->recent_kernel > 0) {
            r->recent_kernel--;
            if (r->recent_kernel > 0)
                for (int d = 0; d < VGPU_MAX_DEVICES; d++) {
                    if (r->uuids[d][0] == 0) continue;          // "Null device condition"
                    active[key(r->uuids[d])][prio(r)]++;
                }
        }
    }
    int changed = 0;
    for (int i = 0; i < n; i++) {
        if (!regions[i]) continue;
        vgpu_shared_region_t *r = regions[i]->r->raw();
        const int p = prio(r);
        bool blocking = false, contended = false;
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) {            // CheckBlocking: decided by the FIRST uuid found in the map
            auto it = active.find(key(r->uuids[d]));
            if (it == active.end()) continue;
            for (int q = 0; q < p; q++) if (it->second[q] > 0) blocking = true;
            break;
        }
        for (int d = 0; d < VGPU_MAX_DEVICES && !contended; d++) {   // CheckPriority: any uuid
            auto it = active.find(key(r->uuids[d]));
            if (it == active.end()) continue;
            for (int q = 0; q < p; q++) if (it->second[q] > 0) contended = true;
            if (it->second[p] > 1) contended = true;
        }
        bool touched = false;
        if (blocking) { if (r->recent_kernel >= 0) { r->recent_kernel = -1; touched = true; } }
        else if (r->recent_kernel < 0) { r->recent_kernel = 0; touched = true; }
        if (contended) { if (r->utilization_switch != 1) { r->utilization_switch = 1; touched = true; } }
        else if (r->utilization_switch != 0) { r->utilization_switch = 0; touched = true; }
        changed += touched;
    }
    return changed;
}

// ------------------------------------------------------------------------------------------------ kernels
VGPU_API int vgpu_pack(const vgpu_seg_t *segs, size_t nseg, void *stream) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    static_assert(sizeof(vgpu_seg_t) == sizeof(PackSegment), "segment layout");
    return launch_pack(k, reinterpret_cast<const PackSegment *>(segs), nseg, static_cast<CUstream>(stream));
}

VGPU_API int vgpu_pack_config(uint32_t tile_bytes, uint32_t stages, uint32_t ctas_per_sm) {
    PackConfig &c = pack_config();
    if (tile_bytes) c.tile_bytes = tile_bytes;
    if (stages) c.stages = stages;
    if (ctas_per_sm) c.ctas_per_sm = ctas_per_sm;
    return CUDA_SUCCESS;
}

VGPU_API int vgpu_victim_scan(uint64_t d_table, uint32_t n, uint64_t need, uint64_t max_touch, void *stream, uint32_t *out_idx,
                              uint32_t out_cap, uint32_t *out_count, uint64_t *freed, int *insufficient) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    // One scanner per context, grown on demand: its scratch (state words, pinned readback) is sized by the row count and
    // allocating it costs far more than the scan itself.
    static std::mutex mu;
    // (leaked on purpose: at static-destruction time the driver may already be torn down)
    static auto *cache = new std::map<const Kernels *, std::pair<uint32_t, std::unique_ptr<VictimScanner>>>();
    std::lock_guard<std::mutex> g(mu);
    auto &slot = (*cache)[k];
    uint32_t want = n ? n : 1;
    CUresult r = CUDA_SUCCESS;
    if (!slot.second || slot.first < want) {
        slot.second.reset(new VictimScanner());
        r = slot.second->init(k, want);
        if (r != CUDA_SUCCESS) { slot.second.reset(); return r; }
        slot.first = want;
    }
    VictimScanner &sc = *slot.second;
    std::vector<uint32_t> v;
    bool ins = false;
    uint64_t fr = 0;
    r = sc.scan(d_table, n, need, max_touch, static_cast<CUstream>(stream), &v, &fr, &ins);
    if (r != CUDA_SUCCESS) return r;
    if (out_count) *out_count = (uint32_t)v.size();
    if (freed) *freed = fr;
    if (insufficient) *insufficient = ins;
    for (size_t i = 0; i < v.size() && i < out_cap; i++) out_idx[i] = v[i];
    return CUDA_SUCCESS;
}

static int wl_launch(CUfunction f, uint64_t nwords_for_grid, void **args, void *stream, int sm) {
    uint64_t blocks = (nwords_for_grid + 255) / 256;
    uint64_t cap = (uint64_t)sm * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return drv().cuLaunchKernel(f, (unsigned)blocks, 1, 1, 256, 1, 1, 0, static_cast<CUstream>(stream), args, nullptr);
}
VGPU_API int vgpu_wl_fill(uint64_t dptr, uint64_t nwords, uint64_t buf_index, void *stream) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    void *a[] = {&dptr, &nwords, &buf_index};
    return wl_launch(k->wl_fill, nwords, a, stream, k->sm_count);
}
VGPU_API int vgpu_wl_touch(uint64_t dptr, uint64_t nwords, void *stream) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    void *a[] = {&dptr, &nwords};
    return wl_launch(k->wl_touch, nwords / 2, a, stream, k->sm_count);
}
VGPU_API int vgpu_wl_touch_indirect(uint64_t d_table, uint32_t nptr, uint64_t nwords, void *stream) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    if (nptr == 0) return CUDA_SUCCESS;
    void *a[] = {&d_table, &nptr, &nwords};
    uint64_t bx = (nwords + 255) / 256, cap = (uint64_t)k->sm_count * 4;
    if (bx > cap) bx = cap;
    if (bx == 0) bx = 1;
    return drv().cuLaunchKernel(k->wl_touch_indirect, (unsigned)bx, nptr < 65535u ? nptr : 65535u, 1, 256, 1, 1, 0, static_cast<CUstream>(stream), a, nullptr);
}
VGPU_API int vgpu_wl_verify(uint64_t dptr, uint64_t nwords, uint64_t buf_index, uint64_t added, uint64_t d_counter, void *stream) {
    const Kernels *k = kernels_for_current_ctx();
    if (!k) return CUDA_ERROR_NOT_INITIALIZED;
    void *a[] = {&dptr, &nwords, &buf_index, &added, &d_counter};
    return wl_launch(k->wl_verify, nwords, a, stream, k->sm_count);
}

// ------------------------------------------------------------------------------------------------ swap engine
static void fill_stats(const SwapStats &s, vgpu_swap_stats_t *o) {
    o->page_out_bytes = s.page_out_bytes; o->page_in_bytes = s.page_in_bytes; o->evictions = s.evictions;
    o->faults = s.faults; o->admissions = s.admissions; o->pack_launches = s.pack_launches;
    o->unpack_launches = s.unpack_launches; o->scan_launches = s.scan_launches; o->scans = s.scans;
    o->resident_bytes = s.resident_bytes; o->live_bytes = s.live_bytes; o->host_bytes = s.host_bytes;
    o->entries = s.entries; o->phys_creates = s.phys_creates; o->phys_reuses = s.phys_reuses;
    o->pack_bytes = s.pack_bytes; o->unpack_bytes = s.unpack_bytes; o->pack_ms = s.pack_ms; o->unpack_ms = s.unpack_ms;
    o->scan_cache_hits = s.scan_cache_hits;
    o->host_admit_ns = s.host_admit_ns; o->host_wait_ns = s.host_wait_ns; o->host_vmm_ns = s.host_vmm_ns;
    o->pager_vmm_ns = s.pager_vmm_ns; o->pager_scan_ns = s.pager_scan_ns; o->pager_packsync_ns = s.pager_packsync_ns;
    o->pager_ring_ns = s.pager_ring_ns; o->pager_busy_ns = s.pager_busy_ns; o->vmm_calls = s.vmm_calls;
    o->pack_span_ms = s.pack_span_ms; o->unpack_span_ms = s.unpack_span_ms;
    o->direct_out_bytes = s.direct_out_bytes; o->direct_in_bytes = s.direct_in_bytes;
    o->prefetch_issued = s.prefetch_issued; o->prefetch_hits = s.prefetch_hits; o->prefetch_wasted = s.prefetch_wasted;
    o->demand_waits = s.demand_waits; o->clean_evictions = s.clean_evictions;
    o->host_slabs = s.host_slabs; o->host_slabs_local = s.host_slabs_local;
    o->pager_unmap_ns = s.pager_unmap_ns; o->pager_setaccess_ns = s.pager_setaccess_ns; o->pager_issue_ns = s.pager_issue_ns;
    o->pager_poll_ns = s.pager_poll_ns; o->pager_lock_ns = s.pager_lock_ns;
    for (int i = 0; i < 5; i++) o->pager_step_ns[i] = s.pager_step_ns[i];
    o->vmm_slow_calls = s.vmm_slow_calls; o->vmm_slow_ns = s.vmm_slow_ns; o->vmm_max_ns = s.vmm_max_ns;
    o->inplace_uses = s.inplace_uses;
}
VGPU_API int vgpu_swap_create(int dev, const vgpu_swap_config_t *cfg, vgpu_swap_t **out) {
    if (!out) return CUDA_ERROR_INVALID_VALUE;
    SwapConfig c = SwapConfig::from_env(cfg ? cfg->resident_cap : 0, cfg ? cfg->virtual_cap : 0);
    if (cfg) {
        if (cfg->host_pool_cap) c.host_pool_cap = cfg->host_pool_cap;
        if (cfg->chunk_bytes) c.chunk_bytes = cfg->chunk_bytes;
        if (cfg->ring_slots) c.ring_slots = (int)cfg->ring_slots;
        c.profile = cfg->profile != 0;
        if (cfg->prefetch_bytes == ~0ull) c.prefetch_bytes = 0;
        else if (cfg->prefetch_bytes) c.prefetch_bytes = cfg->prefetch_bytes;
        if (cfg->copy_bytes) c.copy_bytes = cfg->copy_bytes;
    }
    SwapEngine *e = SwapEngine::create(dev, c);
    if (!e) return CUDA_ERROR_NOT_SUPPORTED;
    *out = new vgpu_swap{e};
    return CUDA_SUCCESS;
}
VGPU_API void vgpu_swap_destroy(vgpu_swap_t *s) { if (s) { delete s->e; delete s; } }
VGPU_API int vgpu_swap_alloc(vgpu_swap_t *s, uint64_t bytes, uint64_t *dptr) {
    if (!s || !dptr) return CUDA_ERROR_INVALID_VALUE;
    CUdeviceptr p = 0;
    CUresult r = s->e->alloc(&p, bytes);
    *dptr = p;
    return r;
}
VGPU_API int vgpu_swap_free(vgpu_swap_t *s, uint64_t dptr) { return s ? s->e->free(dptr) : CUDA_ERROR_INVALID_VALUE; }
static int rows_of(vgpu_swap_t *s, const uint64_t *ptrs, int n, std::vector<int> *rows) {
    for (int i = 0; i < n; i++) {
        int r = s->e->lookup(ptrs[i]);
        if (r < 0) return CUDA_ERROR_INVALID_VALUE;
        bool dup = false;
        for (int x : *rows) dup |= (x == r);
        if (!dup) rows->push_back(r);
    }
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_acquire(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    std::vector<int> rows;
    int rc = rows_of(s, ptrs, n, &rows);
    if (rc) return rc;
    return s->e->ensure_resident(rows.data(), (int)rows.size(), static_cast<CUstream>(stream));
}
VGPU_API int vgpu_swap_release(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    std::vector<int> rows;
    int rc = rows_of(s, ptrs, n, &rows);
    if (rc) return rc;
    s->e->note_use(rows.data(), (int)rows.size(), static_cast<CUstream>(stream));
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_release_ro(vgpu_swap_t *s, const uint64_t *ptrs, int n, void *stream) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    std::vector<int> rows;
    int rc = rows_of(s, ptrs, n, &rows);
    if (rc) return rc;
    s->e->note_use(rows.data(), (int)rows.size(), static_cast<CUstream>(stream), /*writes=*/false);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_advise_read_mostly(vgpu_swap_t *s, uint64_t ptr, int on) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    int row = s->e->lookup(ptr);
    if (row < 0) return CUDA_ERROR_INVALID_VALUE;
    s->e->advise_read_mostly(row, on != 0);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_prefetch(vgpu_swap_t *s, uint64_t ptr, int to_device) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    int row = s->e->lookup(ptr);
    if (row < 0) return CUDA_ERROR_INVALID_VALUE;
    if (to_device) s->e->hint_prefetch(row);
    else s->e->hint_evict(row);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_pin(vgpu_swap_t *s, uint64_t ptr, int on) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    int row = s->e->lookup(ptr);
    if (row < 0) return CUDA_ERROR_INVALID_VALUE;
    return s->e->pin_resident(row, on != 0);
}
VGPU_API int vgpu_swap_stats(vgpu_swap_t *s, vgpu_swap_stats_t *out) {
    if (!s || !out) return CUDA_ERROR_INVALID_VALUE;
    fill_stats(s->e->stats(), out);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_swap_drain(vgpu_swap_t *s) { return s ? s->e->drain() : CUDA_ERROR_INVALID_VALUE; }
VGPU_API int vgpu_swap_set_profile(vgpu_swap_t *s, int on) { if (!s) return CUDA_ERROR_INVALID_VALUE; s->e->set_profile(on != 0); return CUDA_SUCCESS; }
VGPU_API int vgpu_swap_table(vgpu_swap_t *s, vgpu_entry_t *out, uint32_t cap, uint32_t *n) {
    if (!s) return CUDA_ERROR_INVALID_VALUE;
    static_assert(sizeof(vgpu_entry_t) == sizeof(VgpuEntry), "table row layout");
    std::vector<VgpuEntry> t = s->e->snapshot_table();
    if (n) *n = (uint32_t)t.size();
    for (uint32_t i = 0; i < t.size() && i < cap; i++) std::memcpy(&out[i], &t[i], sizeof(VgpuEntry));
    return CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ limiter
static void fill_lim(const LimiterStats &s, vgpu_limiter_stats_t *o) {
    o->launches = s.launches; o->stamps = s.stamps; o->groups = s.groups; o->busy_ns = s.busy_ns;
    o->throttle_ns = s.throttle_ns; o->wall_ns = s.wall_ns; o->limit_percent = s.limit_percent; o->_pad = 0;
}
VGPU_API int vgpu_limiter_create(int percent, vgpu_limiter_t **out) {
    if (!out) return CUDA_ERROR_INVALID_VALUE;
    *out = new vgpu_limiter{new Limiter(percent, nullptr, 1)};
    return CUDA_SUCCESS;
}
VGPU_API void vgpu_limiter_destroy(vgpu_limiter_t *l) { if (l) { delete l->l; delete l; } }
VGPU_API void vgpu_limiter_before_launch(vgpu_limiter_t *l, void *stream) { if (l) l->l->before_launch(static_cast<CUstream>(stream)); }
VGPU_API void vgpu_limiter_after_launch(vgpu_limiter_t *l, void *stream) { if (l) l->l->after_launch(static_cast<CUstream>(stream)); }
VGPU_API int vgpu_limiter_stats(vgpu_limiter_t *l, vgpu_limiter_stats_t *out) {
    if (!l || !out) return CUDA_ERROR_INVALID_VALUE;
    fill_lim(l->l->stats(), out);
    return CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ runtime introspection
VGPU_API int vgpu_runtime_swap_stats(int dev, vgpu_swap_stats_t *out) {
    SwapEngine *e = Runtime::get().swap(dev);
    if (!e || !out) return CUDA_ERROR_NOT_INITIALIZED;
    fill_stats(e->stats(), out);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_runtime_limiter_stats(vgpu_limiter_stats_t *out) {
    Limiter *l = Runtime::get().limiter();
    if (!l || !out) return CUDA_ERROR_NOT_INITIALIZED;
    fill_lim(l->stats(), out);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_runtime_set_swap_profile(int dev, int on) {
    SwapEngine *e = Runtime::get().swap(dev);
    if (!e) return CUDA_ERROR_NOT_INITIALIZED;
    e->set_profile(on != 0);
    return CUDA_SUCCESS;
}
VGPU_API int vgpu_runtime_swap_pin(uint64_t dptr, int on) {
    CUdevice dev = -1;
    if (drv().cuCtxGetDevice(&dev) != CUDA_SUCCESS) return CUDA_ERROR_INVALID_CONTEXT;
    SwapEngine *e = Runtime::get().swap((int)dev);
    if (!e) return CUDA_ERROR_NOT_INITIALIZED;
    int row = e->lookup(dptr);
    if (row < 0) return CUDA_ERROR_INVALID_VALUE;
    return e->pin_resident(row, on != 0);
}
VGPU_API uint64_t vgpu_runtime_context_size(void) { return Runtime::get().context_size(); }
VGPU_API int vgpu_runtime_check_memory_type(uint64_t dptr) { return Runtime::get().check_memory_type(dptr); }
