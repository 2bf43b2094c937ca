// runtime.cc — see runtime.h.
#include <algorithm>
#include "runtime.h"
#include <thread>
#include <csignal>

#include <fcntl.h>
#include <pthread.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <unordered_map>
#include <vector>

#include "driver.h"
#include "limiter.h"
#include "log.h"
#include "swap.h"

namespace vgpu {

static bool env_true(const char *name) {
    const char *e = std::getenv(name);
    return e && (!strcasecmp(e, "true") || !std::strcmp(e, "1"));
}
// Return codes on a quota breach / unknown pointer follow the reference bit for bit by default ((CUresult)-1 from
// add_chunk@0x4005d and remove_chunk@0x409f0); VGPU_STRICT_CUDA_ERRORS=1 switches to CUDA-conformant codes.
// VGPU_REFERENCE_COVERAGE=1: intercept exactly what the reference intercepts (async/pool allocations, cuMemCreate and
// graph launches forwarded untouched)
static bool reference_coverage() {
    static bool on = [] { const char *e = std::getenv("VGPU_REFERENCE_COVERAGE"); return e && *e && *e != '0'; }();
    return on;
}
static bool strict_errors() {
    static bool v = env_true("VGPU_STRICT_CUDA_ERRORS");
    return v;
}
static const CUresult kQuotaBreachAlloc = static_cast<CUresult>(-1);

// load_env_from_file@0x415a4 (called from nvml_preInit@0x24029 with "/overrideEnv"): every line KEY=VALUE of that file,
// trailing newline stripped, split at the FIRST '=', is setenv'd over the process environment before anything is read
// from it — the operator's way to change a running container's limits. (VGPU_OVERRIDE_ENV_FILE moves the path; tests.)
static void load_env_from_file() {
    const char *path = std::getenv("VGPU_OVERRIDE_ENV_FILE");
    FILE *f = std::fopen(path && *path ? path : "/overrideEnv", "r");
    if (!f) return;
    char line[10000];
    while (std::fgets(line, sizeof line, f)) {
        size_t n = std::strlen(line);
        if (n && line[n - 1] == '\n') line[n - 1] = 0;
        char *eq = std::strchr(line, '=');
        if (!eq) continue;
        *eq = 0;
        setenv(line, eq + 1, 1);
        LOG_INFO("SET %s to %s", line, eq + 1);
    }
    std::fclose(f);
}

Config Config::from_env() {
    Config c;
    c.oversubscribe = env_true("CUDA_OVERSUBSCRIBE");
    if (const char *p = std::getenv("GPU_CORE_UTILIZATION_POLICY")) {
        if (!strcasecmp(p, "force")) c.util_policy = 1;
        else if (!strcasecmp(p, "disable")) c.util_policy = 2;
    }
    c.active_oom_killer = env_true("ACTIVE_OOM_KILLER");
    if (const char *p = std::getenv("CUDA_TASK_PRIORITY")) c.priority = std::atoi(p);
    const char *rp = std::getenv("CUDA_DEVICE_MEMORY_SHARED_CACHE");
    c.region_path = rp && *rp ? rp : "/tmp/cudevshr.cache";
    for (int d = 0; d < VGPU_MAX_DEVICES; d++) {
        c.mem_limit[d] = limit_from_env("CUDA_DEVICE_MEMORY_LIMIT", d);
        uint64_t sm = limit_from_env("CUDA_DEVICE_SM_LIMIT", d);
        c.sm_limit[d] = sm ? sm : 100;  // do_init_device_sm_limits@0x41946: default 100
        c.virtual_limit[d] = limit_from_env("CUDA_DEVICE_MEMORY_VIRTUAL_LIMIT", d);
    }
    if (const char *p = std::getenv("VGPU_SWAP_LIMIT_MODE")) c.limit_is_virtual = !strcasecmp(p, "virtual");
    if (c.limit_is_virtual)
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) c.virtual_limit[d] = c.mem_limit[d];
    return c;
}

Runtime &Runtime::get() {
    static Runtime *r = new Runtime();  // intentionally leaked: hooks may run during process teardown
    return *r;
}

static void atexit_trampoline() { Runtime::get().on_exit(); }
static void atfork_child_trampoline() { Runtime::get().on_fork_child(); }

bool Runtime::ensure_initialized() {
    if (inited_.load(std::memory_order_acquire)) return region_ != nullptr;
    std::lock_guard<std::mutex> g(init_mu_);
    if (inited_.load(std::memory_order_relaxed)) return region_ != nullptr;
    load_env_from_file();
    cfg_ = Config::from_env();
    pid_ = getpid();
    char uuids[VGPU_MAX_DEVICES][VGPU_UUID_LEN];
    std::memset(uuids, 0, sizeof uuids);
    int ndev = 0;
    if (nvml_ready()) {  // put_device_info (multiprocess_memory_limit.c:L150)
        unsigned cnt = 0;
        if (nvml().nvmlDeviceGetCount_v2(&cnt) == NVML_SUCCESS) {
            for (unsigned i = 0; i < cnt && i < VGPU_MAX_DEVICES; i++) {
                nvmlDevice_t h;
                if (nvml().nvmlDeviceGetHandleByIndex_v2(i, &h) == NVML_SUCCESS) nvml().nvmlDeviceGetUUID(h, uuids[i], VGPU_UUID_LEN);
            }
            ndev = (int)(cnt < VGPU_MAX_DEVICES ? cnt : VGPU_MAX_DEVICES);
        }
    }
    std::string err;
    Region *R = Region::open(cfg_.region_path.c_str(), true, cfg_.mem_limit, cfg_.sm_limit, cfg_.priority, uuids, ndev, &err);
    if (!R) {
        // Without the region nothing can be accounted. If the container HAS a gpumem quota, running on without it would
        // silently hand the process the whole GPU (a runAsNonRoot container that cannot create its cache file):
        // fail closed — device allocations are refused — unless the operator opts out with VGPU_FAIL_OPEN=1.
        bool limited = false;
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) limited |= cfg_.mem_limit[d] != 0;
        const char *open_env = std::getenv("VGPU_FAIL_OPEN");
        fail_closed_ = limited && !(open_env && *open_env && *open_env != '0');
        LOG_ERROR("shared region %s unavailable (%s): %s", cfg_.region_path.c_str(), err.c_str(),
                  fail_closed_ ? "a gpumem quota is configured and cannot be enforced, device allocations are REFUSED (VGPU_FAIL_OPEN=1 to run unenforced)"
                               : "limits are NOT enforced in this process");
    } else {
        region_.reset(R);
        if (region_->claim_slot(pid_) < 0) LOG_ERROR("no free process slot in %s", cfg_.region_path.c_str());
        region_->sweep_dead(pid_);          // init_proc_slot_withlock@0x43d89 -> clear_proc_slot_nolock: a joining process drops dead siblings' slots
        std::atexit(atexit_trampoline);
        pthread_atfork(nullptr, nullptr, atfork_child_trampoline);
    }
    inited_.store(true, std::memory_order_release);
    return region_ != nullptr;
}

void Runtime::on_exit() {
    // the pager threads must not be inside the driver when it tears itself down behind this handler
    for (int d = 0; d < VGPU_MAX_DEVICES; d++) if (swap_[d]) swap_[d]->stop_pager();
    if (std::getenv("VGPU_PRINT_STATS")) {
        if (limiter_) {
            LimiterStats s = limiter_->stats();
            std::fprintf(stderr, "[vgpu-b200 stats] limiter: limit=%d%% active=%d launches=%lu stamps=%lu groups=%lu busy_ms=%.1f throttle_ms=%.1f wall_ms=%.1f\n",
                         s.limit_percent, (int)limiter_->active(), (unsigned long)s.launches, (unsigned long)s.stamps, (unsigned long)s.groups,
                         s.busy_ns / 1e6, s.throttle_ns / 1e6, s.wall_ns / 1e6);
        } else {
            std::fprintf(stderr, "[vgpu-b200 stats] limiter: not created (cuInit was never intercepted)\n");
        }
        for (int d = 0; d < VGPU_MAX_DEVICES; d++)
            if (swap_[d]) {
                SwapStats s = swap_[d]->stats();
                std::fprintf(stderr, "[vgpu-b200 stats] swap dev %d ms: app admit=%.1f wait=%.1f vmm=%.1f | pager busy=%.1f vmm=%.1f (%lu calls) scan=%.1f packsync=%.1f ringwait=%.1f slabs=%lu local=%lu\n", d,
                             s.host_admit_ns / 1e6, s.host_wait_ns / 1e6, s.host_vmm_ns / 1e6, s.pager_busy_ns / 1e6, s.pager_vmm_ns / 1e6, (unsigned long)s.vmm_calls,
                             s.pager_scan_ns / 1e6, s.pager_packsync_ns / 1e6, s.pager_ring_ns / 1e6, (unsigned long)s.host_slabs, (unsigned long)s.host_slabs_local);
                std::fprintf(stderr, "[vgpu-b200 stats] swap dev %d pager vmm ms: unmap=%.1f setaccess=%.1f map=%.1f create=%.1f | issue=%.1f | steps zombies=%.1f reap=%.1f demand=%.1f prefetch=%.1f evict=%.1f\n", d, s.pager_unmap_ns / 1e6,
                             s.pager_setaccess_ns / 1e6, s.pager_map_ns / 1e6, s.pager_create_ns / 1e6, s.pager_issue_ns / 1e6, s.pager_step_ns[0] / 1e6, s.pager_step_ns[1] / 1e6,
                             s.pager_step_ns[2] / 1e6, s.pager_step_ns[3] / 1e6, s.pager_step_ns[4] / 1e6);
                std::fprintf(stderr, "[vgpu-b200 stats] swap dev %d: prefetch issued=%lu hits=%lu wasted=%lu demand_waits=%lu clean_evictions=%lu inplace_uses=%lu direct in=%lu out=%lu\n", d,
                             (unsigned long)s.prefetch_issued, (unsigned long)s.prefetch_hits, (unsigned long)s.prefetch_wasted, (unsigned long)s.demand_waits,
                             (unsigned long)s.clean_evictions, (unsigned long)s.inplace_uses, (unsigned long)s.direct_in_bytes, (unsigned long)s.direct_out_bytes);
                std::fprintf(stderr, "[vgpu-b200 stats] swap dev %d: in=%lu out=%lu faults=%lu evictions=%lu scans=%lu cache_hits=%lu creates=%lu reuses=%lu\n", d,
                             (unsigned long)s.page_in_bytes, (unsigned long)s.page_out_bytes, (unsigned long)s.faults, (unsigned long)s.evictions,
                             (unsigned long)s.scans, (unsigned long)s.scan_cache_hits, (unsigned long)s.phys_creates, (unsigned long)s.phys_reuses);
            }
    }
    if (region_) region_->release_slot(pid_);
}

void Runtime::on_fork_child() {
    // child_reinit_flag@0x43f54: the child is a new pid with no CUDA state; give it its own slot
    pid_ = getpid();
    {
        std::lock_guard<std::mutex> g(table_mu_);
        table_.clear();
        phys_.clear();
        vmaps_.clear();
    }
    for (auto &c : ctx_charged_) c = false;
    // engines and the limiter own streams, events and threads of the PARENT's contexts: drop them without running
    // destructors (CUDA state does not survive fork); the child builds its own on first use
    for (auto &e : swap_) (void)e.release();
    (void)limiter_.release();
    post_inited_.store(false, std::memory_order_release);
    if (region_) { region_->claim_slot(pid_); region_->sweep_dead(pid_); }
}

int Runtime::current_device() {
    CUdevice d = -1;
    if (drv().cuCtxGetDevice(&d) != CUDA_SUCCESS) return -1;
    return (int)d;
}

// ------------------------------------------------------------------------------------------------ init
static bool unified_lock() {
    // try_lock_unified_lock@0x1612c: the lock IS the existence of /tmp/vgpulock/lock (open O_CREAT|O_EXCL, remove to
    // unlock); the directory is bind-mounted from the host so every container on the node serialises here.
    const char *path = "/tmp/vgpulock/lock";
    for (int i = 0; i < 200; i++) {
        int fd = ::open(path, O_CREAT | O_EXCL, 0700);
        if (fd >= 0) { ::close(fd); return true; }
        if (errno == ENOENT) return false;  // directory not mounted: nothing to serialise against
        usleep(100 * 1000);
    }
    LOG_MSG("unified lock stale for 20 s, removing");
    ::remove(path);
    int fd = ::open(path, O_CREAT | O_EXCL, 0700);
    if (fd >= 0) ::close(fd);
    return fd >= 0;
}
static void unified_unlock() { ::remove("/tmp/vgpulock/lock"); }

// CUDA ordinal -> NVML handle by UUID (the reference keeps cuda_to_nvml_map@0x62c80, built by
// map_cuda_visible_devices@0x17e5a): CUDA_VISIBLE_DEVICES may reorder or hide devices, NVML enumerates all of them.
static bool nvml_handle_for_cuda(int cuda_dev, nvmlDevice_t *h) {
    const NvmlTable &n = nvml();
    const DriverTable &d = drv();
    CUuuid u;
    if (n.nvmlDeviceGetHandleByUUID && d.cuDeviceGetUuid_v2 && d.cuDeviceGetUuid_v2(&u, cuda_dev) == CUDA_SUCCESS) {
        const unsigned char *b = reinterpret_cast<const unsigned char *>(u.bytes);
        char s[64];
        std::snprintf(s, sizeof s, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2], b[3], b[4], b[5],
                      b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
        if (n.nvmlDeviceGetHandleByUUID(s, h) == NVML_SUCCESS) return true;
    }
    return n.nvmlDeviceGetHandleByIndex_v2 && n.nvmlDeviceGetHandleByIndex_v2((unsigned)cuda_dev, h) == NVML_SUCCESS;
}

void Runtime::measure_context_size() {
    // set_task_pid@0x16a7f (utils.c:L135-202): NVML's compute-process list before and after creating the primary
    // context; the new entry is this process as the HOST sees it (pid namespace) and its usedGpuMemory is what a
    // bare context costs on this GPU/driver.
    if (std::getenv("VGPU_SKIP_CONTEXT_MEASURE") || !nvml_ready()) { LOG_WARN("SET_TASK_PID FAILED."); return; }
    const NvmlTable &n = nvml();
    const DriverTable &d = drv();
    nvmlDevice_t h;
    if (!nvml_handle_for_cuda(0, &h)) { LOG_WARN("SET_TASK_PID FAILED."); return; }
    bool locked = unified_lock();
    std::vector<nvmlProcessInfo_t> before(1024), after(1024);
    unsigned nb = (unsigned)before.size(), na = (unsigned)after.size();
    nvmlReturn_t r1 = n.nvmlDeviceGetComputeRunningProcesses_v3(h, &nb, before.data());
    if (r1 != NVML_SUCCESS) nb = 0;
    CUcontext ctx = nullptr;
    if (d.cuDevicePrimaryCtxRetain(&ctx, 0) == CUDA_SUCCESS) {
        nvmlReturn_t r2 = n.nvmlDeviceGetComputeRunningProcesses_v3(h, &na, after.data());
        if (r2 != NVML_SUCCESS) na = 0;
        std::set<unsigned> old;
        for (unsigned i = 0; i < nb; i++) old.insert(before[i].pid);
        int pick = -1;
        for (unsigned i = 0; i < na; i++) if (after[i].pid == (unsigned)pid_) pick = (int)i;         // same pid namespace
        if (pick < 0) for (unsigned i = 0; i < na; i++) if (!old.count(after[i].pid)) { pick = (int)i; break; }
        if (pick >= 0) {
            context_size_ = after[pick].usedGpuMemory;
            pid_found_ = true;
            if (region_) region_->set_hostpid(pid_, (int32_t)after[pick].pid);
            LOG_INFO("hostPid=%u Primary Context Size==%lu", after[pick].pid, (unsigned long)context_size_);
        } else {
            LOG_WARN("host pid is error!");
        }
        d.cuDevicePrimaryCtxRelease_v2(0);
    }
    if (locked) unified_unlock();
    if (!pid_found_) LOG_WARN("SET_TASK_PID FAILED.");
}

void Runtime::post_init(bool late) {
    bool expected = false;
    if (!post_inited_.compare_exchange_strong(expected, true)) return;
    // late = first intercepted call arrived without this library ever seeing cuInit (a runtime that resolved cuInit
    // through a route that is not interposed): the driver is initialised and a context may already hold
    // allocations, so the bare-context measurement would be wrong — skip it, keep everything else.
    if (!late) measure_context_size();
    else LOG_WARN("cuInit was not intercepted; initialising on first use (context size not measured)");
    int pct = (int)(region_ ? region_->sm_limit(0) : cfg_.sm_limit[0]);
    limiter_.reset(new Limiter(pct, region_ ? region_->raw() : nullptr, cfg_.util_policy));
    start_memory_monitor(pct);
    LOG_MSG("Initialized: oversubscribe=%d mem_limit0=%lu sm_limit0=%d", (int)cfg_.oversubscribe,
            (unsigned long)(region_ ? region_->limit(0) : 0), pct);
}

// The reference's NVML-side safety net. Its utilization_watcher thread (only running when 0 < sm_limit < 100,
// init_utilization_watcher@0x46937) calls get_used_gpu_utilization@0x46220 every 120 ms, which feeds every process NVML
// lists on a device into set_gpu_device_memory_monitor@0x42301: the slot whose hostpid matches gets
// monitorused[dev] = usedGpuMemory; under MEMORY_OVERRIDE=1 also offset = used - total, total = used; and when
// (double)used > (double)limit * 1.1 (.rodata@0x56220) with the active OOM killer enabled (ACTIVE_OOM_KILLER, default on:
// set_active_oom_killer@0x43ffa) every process of the container is killed (active_oom_killer@0x41e55). It catches what
// the intercept cannot see. Same thread condition, period, arithmetic and log line here.
static bool oom_killer_enabled() {
    const char *e = std::getenv("ACTIVE_OOM_KILLER");
    if (!e) return true;
    if (!std::strcmp(e, "false") || !std::strcmp(e, "0")) return false;
    return true;
}

void Runtime::start_memory_monitor(int sm_limit_percent) {
    if (sm_limit_percent <= 0 || sm_limit_percent >= 100 || !region_ || !nvml_ready()) return;
    if (cfg_.util_policy == 2) return;   // GPU_CORE_UTILIZATION_POLICY=disable: no watcher thread in the reference either
    bool any = false;
    for (int d = 0; d < VGPU_MAX_DEVICES; d++) any |= region_->limit(d) != 0;
    if (!any) return;
    const bool kill_enabled = oom_killer_enabled();
    const bool memory_override = env_true("MEMORY_OVERRIDE") || (std::getenv("MEMORY_OVERRIDE") && std::atoi(std::getenv("MEMORY_OVERRIDE")) == 1);
    std::thread([this, kill_enabled, memory_override] {
        const NvmlTable &n = nvml();
        if (!n.nvmlDeviceGetCount_v2 || !n.nvmlDeviceGetHandleByIndex_v2 || !n.nvmlDeviceGetComputeRunningProcesses_v3) return;
        for (;;) {
            struct timespec ts = {0, 120000000};   // g_wait .rodata@0x56240
            nanosleep(&ts, nullptr);
            int cnt = 0;
            if (!drv().cuDeviceGetCount || drv().cuDeviceGetCount(&cnt) != CUDA_SUCCESS) continue;
            for (int d = 0; d < cnt && d < VGPU_MAX_DEVICES; d++) {      // d = CUDA ordinal = region lane
                uint64_t lim = region_->limit(d);
                nvmlDevice_t h;
                if (!nvml_handle_for_cuda(d, &h)) continue;
                nvmlProcessInfo_t infos[64];
                unsigned np = 64;
                if (n.nvmlDeviceGetComputeRunningProcesses_v3(h, &np, infos) != NVML_SUCCESS) continue;
                for (unsigned k = 0; k < np; k++) {
                    uint64_t used = infos[k].usedGpuMemory;
                    bool over = false;
                    region_->lock();
                    vgpu_shared_region_t *r = region_->raw();
                    for (int i = 0; i < r->proc_num; i++) {
                        if (r->procs[i].hostpid != (int32_t)infos[k].pid) continue;
                        r->procs[i].monitorused[d] = used;
                        if (memory_override) { r->procs[i].used[d].offset = used - r->procs[i].used[d].total; r->procs[i].used[d].total = used; }
                        if (lim && (double)used > (double)lim * 1.1) over = true;
                    }
                    region_->unlock();
                    if (over && kill_enabled) {
                        LOG_ERROR("device OOM encountered: usage=%lu limit=%lu", (unsigned long)used, (unsigned long)lim);
                        std::vector<int32_t> pids;
                        region_->lock();
                        for (int i = 0; i < r->proc_num; i++) pids.push_back(r->procs[i].pid);
                        region_->unlock();
                        for (int32_t p : pids) if (p > 0 && p != pid_) kill(p, SIGKILL);
                        kill(pid_, SIGKILL);
                    }
                }
            }
        }
    }).detach();
}

CUresult Runtime::init(unsigned flags) {
    ensure_initialized();
    if (!drv().loaded) return CUDA_ERROR_NOT_INITIALIZED;
    CUresult r = drv().cuInit(flags);
    if (r == CUDA_SUCCESS) post_init(false);
    return r;
}

void Runtime::wait_running() {
    // wait_status_self(1)@0x4555e guards every reference wrapper; the status word only ever leaves RUNNING through
    // the reference's dormant SIGUSR2 suspend protocol (nothing calls suspend_all), so there is nothing to wait for.
}

// ------------------------------------------------------------------------------------------------ allocation table
bool Runtime::track(CUdeviceptr base, size_t size, int dev, AllocKind kind) {
    table_[base] = Alloc{size, dev, kind};
    return true;
}

bool Runtime::reference_coverage_mode() const { return reference_coverage(); }

int Runtime::check_memory_type(CUdeviceptr p) {
    std::lock_guard<std::mutex> g(table_mu_);
    auto it = table_.upper_bound(p);
    if (it == table_.begin()) return 1;
    --it;
    return (p >= it->first && p <= it->first + it->second.size) ? 2 : 1;  // inclusive end, like check_memory_type@0x407f2
}

size_t Runtime::table_size() {
    std::lock_guard<std::mutex> g(table_mu_);
    return table_.size();
}

SwapEngine *Runtime::swap(int dev) {
    if (dev < 0 || dev >= VGPU_MAX_DEVICES) return nullptr;
    return swap_[dev].get();
}

// what the engine may keep resident: the quota minus everything that is resident for life (context, non-swappable
// allocations) minus the engine's own staging rings
// resident-for-life bytes = accounted usage minus the live swappable bytes. The lanes use the reference's wrapping u64
// arithmetic (a cross-device free can drive one below zero, see mem_free), so the difference is read as signed.
static uint64_t fixed_bytes(uint64_t usage, uint64_t live) {
    int64_t f = (int64_t)(usage - live);
    return f > 0 ? (uint64_t)f : 0;
}

static uint64_t room_for_engine(uint64_t lim, uint64_t fixed, const SwapEngine *e) {
    uint64_t taken = fixed + (e ? e->device_overhead() : 0);
    return lim > taken ? lim - taken : 0;
}

bool Runtime::charge(int dev, size_t bytes) {
    if (!cfg_.oversubscribe) return region_->try_add(pid_, dev, bytes, VGPU_MEM_BUFFER, true);   // oom_check + add, reference semantics
    if (cfg_.limit_is_virtual) {
        // reference meaning of the limit: a hard cap on live bytes and nothing else. What may be RESIDENT is bounded by
        // the device, not by the limit (the engine sizes itself from the device's free memory and backs off under
        // physical pressure), so nothing is taken out of its room here.
        return region_->try_add(pid_, dev, bytes, VGPU_MEM_BUFFER, true);
    }
    // swap mode: the quota bounds RESIDENT bytes. Non-swappable allocations are resident for life, so they are checked
    // against the quota net of what the swap engine can page out, and they shrink the engine's resident budget.
    uint64_t lim = region_->limit(dev);
    SwapEngine *e = swap(dev);
    if (lim) {
        // check and add in ONE critical section of the region (two processes of a container could both pass a
        // check-then-add): non-swappable bytes = accounted usage - live swappable bytes of EVERY engine of the container on
        // this device (sibling processes publish theirs in the region)
        uint64_t fixed_after = 0;
        if (!region_->try_add_fixed(pid_, dev, bytes, e ? e->live_bytes() : 0, &fixed_after)) {
            LOG_ERROR("Device %d OOM %lu / %lu (non-swappable)", dev, (unsigned long)fixed_after, (unsigned long)lim);
            return false;
        }
        if (e) e->set_resident_cap(room_for_engine(lim, fixed_after, e));
        return true;
    }
    region_->add(pid_, dev, bytes, VGPU_MEM_BUFFER);
    return true;
}

void Runtime::uncharge(int dev, size_t bytes) {
    region_->sub(pid_, dev, bytes, VGPU_MEM_BUFFER);
    if (!cfg_.oversubscribe || cfg_.limit_is_virtual) return;
    uint64_t lim = region_->limit(dev);
    SwapEngine *e = swap(dev);
    if (lim && e) {
        uint64_t u = region_->usage(dev), live = e->live_bytes() + region_->swap_live(dev, pid_);
        uint64_t fixed = fixed_bytes(u, live);
        e->set_resident_cap(room_for_engine(lim, fixed, e));
    }
}

CUresult Runtime::swap_alloc(CUdeviceptr *dptr, size_t bytes, int dev) {
    SwapEngine *e;
    {
        std::lock_guard<std::mutex> g(swap_mu_);
        if (!swap_[dev]) {
            uint64_t lim = region_ ? region_->limit(dev) : 0;
            uint64_t fixed = region_ ? fixed_bytes(region_->usage(dev), region_->swap_live(dev, pid_)) : 0;
            uint64_t cap = lim > fixed ? lim - fixed : 0;
            // virtual mode: the region check is the cap on live bytes; residency is bounded by the device (cap 0 = size
            // from the device's free memory)
            SwapConfig sc = SwapConfig::from_env(lim && !cfg_.limit_is_virtual ? cap : 0, cfg_.limit_is_virtual ? 0 : cfg_.virtual_limit[dev]);
            if (lim && !cfg_.limit_is_virtual) {
                uint64_t overhead = 2ull * sc.ring_slots * sc.chunk_bytes;     // == SwapEngine::device_overhead()
                sc.resident_cap = cap > overhead ? cap - overhead : 0;
                if (sc.resident_cap < (32ull << 20)) { LOG_ERROR("gpumem quota %lu leaves no room for swappable memory", (unsigned long)lim); return CUDA_ERROR_OUT_OF_MEMORY; }
            }
            swap_[dev].reset(SwapEngine::create(dev, sc));
            if (!swap_[dev]) { LOG_ERROR("swap engine unavailable on device %d", dev); return CUDA_ERROR_NOT_SUPPORTED; }
            if (region_) swap_[dev]->set_shared_record(region_->swap_record(pid_, dev));   // counters for the node monitor
            if (region_ && lim && !cfg_.limit_is_virtual) {
                // the resident quota is the CONTAINER's: several processes on this device share it through the region
                // (reference: get_gpu_memory_usage@0x420bd sums the usage of all process slots; UVM arbitrates residency)
                SwapEngine *eng = swap_[dev].get();
                Region *reg = region_.get();
                int32_t pid = pid_;
                eng->set_budget_fn([reg, eng, pid, dev](uint64_t want_total, uint64_t live, bool *granted, int *engines, uint64_t *share) {
                    return reg->swap_reserve(pid, dev, want_total, live, eng->device_overhead(), granted, engines, share);
                });
            }
        }
        e = swap_[dev].get();
    }
    if (cfg_.limit_is_virtual && region_) {
        // oom_check(dev, size) of the reference: usage + size > limit refuses, oversubscribed or not
        if (!region_->try_add(pid_, dev, bytes, VGPU_MEM_BUFFER, true)) return CUDA_ERROR_OUT_OF_MEMORY;
        CUresult r = e->alloc(dptr, bytes);
        if (r != CUDA_SUCCESS) { region_->sub(pid_, dev, bytes, VGPU_MEM_BUFFER); return r; }
        track(*dptr, bytes, dev, AllocKind::Swap);
        return CUDA_SUCCESS;
    }
    CUresult r = e->alloc(dptr, bytes);
    if (r != CUDA_SUCCESS) return r;
    if (region_) region_->add(pid_, dev, bytes, VGPU_MEM_BUFFER);
    track(*dptr, bytes, dev, AllocKind::Swap);
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_alloc(CUdeviceptr *dptr, size_t bytes) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!region_) return fail_closed_ ? CUDA_ERROR_OUT_OF_MEMORY : d.cuMemAlloc_v2(dptr, bytes);
    int dev = current_device();
    if (dev < 0) return d.cuMemAlloc_v2(dptr, bytes);  // no context: let the driver report it
    std::lock_guard<std::mutex> g(table_mu_);          // allocate_raw@0x40a10 holds the allocator mutex across add_chunk
    if (cfg_.oversubscribe && bytes > kIpcSize) {
        // the reference's swap switch: cuMemoryAllocate@0x315da allocmode 0 (there: cuMemAllocManaged + UVM)
        CUresult r = swap_alloc(dptr, bytes, dev);
        if (r == CUDA_ERROR_OUT_OF_MEMORY && !strict_errors()) return kQuotaBreachAlloc;
        return r;
    }
    if (!charge(dev, bytes)) return strict_errors() ? CUDA_ERROR_OUT_OF_MEMORY : kQuotaBreachAlloc;   // add_chunk@0x4005d returns -1
    CUresult r = d.cuMemAlloc_v2(dptr, bytes);
    if (r != CUDA_SUCCESS) {
        uncharge(dev, bytes);
        LOG_ERROR("cuMemoryAllocate failed res=%d", (int)r);
        return r;
    }
    track(*dptr, bytes, dev, AllocKind::Device);
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_alloc_managed(CUdeviceptr *dptr, size_t bytes, unsigned flags) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!region_) return fail_closed_ ? CUDA_ERROR_OUT_OF_MEMORY : d.cuMemAllocManaged(dptr, bytes, flags);
    int dev = current_device();
    if (dev < 0) return d.cuMemAllocManaged(dptr, bytes, flags);
    if (!charge(dev, bytes)) return CUDA_ERROR_OUT_OF_MEMORY;  // @0x31eab
    CUresult r = d.cuMemAllocManaged(dptr, bytes, flags);
    if (r != CUDA_SUCCESS) { uncharge(dev, bytes); return r; }
    std::lock_guard<std::mutex> g(table_mu_);   // add_chunk_only@0x404a5
    track(*dptr, bytes, dev, AllocKind::Managed);
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_alloc_pitch(CUdeviceptr *dptr, size_t *pitch, size_t width, size_t height, unsigned elem) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!region_ && fail_closed_) return CUDA_ERROR_OUT_OF_MEMORY;
    if (!region_ || elem == 0) return d.cuMemAllocPitch_v2(dptr, pitch, width, height, elem);
    int dev = current_device();
    if (dev < 0) return d.cuMemAllocPitch_v2(dptr, pitch, width, height, elem);
    // cuMemAllocPitch_v2@0x3206b: guess_pitch = ((W-1)/elem + 1) * elem ; bytesize = guess_pitch * H — the GUESS is
    // what gets charged, not the pitch the driver picks
    size_t guess = ((width - 1) / elem + 1) * (size_t)elem;
    size_t bytes = guess * height;
    if (!charge(dev, bytes)) return CUDA_ERROR_OUT_OF_MEMORY;  // @0x321ef
    CUresult r = d.cuMemAllocPitch_v2(dptr, pitch, width, height, elem);
    if (r != CUDA_SUCCESS) { uncharge(dev, bytes); return r; }
    std::lock_guard<std::mutex> g(table_mu_);
    track(*dptr, bytes, dev, AllocKind::Pitch);
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_free(CUdeviceptr dptr) {
    if (!dptr) return CUDA_SUCCESS;                 // cuMemFree_v2@0x32383
    ensure_initialized();
    const DriverTable &d = drv();
    if (!region_) return d.cuMemFree_v2(dptr);
    std::lock_guard<std::mutex> g(table_mu_);
    auto it = table_.find(dptr);
    if (it == table_.end()) {
        // remove_chunk@0x409f0 answers -1 for anything it did not hand out (including pointers from unhooked
        // allocators); strict mode lets the driver decide instead
        return strict_errors() ? d.cuMemFree_v2(dptr) : static_cast<CUresult>(-1);
    }
    Alloc a = it->second;
    CUresult r;
    if (a.kind == AllocKind::Swap) {
        SwapEngine *e = swap(a.dev);
        r = e ? e->free(dptr) : CUDA_ERROR_INVALID_VALUE;
    } else {
        r = d.cuMemFree_v2(dptr);
    }
    // remove_chunk unlinks and un-accounts whatever the real free returned — and it credits the CURRENT device
    // (remove_chunk@0x40871: cuCtxGetDevice(&dev); rm_gpu_device_memory_usage(getpid(), dev, size, 2)), not the one the
    // buffer was charged to: freeing on device 0 what was allocated on device 1 leaves device 1 charged for ever and
    // wraps device 0's lane downwards. Reproduced by default for bit-exact accounting; VGPU_STRICT_CUDA_ERRORS=1
    // credits the owner.
    table_.erase(it);
    int credit = a.dev;
    if (!strict_errors()) { int cur = current_device(); if (cur >= 0 && cur < VGPU_MAX_DEVICES) credit = cur; }
    if (a.kind == AllocKind::Swap) region_->sub(pid_, credit, a.size, VGPU_MEM_BUFFER);
    else uncharge(credit, a.size);
    return r == CUDA_SUCCESS ? CUDA_SUCCESS : r;
}

bool Runtime::check_oom() {
    if (!ensure_initialized()) return false;
    int dev = current_device();
    if (dev < 0) return false;
    if (cfg_.oversubscribe && !cfg_.limit_is_virtual) {
        // swap mode: the quota bounds resident bytes; live swappable bytes above it are the point of the mode, so only
        // the non-swappable part is held against the limit (same rule as charge())
        uint64_t lim = region_->limit(dev);
        if (!lim) return false;
        SwapEngine *e = swap(dev);
        uint64_t u = region_->usage(dev), live = (e ? e->live_bytes() : 0) + region_->swap_live(dev, pid_);
        return fixed_bytes(u, live) > lim;
    }
    return !region_->try_add(pid_, dev, 0, VGPU_MEM_BUFFER, true, true);
}

CUresult Runtime::mem_get_info(size_t *free_b, size_t *total_b) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!region_) return d.cuMemGetInfo_v2(free_b, total_b);
    int dev = current_device();
    if (dev < 0) return d.cuMemGetInfo_v2(free_b, total_b);
    // cuMemGetInfo_v2@0x367dc (memory.c:L549-566)
    uint64_t usage = region_->usage(dev);
    uint64_t limit = region_->limit(dev);
    size_t rf = 0, rt = 0;
    if (limit == 0) {
        CUresult r = d.cuMemGetInfo_v2(&rf, &rt);
        if (r != CUDA_SUCCESS) return r;
        if (total_b) *total_b = rt;
        if (free_b) *free_b = rt - usage;
        return CUDA_SUCCESS;
    }
    if (cfg_.oversubscribe && !cfg_.limit_is_virtual) {
        // swap mode: the quota bounds residency, not live bytes (DESIGN.md "quota semantics"); report the virtual
        // capacity and never the reference's CUDA_ERROR_INVALID_VALUE for usage > limit (@0x36b3a)
        uint64_t vcap = cfg_.virtual_limit[dev];
        if (!vcap) {
            SwapEngine *e = swap(dev);
            uint64_t pool = e ? e->config().host_pool_cap : 0;
            CUresult r = d.cuMemGetInfo_v2(&rf, &rt);
            if (r != CUDA_SUCCESS) return r;
            vcap = pool ? limit + pool : (uint64_t)rt + limit;   // unbounded pool: advertise one device's worth beyond the quota
        }
        if (total_b) *total_b = vcap;
        if (free_b) *free_b = vcap > usage ? vcap - usage : 0;
        return CUDA_SUCCESS;
    }
    if (limit < usage) return CUDA_ERROR_INVALID_VALUE;
    CUresult r = d.cuMemGetInfo_v2(&rf, &rt);
    if (r != CUDA_SUCCESS) return r;
    if (free_b) *free_b = limit - usage;
    if (total_b) *total_b = limit;
    return CUDA_SUCCESS;
}

CUresult Runtime::device_total_mem(size_t *bytes, CUdevice dev) {
    ensure_initialized();
    uint64_t limit = (region_ && dev >= 0 && dev < VGPU_MAX_DEVICES) ? region_->limit(dev) : 0;
    // cuDeviceTotalMem_v2@0x2d3f8 stores the limit unconditionally — 0 bytes for an unlimited container. That is
    // a defect, not a contract: unlimited containers get the driver's answer here (DESIGN.md "deviations");
    // VGPU_REFERENCE_COVERAGE=1 answers like the binary (differential fuzzing with unlimited lanes).
    if (limit == 0 && reference_coverage()) { if (bytes) *bytes = 0; return CUDA_SUCCESS; }
    if (limit == 0) return drv().cuDeviceTotalMem_v2(bytes, dev);
    if (bytes) *bytes = limit;
    return CUDA_SUCCESS;
}

// Context accounting. The reference virtualises contexts (context.c: one `vdevices` record per device, filled by whichever
// of cuDevicePrimaryCtxRetain@0x28c55 / cuCtxCreate_v2@0x29c75 / cuCtxSetCurrent@0x2afa4 reaches the device first) and
// adds context_size as type 0 when it fills the record — so ONCE PER DEVICE per process, however the context came to be
// and however many more are created ("Duplicate cuCtxCreate, this may indicate errors", context.c:144) or destroyed and
// re-created; nothing is ever given back. Checked against the binary: tests/test_hook_parity_cpu.py
// test_context_accounting_is_once_per_device.
void Runtime::charge_context_once(CUdevice dev) {
    if (!region_ || dev < 0 || dev >= VGPU_MAX_DEVICES) return;
    std::lock_guard<std::mutex> g(ctx_mu_);
    if (ctx_charged_[dev]) return;
    ctx_charged_[dev] = true;
    if (context_size_) region_->add(pid_, dev, context_size_, VGPU_MEM_CONTEXT);
}

CUresult Runtime::primary_ctx_retain(CUcontext *ctx, CUdevice dev) {
    ensure_initialized();
    CUresult r = drv().cuDevicePrimaryCtxRetain(ctx, dev);
    if (r == CUDA_SUCCESS) charge_context_once(dev);
    return r;
}

CUresult Runtime::ctx_create(CUcontext *ctx, unsigned flags, CUdevice dev) {
    ensure_initialized();
    CUresult r = drv().cuCtxCreate_v2(ctx, flags, dev);
    if (r == CUDA_SUCCESS) charge_context_once(dev);
    return r;
}

bool Runtime::nvml_memory_view(int idx, unsigned long long *total, unsigned long long *free_b, unsigned long long *used) {
    // nvmlDeviceGetMemoryInfo@0x24069 (nvml/hook.c:L327-334): usage = Σ used[d].total; under MEMORY_OVERRIDE=1 the
    // NVML-monitored figure wins when it is larger; limit == 0 -> only `used` is replaced (by the container's usage),
    // otherwise total = limit, free = limit - usage, used = usage.
    if (!ensure_initialized() || idx < 0 || idx >= VGPU_MAX_DEVICES) return false;
    uint64_t limit = region_->limit(idx);
    uint64_t usage = region_->usage(idx);
    static const bool override_on = [] { const char *e = std::getenv("MEMORY_OVERRIDE"); return e && std::atoi(e) == 1; }();
    if (override_on) {
        uint64_t monitor = 0;
        vgpu_shared_region_t *r = region_->raw();
        for (int i = 0; i < r->proc_num; i++) monitor += r->procs[i].monitorused[idx];
        if (usage < monitor) usage = monitor;
    }
    *used = usage;
    if (limit == 0) return true;           // caller keeps the driver's total/free (signalled by *total == 0)
    *total = limit;
    *free_b = limit - usage;               // unsigned wrap when usage > limit, like the reference (@0x244b0-0x244b8)
    return true;
}

// ------------------------------------------------------------------------------------------------ launches
namespace {
struct ParamLayout { std::vector<std::pair<size_t, size_t>> p; bool known = false; };
std::mutex g_layout_mu;
std::unordered_map<CUfunction, std::shared_ptr<const ParamLayout>> g_layouts;

// shared_ptr: cuModuleUnload may clear the cache while another thread is still scanning with a layout
std::shared_ptr<const ParamLayout> layout_of(CUfunction f) {
    std::lock_guard<std::mutex> g(g_layout_mu);
    auto it = g_layouts.find(f);
    if (it != g_layouts.end()) return it->second;
    ParamLayout L;
    const DriverTable &d = drv();
    if (d.cuFuncGetParamInfo) {
        L.known = true;
        for (size_t i = 0; i < 1024; i++) {
            size_t off = 0, sz = 0;
            if (d.cuFuncGetParamInfo(f, i, &off, &sz) != CUDA_SUCCESS) break;
            L.p.emplace_back(off, sz);
        }
    }
    auto sp = std::make_shared<const ParamLayout>(std::move(L));
    g_layouts.emplace(f, sp);
    return sp;
}
}  // namespace

void Runtime::forget_function_layouts() {
    std::lock_guard<std::mutex> g(g_layout_mu);
    g_layouts.clear();
}

// Collects the swap-table rows referenced by a launch's arguments (pointers falling inside the swap arena).
static void collect_launch_rows(SwapEngine *e, CUfunction f, void **params, void **extra, std::vector<int> *rows) {
    std::shared_ptr<const ParamLayout> Lp = layout_of(f);
    const ParamLayout &L = *Lp;
    if (params && L.known) {
        for (size_t i = 0; i < L.p.size(); i++)
            if (L.p[i].second >= 8 && params[i]) e->collect_rows(params[i], L.p[i].second, rows);
    } else if (extra) {
        // CU_LAUNCH_PARAM_BUFFER_POINTER / _SIZE pairs
        void *buf = nullptr; size_t size = 0;
        for (int i = 0; i < 16 && extra[i] != CU_LAUNCH_PARAM_END; i += 2) {
            if (extra[i] == CU_LAUNCH_PARAM_BUFFER_POINTER) buf = extra[i + 1];
            else if (extra[i] == CU_LAUNCH_PARAM_BUFFER_SIZE) size = *static_cast<size_t *>(extra[i + 1]);
        }
        if (buf && size) e->collect_rows(buf, size, rows);
    } else if (params && !L.known) {
        static bool warned = false;
        if (!warned) { warned = true; LOG_ERROR("driver lacks cuFuncGetParamInfo: kernel arguments cannot be scanned, swapped-out buffers may fault"); }
    }
}

static bool stream_is_capturing(CUstream st) {
    if (!drv().cuStreamIsCapturing) return false;
    CUstreamCaptureStatus cs = CU_STREAM_CAPTURE_STATUS_NONE;
    if (drv().cuStreamIsCapturing(st, &cs) != CUDA_SUCCESS) return false;   // legacy stream during a global capture: the launch itself will fail
    return cs != CU_STREAM_CAPTURE_STATUS_NONE;
}

// Shared shape of every launch intercept: limiter gate -> swap admission -> real launch -> bookkeeping.
template <typename RealLaunch>
static CUresult guarded_launch(Runtime &rt, const Config &cfg, Limiter *lim, CUfunction f, void **params, void **extra,
                               CUstream st, RealLaunch real) {
    SwapEngine *e = nullptr;
    thread_local std::vector<int> rows;
    rows.clear();
    if ((lim || cfg.oversubscribe) && stream_is_capturing(st)) {
        // Stream capture: nothing executes now, and a capturing stream may neither carry the limiter's stamp kernels
        // (they would be replayed with every graph launch) nor wait on the engine's page-in events (capture
        // isolation). The launch is billed when the graph runs (graph_launch); its swappable operands are made
        // resident on the host's time and stay pinned, since a replay cannot fault them back in.
        if (cfg.oversubscribe) {
            CUdevice dev = -1;
            if (drv().cuCtxGetDevice(&dev) == CUDA_SUCCESS) e = rt.swap((int)dev);
            if (e) {
                collect_launch_rows(e, f, params, extra, &rows);
                if (!rows.empty()) {
                    CUresult r = e->ensure_resident(rows.data(), (int)rows.size(), SwapEngine::kHostWait);
                    if (r != CUDA_SUCCESS) return r;
                }
            }
        }
        return real();
    }
    if (lim) lim->before_launch(st);     // rate_limiter@0x4591a position: before the real launch
    if (cfg.oversubscribe) {
        CUdevice dev = -1;
        if (drv().cuCtxGetDevice(&dev) == CUDA_SUCCESS) e = rt.swap((int)dev);
        if (e) {
            collect_launch_rows(e, f, params, extra, &rows);
            if (!rows.empty()) {
                CUresult r = e->ensure_resident(rows.data(), (int)rows.size(), st);
                if (r != CUDA_SUCCESS) return r;
            }
        }
    }
    CUresult r = real();
    if (e && !rows.empty()) e->note_use(rows.data(), (int)rows.size(), st);
    if (lim) lim->after_launch(st);
    return r;
}

CUresult Runtime::launch_kernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                unsigned smem, CUstream st, void **params, void **extra) {
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, extra, st,
                          [&] { return drv().cuLaunchKernel(f, gx, gy, gz, bx, by, bz, smem, st, params, extra); });
}

CUresult Runtime::launch_kernel_ex(const CUlaunchConfig *cfg, CUfunction f, void **params, void **extra) {
    if (!cfg || !drv().cuLaunchKernelEx) return CUDA_ERROR_NOT_SUPPORTED;
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, extra, cfg->hStream,
                          [&] { return drv().cuLaunchKernelEx(cfg, f, params, extra); });
}

CUresult Runtime::launch_cooperative(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                     unsigned bz, unsigned smem, CUstream st, void **params) {
    // the reference lets cooperative launches bypass rate_limiter (Appendix E); here they are limited and admitted
    // like any other launch
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, nullptr, st,
                          [&] { return drv().cuLaunchCooperativeKernel(f, gx, gy, gz, bx, by, bz, smem, st, params); });
}

static inline CUstream pt(CUstream st) { return st ? st : CU_STREAM_PER_THREAD; }

CUresult Runtime::launch_kernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                     unsigned smem, CUstream st, void **params, void **extra) {
    if (!drv().cuLaunchKernel_ptsz) return CUDA_ERROR_NOT_SUPPORTED;
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, extra, pt(st),
                          [&] { return drv().cuLaunchKernel_ptsz(f, gx, gy, gz, bx, by, bz, smem, st, params, extra); });
}

CUresult Runtime::launch_kernel_ex_ptsz(const CUlaunchConfig *cfg, CUfunction f, void **params, void **extra) {
    if (!cfg || !drv().cuLaunchKernelEx_ptsz) return CUDA_ERROR_NOT_SUPPORTED;
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, extra, pt(cfg->hStream),
                          [&] { return drv().cuLaunchKernelEx_ptsz(cfg, f, params, extra); });
}

CUresult Runtime::launch_cooperative_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                          unsigned bz, unsigned smem, CUstream st, void **params) {
    if (!drv().cuLaunchCooperativeKernel_ptsz) return CUDA_ERROR_NOT_SUPPORTED;
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    return guarded_launch(*this, cfg_, limiter_.get(), f, params, nullptr, pt(st),
                          [&] { return drv().cuLaunchCooperativeKernel_ptsz(f, gx, gy, gz, bx, by, bz, smem, st, params); });
}

// A graph launch is one unit of GPU work for the core limiter: gated like a kernel launch and billed by the same
// device-side stamps (the limiter measures busy time around whatever was enqueued, so a whole graph is covered). The
// reference has no hook on graph launches at all — a captured decode loop runs unthrottled there.
CUresult Runtime::graph_launch(CUgraphExec g, CUstream st, bool ptsz) {
    auto real = ptsz ? drv().cuGraphLaunch_ptsz : drv().cuGraphLaunch;
    if (!real) return CUDA_ERROR_NOT_SUPPORTED;
    if (!initialized()) ensure_initialized();
    if (!post_inited_.load(std::memory_order_acquire)) post_init(true);
    Limiter *lim = reference_coverage() ? nullptr : limiter_.get();
    CUstream eff = ptsz ? pt(st) : st;
    if (lim) lim->before_launch(eff);
    CUresult r = real(g, st);
    if (lim) lim->after_launch(eff, /*heavy=*/true);
    return r;
}

// Stream-ordered allocations come out of a driver memory pool: charged to the quota like cuMemAlloc (requested bytes),
// tracked in the same table so cuMemFree_v2 / cuMemFreeAsync / check_memory_type see them. Never swappable.
CUresult Runtime::mem_alloc_async(CUdeviceptr *dptr, size_t bytes, CUmemoryPool pool, bool from_pool, CUstream st, bool ptsz) {
    ensure_initialized();
    const DriverTable &d = drv();
    auto real = [&]() -> CUresult {
        if (from_pool) {
            auto f = ptsz ? d.cuMemAllocFromPoolAsync_ptsz : d.cuMemAllocFromPoolAsync;
            return f ? f(dptr, bytes, pool, st) : CUDA_ERROR_NOT_SUPPORTED;
        }
        auto f = ptsz ? d.cuMemAllocAsync_ptsz : d.cuMemAllocAsync;
        return f ? f(dptr, bytes, st) : CUDA_ERROR_NOT_SUPPORTED;
    };
    if (!region_ && fail_closed_) return CUDA_ERROR_OUT_OF_MEMORY;
    if (!region_ || reference_coverage()) return real();
    int dev = current_device();
    if (dev < 0) return real();
    if (!charge(dev, bytes)) return CUDA_ERROR_OUT_OF_MEMORY;
    CUresult r = real();
    if (r != CUDA_SUCCESS) { uncharge(dev, bytes); return r; }
    std::lock_guard<std::mutex> g(table_mu_);
    track(*dptr, bytes, dev, AllocKind::Async);
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_free_async(CUdeviceptr dptr, CUstream st, bool ptsz) {
    ensure_initialized();
    const DriverTable &d = drv();
    auto real = ptsz ? d.cuMemFreeAsync_ptsz : d.cuMemFreeAsync;
    if (!real) return CUDA_ERROR_NOT_SUPPORTED;
    if (!dptr || !region_) return real(dptr, st);
    Alloc a;
    {
        std::lock_guard<std::mutex> g(table_mu_);
        auto it = table_.find(dptr);
        if (it == table_.end()) return real(dptr, st);          // not ours (allocated before the hook, or coverage off)
        a = it->second;
        if (a.kind == AllocKind::Swap) {
            // a swappable buffer has no driver allocation behind its address: order the free behind the stream, then
            // let the engine release it
            d.cuStreamSynchronize(ptsz ? pt(st) : st);
            SwapEngine *e = swap(a.dev);
            CUresult r = e ? e->free(dptr) : CUDA_ERROR_INVALID_VALUE;
            table_.erase(it);
            region_->sub(pid_, a.dev, a.size, VGPU_MEM_BUFFER);
            return r;
        }
        table_.erase(it);
    }
    CUresult r = real(dptr, st);
    uncharge(a.dev, a.size);
    return r;
}

// VMM physical allocations (PyTorch expandable segments, NCCL user buffers): the handle is what occupies HBM, mapped
// or not, so the handle is what gets charged. Host-located handles are not device memory.
CUresult Runtime::mem_create(CUmemGenericAllocationHandle *h, size_t bytes, const CUmemAllocationProp *prop, unsigned long long flags) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!d.cuMemCreate) return CUDA_ERROR_NOT_SUPPORTED;
    if (!region_ && fail_closed_ && prop && prop->location.type == CU_MEM_LOCATION_TYPE_DEVICE) return CUDA_ERROR_OUT_OF_MEMORY;
    if (!region_ || reference_coverage() || !prop || prop->location.type != CU_MEM_LOCATION_TYPE_DEVICE)
        return d.cuMemCreate(h, bytes, prop, flags);
    int dev = prop->location.id;
    if (dev < 0 || dev >= VGPU_MAX_DEVICES) return d.cuMemCreate(h, bytes, prop, flags);
    if (!charge(dev, bytes)) return CUDA_ERROR_OUT_OF_MEMORY;
    CUresult r = d.cuMemCreate(h, bytes, prop, flags);
    if (r != CUDA_SUCCESS) { uncharge(dev, bytes); return r; }
    std::lock_guard<std::mutex> g(table_mu_);
    phys_[*h] = PhysAlloc{bytes, dev, 0, false};
    return CUDA_SUCCESS;
}

CUresult Runtime::mem_release(CUmemGenericAllocationHandle h) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!d.cuMemRelease) return CUDA_ERROR_NOT_SUPPORTED;
    CUresult r = d.cuMemRelease(h);
    if (!region_ || r != CUDA_SUCCESS) return r;
    PhysAlloc a;
    {
        std::lock_guard<std::mutex> g(table_mu_);
        auto it = phys_.find(h);
        if (it == phys_.end()) return r;                        // imported / retained handle, or coverage off
        if (it->second.maps > 0) { it->second.released = true; return r; }   // still mapped: the memory stays, so does the charge
        a = it->second;
        phys_.erase(it);
    }
    uncharge(a.dev, a.size);
    return r;
}

// cuMemMap / cuMemUnmap of the APPLICATION's own ranges (the swap engine calls the driver directly): only bookkeeping, so
// that a handle released while mapped — the idiom of the CUDA VMM samples — keeps its charge until the last mapping goes.
CUresult Runtime::mem_map(CUdeviceptr ptr, size_t size, size_t offset, CUmemGenericAllocationHandle h, unsigned long long flags) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!d.cuMemMap) return CUDA_ERROR_NOT_SUPPORTED;
    CUresult r = d.cuMemMap(ptr, size, offset, h, flags);
    if (r != CUDA_SUCCESS || !region_) return r;
    std::lock_guard<std::mutex> g(table_mu_);
    auto it = phys_.find(h);
    if (it == phys_.end()) return r;
    it->second.maps++;
    vmaps_[ptr] = {size, h};
    return r;
}

CUresult Runtime::mem_unmap(CUdeviceptr ptr, size_t size) {
    ensure_initialized();
    const DriverTable &d = drv();
    if (!d.cuMemUnmap) return CUDA_ERROR_NOT_SUPPORTED;
    CUresult r = d.cuMemUnmap(ptr, size);
    if (r != CUDA_SUCCESS || !region_) return r;
    std::vector<PhysAlloc> gone;
    {
        std::lock_guard<std::mutex> g(table_mu_);
        for (auto m = vmaps_.lower_bound(ptr); m != vmaps_.end() && m->first < ptr + size;) {
            auto it = phys_.find(m->second.second);
            if (it != phys_.end() && --it->second.maps <= 0 && it->second.released) { gone.push_back(it->second); phys_.erase(it); }
            m = vmaps_.erase(m);
        }
    }
    for (const PhysAlloc &a : gone) uncharge(a.dev, a.size);
    return r;
}

// Rows pinned between touch_range*() / touch_batch() and touch_done(). A copy can touch two engines: a peer copy between
// two devices of the container, or any copy issued while another device's context is current — every pointer is looked
// up in the engine whose arena holds it, not in the current device's.
struct TouchRef { SwapEngine *e; int row; bool writes; };
static thread_local std::vector<TouchRef> t_touch;

SwapEngine *Runtime::engine_of(CUdeviceptr p) {
    if (!p) return nullptr;
    for (int d = 0; d < VGPU_MAX_DEVICES; d++)
        if (SwapEngine *e = swap_[d].get())
            if (e->owns(p)) return e;
    return nullptr;
}

void Runtime::touch_range(CUdeviceptr p, size_t bytes, CUstream st, bool writes) {
    if (writes) touch_range2(p, bytes, 0, 0, st);
    else touch_range2(0, 0, p, bytes, st);
}

void Runtime::touch_range2(CUdeviceptr dst, size_t dbytes, CUdeviceptr src, size_t sbytes, CUstream st) {
    if (!cfg_.oversubscribe) return;
    (void)dbytes; (void)sbytes;
    SwapEngine *ed = engine_of(dst), *es = engine_of(src);
    int rd = ed ? ed->lookup(dst) : -1, rs = es ? es->lookup(src) : -1;
    if (rd < 0 && rs < 0) return;
    const bool capturing = stream_is_capturing(st);   // captured copy node: operands pinned resident, nothing recorded into the capture
    CUstream how = capturing ? SwapEngine::kHostWait : st;
    t_touch.clear();
    auto admit = [&](SwapEngine *e, const int *rows, int n) {
        if (e->ensure_resident(rows, n, how) != CUDA_SUCCESS) { LOG_ERROR("memcpy/memset operand could not be made resident"); return false; }
        return true;
    };
    if (rd >= 0 && rs >= 0 && ed == es) {
        int rows[2] = {rd, rs};
        if (!admit(ed, rows, rd == rs ? 1 : 2) || capturing) return;
        t_touch.push_back(TouchRef{ed, rd, true});
        if (rs != rd) t_touch.push_back(TouchRef{es, rs, false});
        return;
    }
    // one operand, or two operands in two engines: one admission per engine
    if (rd >= 0 && admit(ed, &rd, 1) && !capturing) t_touch.push_back(TouchRef{ed, rd, true});
    if (rs >= 0 && admit(es, &rs, 1) && !capturing) t_touch.push_back(TouchRef{es, rs, false});
    // the copy itself is enqueued by the caller right after this returns; note_use after it keeps the rows pinned
}

bool Runtime::touch_batch(const CUdeviceptr *written, size_t nw, const CUdeviceptr *read, size_t nr, CUstream st) {
    t_touch.clear();
    if (!cfg_.oversubscribe) return true;
    SwapEngine *e = nullptr;
    bool several = false;
    std::vector<int> w, r;
    auto add = [&](std::vector<int> &v, CUdeviceptr p) {
        SwapEngine *pe = engine_of(p);
        if (!pe) return;
        if (e && pe != e) { several = true; return; }
        e = pe;
        int row = e->lookup(p);
        if (row < 0) return;
        if (std::find(w.begin(), w.end(), row) != w.end()) return;     // written wins over read
        if (std::find(v.begin(), v.end(), row) == v.end()) v.push_back(row);
    };
    for (size_t i = 0; i < nw; i++) add(w, written[i]);
    for (size_t i = 0; i < nr; i++) add(r, read[i]);
    if (several) return false;                     // operands in two engines: the caller issues the copies one by one
    if (w.empty() && r.empty()) return true;
    std::vector<int> all(w);
    all.insert(all.end(), r.begin(), r.end());
    if (!e->fits_together(all.data(), (int)all.size())) return false;       // likewise: nothing to report
    bool capturing = stream_is_capturing(st);
    CUresult rc = e->ensure_resident(all.data(), (int)all.size(), capturing ? SwapEngine::kHostWait : st);
    if (rc != CUDA_SUCCESS) return false;
    if (capturing) return true;                    // captured: the operands stay pinned, nothing is recorded into the capture
    for (int row : w) t_touch.push_back(TouchRef{e, row, true});
    for (int row : r) t_touch.push_back(TouchRef{e, row, false});
    return true;                                   // touch_done() after the real call unpins and records the use
}

void Runtime::touch_done(CUstream st) {
    if (t_touch.empty()) return;
    // per engine: its written rows, then its read-only rows (one admission each: the second call does not close another)
    while (!t_touch.empty()) {
        SwapEngine *e = t_touch.front().e;
        std::vector<int> w, r;
        for (auto it = t_touch.begin(); it != t_touch.end();) {
            if (it->e != e) { ++it; continue; }
            (it->writes ? w : r).push_back(it->row);
            it = t_touch.erase(it);
        }
        if (!w.empty()) e->note_use(w.data(), (int)w.size(), st, true);
        if (!r.empty()) e->note_use(r.data(), (int)r.size(), st, false, /*closes_admission=*/w.empty());
    }
}

bool Runtime::swap_advise(CUdeviceptr p, CUmem_advise advice) {
    if (!cfg_.oversubscribe) return false;
    SwapEngine *e = engine_of(p);
    if (!e) return false;
    int row = e->lookup(p);
    if (row < 0) return false;
    // the only advice with a meaning for explicit paging: a read-mostly range is not dirtied by kernel launches, so
    // evicting it needs no write-back (UVM: read duplication). Location / accessed-by hints are accepted and ignored.
    if (advice == CU_MEM_ADVISE_SET_READ_MOSTLY) e->advise_read_mostly(row, true);
    else if (advice == CU_MEM_ADVISE_UNSET_READ_MOSTLY) e->advise_read_mostly(row, false);
    return true;
}

CUresult Runtime::pin_graph_kernel(CUfunction f, void **params, void **extra) {
    if (!cfg_.oversubscribe || !f) return CUDA_SUCCESS;
    SwapEngine *e = swap(current_device());
    if (!e) return CUDA_SUCCESS;
    std::vector<int> rows;
    collect_launch_rows(e, f, params, extra, &rows);
    if (rows.empty()) return CUDA_SUCCESS;
    return e->ensure_resident(rows.data(), (int)rows.size(), SwapEngine::kHostWait);
}

CUresult Runtime::pin_graph_ptrs(const CUdeviceptr *p, size_t n) {
    if (!cfg_.oversubscribe) return CUDA_SUCCESS;
    CUresult rc = CUDA_SUCCESS;
    for (size_t i = 0; i < n; i++) {               // one admission per pointer: the operands of a copy node may live in two engines
        SwapEngine *e = engine_of(p[i]);
        int r = e ? e->lookup(p[i]) : -1;
        if (r < 0 || (i > 0 && p[i] == p[i - 1])) continue;
        CUresult one = e->ensure_resident(&r, 1, SwapEngine::kHostWait);
        if (one != CUDA_SUCCESS) rc = one;
    }
    return rc;
}

bool Runtime::swap_address_range(CUdeviceptr p, CUdeviceptr *base, size_t *size) {
    if (!cfg_.oversubscribe) return false;
    SwapEngine *e = engine_of(p);
    return e && e->range_of(p, base, size);
}

bool Runtime::swap_prefetch(CUdeviceptr p, bool to_device) {
    if (!cfg_.oversubscribe) return false;
    SwapEngine *e = engine_of(p);
    if (!e) return false;
    int row = e->lookup(p);
    if (row < 0) return false;
    if (to_device) e->hint_prefetch(row);
    else e->hint_evict(row);                     // "to the host": the application is done with it on the device for now
    return true;
}

}  // namespace vgpu
