// region.cc — see region.h.
#include "region.h"

#include <fcntl.h>
#include <semaphore.h>
#include <sys/file.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "log.h"

namespace vgpu {

// strtoul(text, NULL, 0) as the reference binary gets it: its call binds to the classic strtoul@GLIBC_2.2.5, while code
// compiled today against glibc >= 2.38 is redirected to __isoc23_strtoul, which also accepts "0b"/"0B" binary literals
// ("0B1k" would parse as 1 KiB here and as 0 there). Spelled out so the result does not depend on the C library's
// vintage: leading white space, optional sign, 0x/0X hex, leading 0 octal, else decimal; overflow -> ULONG_MAX; a
// negative number wraps like the library's.
static uint64_t classic_strtoul_base0(const char *p) {
    while (*p == ' ' || (*p >= '\t' && *p <= '\r')) p++;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; p++; }
    unsigned base = 10;
    if (*p == '0') {
        auto hex = [](char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); };
        if ((p[1] == 'x' || p[1] == 'X') && hex(p[2])) { base = 16; p += 2; }
        else base = 8;
    }
    uint64_t acc = 0;
    bool overflow = false;
    for (;; p++) {
        unsigned d;
        if (*p >= '0' && *p <= '9') d = (unsigned)(*p - '0');
        else if (*p >= 'a' && *p <= 'z') d = (unsigned)(*p - 'a') + 10;
        else if (*p >= 'A' && *p <= 'Z') d = (unsigned)(*p - 'A') + 10;
        else break;
        if (d >= base) break;
        if (acc > (UINT64_MAX - d) / base) overflow = true;
        acc = acc * base + d;
    }
    if (overflow) return UINT64_MAX;
    return neg ? (uint64_t)0 - acc : acc;
}

uint64_t parse_limit(const char *v) {
    // get_limit_from_env@0x40d00 (multiprocess_memory_limit.c:L101-111): unit from the LAST character, number by
    // strtoul(base 0), zero product or overflowing product -> 0 (= unlimited)
    if (!v) return 0;
    size_t len = std::strlen(v);
    if (!len) return 0;
    uint64_t scalar = 1;
    switch (v[len - 1]) {
        case 'G': case 'g': scalar = 1ull << 30; break;
        case 'M': case 'm': scalar = 1ull << 20; break;
        case 'K': case 'k': scalar = 1ull << 10; break;
        default: break;
    }
    uint64_t n = classic_strtoul_base0(v);
    uint64_t prod = n * scalar;
    if (prod == 0 || prod / scalar != n) return 0;
    return prod;
}

uint64_t limit_from_env(const char *base, int dev) {
    char name[96];
    std::snprintf(name, sizeof name, "%s_%d", base, dev);
    uint64_t v = parse_limit(std::getenv(name));
    if (v) return v;
    return parse_limit(std::getenv(base));
}

static sem_t *sem_of(vgpu_shared_region_t *r) { return reinterpret_cast<sem_t *>(r->sem); }
static_assert(sizeof(sem_t) <= 32, "sem_t must fit the 32-byte lane of the region");

static bool pid_alive(int32_t pid) {
    // proc_alive@0x40aac parses /proc/<pid>/stat; existence of the directory is the same predicate
    if (pid <= 0) return false;
    char p[64];
    std::snprintf(p, sizeof p, "/proc/%d/stat", pid);
    FILE *f = std::fopen(p, "r");
    if (!f) return false;
    char state = '?';
    int rd = std::fscanf(f, "%*d %*[^)]%*c %c", &state);
    std::fclose(f);
    return rd != 1 || (state != 'Z' && state != 'X');
}

Region *Region::open(const char *path, bool create, const uint64_t *mem_limits, const uint64_t *sm_limits, int priority,
                     const char (*uuids)[VGPU_UUID_LEN], int ndev, std::string *err) {
    mode_t old = umask(0);  // file must be usable by every uid of the container and by the monitor (0666, @0x4429e)
    int fd = ::open(path, create ? (O_RDWR | O_CREAT) : O_RDWR, 0666);
    umask(old);
    if (fd < 0) {
        if (err) *err = std::string("open ") + path + ": " + std::strerror(errno);
        return nullptr;
    }
    // size check and growth under the whole-file lock, and never downwards: a second creator that saw size 0 must not cut
    // off the extension block a first creator has already appended behind the reference's bytes
    if (lockf(fd, F_LOCK, 0) != 0) LOG_WARN("lockf(%s): %s", path, std::strerror(errno));
    struct stat st;
    bool size_ok = fstat(fd, &st) == 0 && static_cast<uint64_t>(st.st_size) >= VGPU_REGION_SIZE;
    if (!size_ok && create) size_ok = ftruncate(fd, VGPU_REGION_SIZE) == 0;
    if (lockf(fd, F_ULOCK, 0) != 0) {}
    if (!size_ok) {
        if (err) *err = "region file too small";
        ::close(fd);
        return nullptr;
    }
    void *m = mmap(nullptr, VGPU_REGION_SIZE, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) {
        if (err) *err = std::string("mmap: ") + std::strerror(errno);
        ::close(fd);
        return nullptr;
    }
    Region *R = new Region();
    R->r_ = static_cast<vgpu_shared_region_t *>(m);
    R->fd_ = fd;
    R->path_ = path;
    vgpu_shared_region_t *r = R->r_;

    // creation race: whole-file advisory lock while the magic is examined/written (lockf @0x4458c)
    if (lockf(fd, F_LOCK, VGPU_REGION_SIZE) != 0) LOG_WARN("lockf(%s): %s", path, std::strerror(errno));
    if (r->initialized_flag != VGPU_REGION_MAGIC) {
        if (!create) {
            if (lockf(fd, F_ULOCK, VGPU_REGION_SIZE) != 0) {}
            if (err) *err = "region not initialised";
            delete R;
            return nullptr;
        }
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) {
            r->limit[d] = mem_limits ? mem_limits[d] : 0;
            r->sm_limit[d] = sm_limits ? sm_limits[d] : 100;
        }
        if (sem_init(sem_of(r), 1, 1) != 0) LOG_ERROR("sem_init: %s", std::strerror(errno));
        r->owner_pid = 0;
        r->sm_init_flag = 0;
        r->utilization_switch = 1;
        r->recent_kernel = 2;
        r->priority = priority;
        r->proc_num = 0;
        if (uuids && ndev > 0) {
            r->device_num = static_cast<uint64_t>(ndev);
            for (int d = 0; d < ndev && d < VGPU_MAX_DEVICES; d++) std::memcpy(r->uuids[d], uuids[d], VGPU_UUID_LEN);
        }
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        r->initialized_flag = VGPU_REGION_MAGIC;
    } else if (mem_limits) {
        for (int d = 0; d < VGPU_MAX_DEVICES; d++) {
            if (mem_limits[d] != r->limit[d])
                LOG_ERROR("Limit inconsistency detected for %dth device, %lu expected, get %lu", d,
                          (unsigned long)r->limit[d], (unsigned long)mem_limits[d]);
            if (sm_limits && sm_limits[d] != r->sm_limit[d])
                LOG_INFO("SM limit inconsistency for device %d: %lu stored, %lu in env", d, (unsigned long)r->sm_limit[d],
                         (unsigned long)sm_limits[d]);
        }
    }
    // extension block: grow the file when it ends before it (first creator, or a file made by the reference hook)
    {
        const off_t need = (off_t)VGPU_REGION_EXT_OFFSET + (off_t)sizeof(vgpu_region_ext_t);
        struct stat st2;
        bool ok = fstat(fd, &st2) == 0 && (st2.st_size >= need || (create && ftruncate(fd, need) == 0));
        if (ok) {
            void *em = mmap(nullptr, sizeof(vgpu_region_ext_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, VGPU_REGION_EXT_OFFSET);
            if (em != MAP_FAILED) {
                R->ext_ = static_cast<vgpu_region_ext_t *>(em);
                if (R->ext_->magic != VGPU_REGION_EXT_MAGIC) {
                    if (create) {
                        std::memset(R->ext_, 0, sizeof(vgpu_region_ext_t));
                        R->ext_->version = 1;
                        __atomic_thread_fence(__ATOMIC_SEQ_CST);
                        R->ext_->magic = VGPU_REGION_EXT_MAGIC;
                    } else {
                        munmap(em, sizeof(vgpu_region_ext_t));
                        R->ext_ = nullptr;
                    }
                }
            }
        }
    }
    if (lockf(fd, F_ULOCK, VGPU_REGION_SIZE) != 0) {}
    return R;
}

vgpu_swap_record_t *Region::swap_record(int32_t pid, int dev) {
    if (!ext_) return nullptr;
    lock();
    vgpu_swap_record_t *hit = nullptr, *empty = nullptr;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++) {
        vgpu_swap_record_t &rec = ext_->swap[i];
        if (rec.pid == pid && rec.dev == dev) { hit = &rec; break; }
        if (!empty && rec.pid == 0) empty = &rec;
    }
    if (!hit && !empty) {          // full: records of dead processes are reclaimed with their slots
        reap_dead_locked();
        for (int i = 0; i < VGPU_REGION_EXT_RECORDS && !empty; i++)
            if (ext_->swap[i].pid == 0) empty = &ext_->swap[i];
    }
    if (!hit && empty) {
        std::memset(empty, 0, sizeof *empty);
        empty->dev = dev;
        __atomic_thread_fence(__ATOMIC_RELEASE);
        empty->pid = pid;
        hit = empty;
    }
    unlock();
    return hit;
}

bool Region::swap_counters(int dev, vgpu_swap_record_t *out) {
    std::memset(out, 0, sizeof *out);
    out->dev = dev;
    if (!ext_) return false;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++) {
        const vgpu_swap_record_t &rec = ext_->swap[i];
        if (rec.pid == 0 || rec.dev != dev) continue;
        out->pid++;
        out->page_out_bytes += rec.page_out_bytes; out->page_in_bytes += rec.page_in_bytes;
        out->evictions += rec.evictions; out->faults += rec.faults;
        out->resident_bytes += rec.resident_bytes; out->live_bytes += rec.live_bytes; out->host_bytes += rec.host_bytes;
    }
    return true;
}

uint64_t Region::swap_live(int dev, int32_t except_pid) const {
    if (!ext_) return 0;
    uint64_t sum = 0;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++) {
        const vgpu_swap_record_t &rec = ext_->swap[i];
        if (rec.pid == 0 || rec.dev != dev || rec.pid == except_pid) continue;
        sum += __atomic_load_n(&rec.live_bytes, __ATOMIC_RELAXED);
    }
    return sum;
}

uint64_t Region::swap_reserve(int32_t pid, int dev, uint64_t want_total, uint64_t live_mapped, uint64_t overhead, bool *granted, int *engines, uint64_t *share_out) {
    *granted = false;
    if (engines) *engines = 1;
    if (!ext_) { *granted = true; return ~0ull; }
    lock();
    bool swept = false;
again:
    const uint64_t lim = r_->limit[dev];
    vgpu_swap_record_t *mine = nullptr;
    uint64_t live_all = 0;
    int n = 0;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++) {
        vgpu_swap_record_t &rec = ext_->swap[i];
        if (rec.pid == 0 || rec.dev != dev) continue;
        if (rec.pid == pid) { mine = &rec; rec.live_bytes = live_mapped; }
        if (rec.live_bytes == 0 && rec.resident_bytes == 0 && rec.pid != pid) continue;   // an engine that holds nothing takes no share
        n++;
        live_all += rec.live_bytes;
    }
    if (n == 0) n = 1;
    if (engines) *engines = n;
    if (lim == 0) { if (mine) mine->resident_bytes = want_total; unlock(); *granted = true; return ~0ull; }
    // non-swappable bytes of the whole container = accounted usage - every engine's live swappable bytes (the lanes use the
    // reference's wrapping arithmetic, so the difference is read as signed)
    int64_t fixed = (int64_t)(usage_locked(dev) - live_all);
    if (fixed < 0) fixed = 0;
    uint64_t taken = (uint64_t)fixed + (uint64_t)n * overhead;
    uint64_t room = lim > taken ? lim - taken : 0;
    uint64_t share = room / (uint64_t)n;
    if (share_out) *share_out = share;
    uint64_t others = 0;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++) {
        const vgpu_swap_record_t &rec = ext_->swap[i];
        if (rec.pid == 0 || rec.dev != dev || rec.pid == pid) continue;
        if (rec.live_bytes == 0 && rec.resident_bytes == 0) continue;
        uint64_t entitled = rec.live_bytes < share ? rec.live_bytes : share;
        others += rec.resident_bytes > entitled ? rec.resident_bytes : entitled;
    }
    uint64_t cap = room > others ? room - others : 0;
    if (n > 1 && !swept) {
        // What siblings hold or are entitled to shapes this engine's cap: a sibling that was killed (no exit handler ran)
        // still has its record and its slot. Look — /proc, at most every 100 ms — and judge again without the dead.
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        uint64_t now = (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
        if (now - last_dead_sweep_ns_ > 100000000ull) {
            last_dead_sweep_ns_ = now;
            swept = true;
            if (reap_dead_locked() > 0) goto again;
        }
    }
    if (want_total <= cap) { *granted = true; if (mine) mine->resident_bytes = want_total; }
    unlock();
    return cap;
}

void Region::clear_swap_records_locked(int32_t pid) {
    if (!ext_) return;
    for (int i = 0; i < VGPU_REGION_EXT_RECORDS; i++)
        if (ext_->swap[i].pid == pid) std::memset(&ext_->swap[i], 0, sizeof(vgpu_swap_record_t));
}

Region::~Region() {
    if (ext_) munmap(ext_, sizeof(vgpu_region_ext_t));
    if (r_) munmap(r_, VGPU_REGION_SIZE);
    if (fd_ >= 0) ::close(fd_);
}

void Region::lock() {
    // lock_shrreg@0x437d3 (multiprocess_memory_limit.c:L518-542): sem_timedwait(10 s); on timeout take the lock
    // over when its recorded owner is this process or a dead one (fix_lock_shrreg@0x4317c, under lockf), or after
    // 30 timeouts with no recorded owner.
    int trials = 0;
    for (;;) {
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        ts.tv_sec += 10;
        if (sem_timedwait(sem_of(r_), &ts) == 0) {
            r_->owner_pid = static_cast<uint64_t>(getpid());
            __atomic_thread_fence(__ATOMIC_SEQ_CST);
            return;
        }
        if (errno == EINTR) continue;
        if (errno != ETIMEDOUT) {
            LOG_ERROR("region lock: %s", std::strerror(errno));
            continue;
        }
        int32_t owner = static_cast<int32_t>(r_->owner_pid);
        bool takeover = false;
        ++trials;
        if (owner == getpid() || (owner != 0 && !pid_alive(owner))) takeover = true;
        else if (owner == 0 && trials >= 30) takeover = true;    // 30 timeouts (5 min) like the reference: a holder that
                                                                 // never recorded itself died inside a microsecond window
        if (takeover) {
            if (lockf(fd_, F_LOCK, VGPU_REGION_SIZE) == 0) {
                int32_t now = static_cast<int32_t>(r_->owner_pid);
                bool still = (now == owner);
                // owner 0 is also the normal state between unlock()'s store and its sem_post: if the semaphore is
                // available after all, take it the ordinary way so that the count can never reach 2
                int sv = 0;
                if (still && owner == 0 && sem_getvalue(sem_of(r_), &sv) == 0 && sv > 0) still = false;
                if (still) r_->owner_pid = static_cast<uint64_t>(getpid());
                if (lockf(fd_, F_ULOCK, VGPU_REGION_SIZE) != 0) {}
                if (still) {
                    LOG_WARN("region lock taken over from dead owner %d", owner);
                    return;  // lock held WITHOUT a sem_wait: the dead owner's count is ours now
                }
            }
        }
    }
}

void Region::unlock() {
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    r_->owner_pid = 0;
    sem_post(sem_of(r_));
}

int Region::find_slot_locked(int32_t pid) {
    int n = r_->proc_num;
    if (cached_slot_ >= 0 && cached_slot_ < n && r_->procs[cached_slot_].pid == pid) return cached_slot_;
    for (int i = 0; i < n; i++)
        if (r_->procs[i].pid == pid) return cached_slot_ = i;
    return -1;
}

int Region::claim_slot(int32_t pid) {
    lock();
    int s = find_slot_locked(pid);
    if (s < 0) {
        if (r_->proc_num >= VGPU_MAX_PROCS) reap_dead_locked();
        if (r_->proc_num < VGPU_MAX_PROCS) {
            s = r_->proc_num;
            std::memset(&r_->procs[s], 0, sizeof(vgpu_proc_slot_t));
            r_->procs[s].pid = pid;
            r_->procs[s].hostpid = 0;
            r_->procs[s].status = VGPU_STATUS_RUNNING;
            __atomic_thread_fence(__ATOMIC_RELEASE);
            r_->proc_num = s + 1;
            cached_slot_ = s;
        }
    } else {
        r_->procs[s].status = VGPU_STATUS_RUNNING;
    }
    unlock();
    return s;
}

void Region::release_slot(int32_t pid) {
    lock();
    int s = find_slot_locked(pid);
    if (s >= 0) {
        int last = r_->proc_num - 1;
        std::memset(&r_->procs[s], 0, sizeof(vgpu_proc_slot_t));
        r_->proc_num = last;
        if (s != last) {
            std::memcpy(&r_->procs[s], &r_->procs[last], sizeof(vgpu_proc_slot_t));
            std::memset(&r_->procs[last], 0, sizeof(vgpu_proc_slot_t));
        }
        cached_slot_ = -1;
    }
    clear_swap_records_locked(pid);
    unlock();
}

void Region::set_hostpid(int32_t pid, int32_t hostpid) {
    lock();
    int s = find_slot_locked(pid);
    if (s >= 0) r_->procs[s].hostpid = hostpid;
    unlock();
}

uint64_t Region::usage_locked(int dev) const {
    uint64_t sum = 0;
    int n = r_->proc_num;
    for (int i = 0; i < n; i++) sum += r_->procs[i].used[dev].total;
    return sum;  // reference adds initial_offset (a process-local constant 0 in the shipped binary)
}

uint64_t Region::usage(int dev) {
    lock();
    uint64_t u = usage_locked(dev);
    unlock();
    return u;
}

// init_proc_slot_withlock@0x43d89 ends with clear_proc_slot_nolock(pid, 1)@0x43f42: every process that joins the container
// through the hook sweeps the slots of processes that died without their exit handler (SIGKILL, crash) — their bytes stop
// counting against the quota as soon as a sibling starts, not only at the next quota breach (rm_quitted_process). Same
// compaction (last slot moved into the hole), so the caller's own slot may move. Called by the hook's initialisation
// right after claim_slot(); host tools that claim slots on behalf of other pids (vgpu_region_claim) do not sweep.
int Region::sweep_dead(int32_t keep) {
    lock();
    int n = reap_dead_locked(keep);
    unlock();
    return n;
}

int Region::reap_dead_locked(int32_t keep) {
    // rm_quitted_process@0x41a8e runs `ps ax` through popen on every quota breach; /proc/<pid> answers the same
    // question without forking a shell from inside a CUDA allocation call.
    int reaped = 0;
    int32_t self = getpid();
    for (int i = 0; i < r_->proc_num;) {
        int32_t p = r_->procs[i].pid;
        if (p != self && p != keep && !pid_alive(p)) {
            clear_swap_records_locked(p);
            int last = r_->proc_num - 1;
            if (i != last) std::memcpy(&r_->procs[i], &r_->procs[last], sizeof(vgpu_proc_slot_t));
            std::memset(&r_->procs[last], 0, sizeof(vgpu_proc_slot_t));
            r_->proc_num = last;
            cached_slot_ = -1;
            reaped++;
        } else {
            i++;
        }
    }
    return reaped;
}

static uint64_t *lane(vgpu_device_memory_t &m, int type) {
    switch (type) {
        case VGPU_MEM_CONTEXT: return &m.context_size;
        case VGPU_MEM_MODULE: return &m.module_size;
        case VGPU_MEM_BUFFER: return &m.buffer_size;
        default: return nullptr;
    }
}

bool Region::try_add_fixed(int32_t pid, int dev, uint64_t bytes, uint64_t own_live, uint64_t *fixed_after) {
    lock();
    uint64_t lim = r_->limit[dev];
    int64_t fixed = (int64_t)(usage_locked(dev) - own_live - swap_live(dev, pid));
    if (fixed < 0) fixed = 0;
    if (fixed_after) *fixed_after = (uint64_t)fixed + bytes;
    if (lim && (uint64_t)fixed + bytes > lim) {
        bool ok = false;
        if (reap_dead_locked() > 0) {
            fixed = (int64_t)(usage_locked(dev) - own_live - swap_live(dev, pid));
            if (fixed < 0) fixed = 0;
            ok = !((uint64_t)fixed + bytes > lim);
            if (fixed_after) *fixed_after = (uint64_t)fixed + bytes;
        }
        if (!ok) { unlock(); return false; }
    }
    int s = find_slot_locked(pid);
    if (s >= 0) {
        vgpu_device_memory_t &m = r_->procs[s].used[dev];
        m.total += bytes;
        if (uint64_t *l = lane(m, VGPU_MEM_BUFFER)) *l += bytes;
    }
    unlock();
    return true;
}

bool Region::try_add(int32_t pid, int dev, uint64_t bytes, int type, bool enforce, bool check_only) {
    lock();
    if (enforce) {
        uint64_t lim = r_->limit[dev];
        if (lim != 0) {
            uint64_t u = usage_locked(dev);
            if (u + bytes > lim) {  // strict: usage+size == limit is admitted (ja @0x3fb88)
                bool ok = false;
                if (reap_dead_locked() > 0) {
                    u = usage_locked(dev);
                    ok = !(u + bytes > lim);
                }
                if (!ok) {
                    unlock();
                    LOG_ERROR("Device %d OOM %lu / %lu", dev, (unsigned long)(u + bytes), (unsigned long)lim);
                    return false;
                }
            }
        }
    }
    if (!check_only) {
        int s = find_slot_locked(pid);
        if (s >= 0) {
            vgpu_device_memory_t &m = r_->procs[s].used[dev];
            m.total += bytes;
            if (uint64_t *l = lane(m, type)) *l += bytes;
        } else {
            LOG_WARN("add: no slot for pid %d", pid);
        }
    }
    unlock();
    return true;
}

void Region::sub(int32_t pid, int dev, uint64_t bytes, int type) {
    lock();
    int s = find_slot_locked(pid);
    if (s >= 0) {
        vgpu_device_memory_t &m = r_->procs[s].used[dev];
        m.total -= bytes;  // wrapping u64, like the reference
        if (uint64_t *l = lane(m, type)) *l -= bytes;
    }
    unlock();
}

}  // namespace vgpu
