// region.h — the per-container shared region: create/attach, cross-process lock, process slots, byte accounting.
//
// Same file format and lock protocol as the reference (include/vgpu_region.h; reference functions:
// try_create_shrreg libvgpu.so@0x44138, lock_shrreg@0x437d3, unlock_shrreg@0x43bbb, fix_lock_shrreg@0x4317c,
// init_proc_slot_withlock@0x43d89, exit_handler@0x43501, add_gpu_device_memory_usage@0x42a0e,
// rm_gpu_device_memory_usage@0x42de1, get_gpu_memory_usage@0x420bd, rm_quitted_process@0x41a8e), so reference
// monitors — and reference-hooked processes of the same container — can share the file. What changes is the
// cost: the own slot index is cached and revalidated under the lock instead of a linear pid scan with a log
// macro per slot, dead-process reclaim reads /proc/<pid> instead of popen("ps ax"), and nothing calls getenv.
#pragma once
#include <cstdint>
#include <string>

#include "vgpu_region.h"

namespace vgpu {

// get_limit_from_env@0x40d00 applied to a VALUE string ("8192m" -> 8589934592); 0 = unlimited/invalid
uint64_t parse_limit(const char *value);
// CUDA_DEVICE_MEMORY_LIMIT_<i> overrides CUDA_DEVICE_MEMORY_LIMIT (do_init_device_memory_limits@0x41806)
uint64_t limit_from_env(const char *base_name, int dev);

class Region {
   public:
    // Opens (creating if needed) the region file. env-derived limits are applied by the first creator and
    // verified by later joiners ("Limit inconsistency", multiprocess_memory_limit.c:L725).
    // uuids/ndev: device identities to publish (may be null/0 when NVML is unavailable).
    static Region *open(const char *path, bool create, const uint64_t *mem_limits, const uint64_t *sm_limits,
                        int priority, const char (*uuids)[VGPU_UUID_LEN], int ndev, std::string *err);
    ~Region();

    vgpu_shared_region_t *raw() { return r_; }
    const std::string &path() const { return path_; }

    void lock();
    void unlock();

    // Claims (or re-finds) the slot of `pid`, status RUNNING. Returns slot index or -1 when all 1024 are taken.
    int claim_slot(int32_t pid);
    // exit_handler@0x43501: zero own slot, move the last slot into the hole, procnum--
    void release_slot(int32_t pid);
    void set_hostpid(int32_t pid, int32_t hostpid);

    uint64_t limit(int dev) const { return r_->limit[dev]; }
    uint64_t sm_limit(int dev) const { return r_->sm_limit[dev]; }

    // Σ over slots of used[dev].total (get_gpu_memory_usage@0x420bd). Takes the lock.
    uint64_t usage(int dev);
    // oom_check@0x3f9bc + add_gpu_device_memory_usage@0x42a0e as ONE critical section: if limit != 0 and
    // usage+bytes > limit (strict), first reclaims slots of dead pids (rm_quitted_process) and re-checks;
    // returns false on quota breach (nothing added). check_only: oom_check without the add.
    bool try_add(int32_t pid, int dev, uint64_t bytes, int type, bool enforce, bool check_only = false);
    void add(int32_t pid, int dev, uint64_t bytes, int type) { try_add(pid, dev, bytes, type, false); }
    void sub(int32_t pid, int dev, uint64_t bytes, int type);  // rm_gpu_device_memory_usage@0x42de1
    // drop slots whose pid no longer exists; returns how many were reclaimed (lock must be held)
    int reap_dead_locked(int32_t keep = 0);
    int sweep_dead(int32_t keep = 0);        // lock + reap_dead_locked: what a joining process does (clear_proc_slot_nolock@0x43bf4)

    // extension block (vgpu_region.h): nullptr when the file has none and cannot be grown
    vgpu_region_ext_t *ext() { return ext_; }
    // the (pid, dev) record, claimed on first use; nullptr when there is no extension block or it is full
    vgpu_swap_record_t *swap_record(int32_t pid, int dev);
    // sum of the live records of `dev`; out->pid = records summed. false: no extension block
    bool swap_counters(int dev, vgpu_swap_record_t *out);
    uint64_t swap_live(int dev, int32_t except_pid = 0) const;
    // swap mode, non-swappable charge: under the lock, (usage - live swappable bytes of all engines) + bytes <= limit ? add : refuse
    bool try_add_fixed(int32_t pid, int dev, uint64_t bytes, uint64_t own_live, uint64_t *fixed_after);   // live swappable bytes of the container's engines on dev (relaxed reads)
    // Shared RESIDENT budget of the container's swap engines on one device (several processes, one quota), under the region
    // lock: the room for swappable memory is limit - everything non-swappable - every engine's staging rings; of that an
    // engine may hold what the OTHERS neither hold nor are entitled to (their fair share, bounded by what they have
    // live). Grants `want_total` (publishes it as the caller's resident_bytes, i.e. reserves it) when it fits; returns the
    // caller's cap either way. live_mapped: the caller's live swappable bytes; overhead: its staging rings.
    uint64_t swap_reserve(int32_t pid, int dev, uint64_t want_total, uint64_t live_mapped, uint64_t overhead, bool *granted, int *engines, uint64_t *share = nullptr);

   private:
    Region() = default;
    int find_slot_locked(int32_t pid);
    uint64_t usage_locked(int dev) const;

    void clear_swap_records_locked(int32_t pid);

    vgpu_shared_region_t *r_ = nullptr;
    vgpu_region_ext_t *ext_ = nullptr;
    int fd_ = -1;
    std::string path_;
    int cached_slot_ = -1;
    uint64_t last_dead_sweep_ns_ = 0;           // swap_reserve: when the siblings' liveness was last checked
};

}  // namespace vgpu
