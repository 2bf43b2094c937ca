/*
 * vgpu_region.h — byte layout of the per-container shared region (the ".cache" file).
 *
 * This is a wire format, not an implementation detail: the reference's node monitor mmaps the very same file
 * and reads/writes it without taking the lock (reference: cmd/vGPUmonitor/cudevshr.go:18-58 is the Go mirror,
 * metrics.go:117-133 reads limit/used, feedback.go:197-255 writes recentKernel/utilizationSwitch/hostpid).
 * Offsets were recovered from the shipped hook binary (lib/nvidia/libvgpu.so; SURVEY.md Appendix A cites the
 * instruction addresses) and are pinned below with static assertions, so a layout drift is a compile error
 * instead of silently wrong Prometheus numbers (the reference's Go mirror has no such check).
 */
#ifndef VGPU_REGION_H
#define VGPU_REGION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGPU_REGION_MAGIC 19920718            /* cudevshr.go:15; cmp 0x12ff74e libvgpu.so@0x4460e */
#define VGPU_MAX_DEVICES 16                   /* cudevshr.go:16 */
#define VGPU_MAX_PROCS 1024                   /* cudevshr.go:52 */
#define VGPU_UUID_LEN 96                      /* cudevshr.go:35 */
#define VGPU_REGION_SIZE 0xC4748u             /* lseek/mmap/lockf constant libvgpu.so@0x44371 */

#define VGPU_STATUS_RUNNING 1                 /* procs[i].status */
#define VGPU_STATUS_SWAPPED 2

/* accounting "type" argument of add/rm usage (add_gpu_device_memory_usage libvgpu.so@0x42a0e) */
#define VGPU_MEM_CONTEXT 0
#define VGPU_MEM_MODULE 1
#define VGPU_MEM_BUFFER 2

typedef struct vgpu_device_memory {           /* cudevshr.go:18-24, stride 40 */
    uint64_t context_size;
    uint64_t module_size;
    uint64_t buffer_size;
    uint64_t offset;
    uint64_t total;
} vgpu_device_memory_t;

typedef struct vgpu_proc_slot {               /* cudevshr.go:26-32, stride 0x310 */
    int32_t pid;
    int32_t hostpid;
    vgpu_device_memory_t used[VGPU_MAX_DEVICES];
    uint64_t monitorused[VGPU_MAX_DEVICES];
    int32_t status;
    int32_t _pad;
} vgpu_proc_slot_t;

typedef struct vgpu_shared_region {           /* cudevshr.go:42-58 */
    int32_t initialized_flag;                 /* == VGPU_REGION_MAGIC once set up */
    int32_t sm_init_flag;
    uint64_t owner_pid;                       /* lock owner (C writes 8 bytes, libvgpu.so@0x43845) */
    unsigned char sem[32];                    /* sem_t, pshared=1, value 1 (sem_init libvgpu.so@0x44665) */
    uint64_t device_num;
    char uuids[VGPU_MAX_DEVICES][VGPU_UUID_LEN];
    uint64_t limit[VGPU_MAX_DEVICES];         /* bytes; 0 = unlimited */
    uint64_t sm_limit[VGPU_MAX_DEVICES];      /* percent; 0 or >=100 = unlimited */
    vgpu_proc_slot_t procs[VGPU_MAX_PROCS];
    int32_t proc_num;
    int32_t utilization_switch;               /* written by the monitor (feedback.go:236-251) */
    int32_t recent_kernel;                    /* hook sets 2 on launch; monitor decrements / sets -1 to block */
    int32_t priority;                         /* CUDA_TASK_PRIORITY, default 1 */
} vgpu_shared_region_t;

#ifdef __cplusplus
#define VGPU_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define VGPU_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
VGPU_STATIC_ASSERT(sizeof(vgpu_device_memory_t) == 40, "used[] stride (Appendix A)");
VGPU_STATIC_ASSERT(sizeof(vgpu_proc_slot_t) == 0x310, "slot stride imul 0x310 @0x42af2");
VGPU_STATIC_ASSERT(offsetof(vgpu_proc_slot_t, used) == 0x8, "used lane");
VGPU_STATIC_ASSERT(offsetof(vgpu_proc_slot_t, monitorused) == 0x288, "monitorused lane @0x42439");
VGPU_STATIC_ASSERT(offsetof(vgpu_proc_slot_t, status) == 0x308, "status @0x40ca2");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, owner_pid) == 0x8, "owner_pid @0x43845");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, sem) == 0x10, "sem @0x44665");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, device_num) == 0x30, "device_num");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, uuids) == 0x38, "uuids");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, limit) == 0x638, "limit @0x44620");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, sm_limit) == 0x6B8, "sm_limit @0x4463a");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, procs) == 0x738, "procs @0x42af2");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, proc_num) == 0xC4738, "proc_num @0x42d23");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, utilization_switch) == 0xC473C, "utilization_switch @0x446e0");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, recent_kernel) == 0xC4740, "recent_kernel @0x446f1");
VGPU_STATIC_ASSERT(offsetof(vgpu_shared_region_t, priority) == 0xC4744, "priority @0x44702");
VGPU_STATIC_ASSERT(sizeof(vgpu_shared_region_t) == VGPU_REGION_SIZE, "region file size");

/*
 * Extension block (not in the reference): swap-engine counters of the container, for the node monitor (SURVEY.md §8(f) #2
 * "extend with swap counters via the new C ABI"). It lives in the SAME file, page-aligned behind the reference's region,
 * so the directory contract of the reference monitor (at most two entries, checkfiles pathmonitor.go:38-71) and every
 * reference offset above stay untouched; consumers that map VGPU_REGION_SIZE bytes never see it. One record per
 * (process, device) that runs a swap engine; the owning process overwrites its record with relaxed stores, readers sum
 * the records of a device. Records of exited processes are cleared together with their process slot.
 */
#define VGPU_REGION_EXT_OFFSET 0xC5000u
#define VGPU_REGION_EXT_MAGIC 0x30303242u      /* "B200" */
#define VGPU_REGION_EXT_RECORDS 1024
typedef struct vgpu_swap_record {
    int32_t pid;                               /* 0 = free */
    int32_t dev;
    uint64_t page_out_bytes, page_in_bytes;    /* cumulative */
    uint64_t evictions, faults;                /* cumulative */
    uint64_t resident_bytes, live_bytes, host_bytes;   /* gauges */
} vgpu_swap_record_t;
typedef struct vgpu_region_ext {
    uint32_t magic, version;
    uint64_t reserved[7];
    vgpu_swap_record_t swap[VGPU_REGION_EXT_RECORDS];
} vgpu_region_ext_t;
VGPU_STATIC_ASSERT(sizeof(vgpu_swap_record_t) == 64, "swap record");
VGPU_STATIC_ASSERT(VGPU_REGION_EXT_OFFSET >= VGPU_REGION_SIZE && VGPU_REGION_EXT_OFFSET % 4096 == 0, "extension placement");

#ifdef __cplusplus
}
#endif
#endif /* VGPU_REGION_H */
