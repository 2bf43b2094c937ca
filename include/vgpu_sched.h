/*
 * vgpu_sched.h — C ABI of the scheduler-extender scoring core (the caller that produces the
 * annotations Allocate consumes).
 *
 * Reference (Go, no toolchain here): pkg/scheduler/score.go:36-226 (device ordering, checkType, fitInCertainDevice,
 * fitInDevices, calcScore), pkg/scheduler/scheduler.go:250-313 (getNodesUsage), pkg/device/nvidia/device.go:69-118
 * (checkGPUtype / assertNuma / CheckType). The logic lives in C++ (csrc/sched_core.cc); the HTTP/JSON transport of
 * /filter /bind /webhook (pkg/scheduler/routes/route.go:41-134) is k8s-device-plugin_b200/plugin/scheduler.py, and a Go
 * extender binds these entry points with cgo exactly like INTEGRATION.md shows for the plugin core.
 *
 * All memory/core quantities are the reference's units: MiB and percent, int32.
 */
#ifndef VGPU_SCHED_H
#define VGPU_SCHED_H
#include <stddef.h>
#include <stdint.h>

#include "vgpu_plugin.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgpu_device_usage {          /* util.DeviceUsage, pkg/util/types.go:110-122 */
    char id[VGPU_PLUGIN_MAX_STR];
    char type[VGPU_PLUGIN_MAX_STR];
    uint32_t index;
    int32_t used, count, usedmem, totalmem, totalcore, usedcores, numa, health;
} vgpu_device_usage_t;

typedef struct vgpu_device_request {        /* util.ContainerDeviceRequest, pkg/util/types.go:93-99 */
    int32_t nums;
    char type[VGPU_PLUGIN_MAX_STR];
    int32_t memreq;                         /* MiB; 0 = use the percentage */
    int32_t mem_percentage_req;             /* 101 = unset (device.go:143) */
    int32_t coresreq;
} vgpu_device_request_t;

typedef struct vgpu_sched_annotations {     /* the pod annotations the NVIDIA CheckType reads; NULL = key absent */
    const char *use_gputype;                /* nvidia.com/use-gputype   (device.go:21) */
    const char *nouse_gputype;              /* nvidia.com/nouse-gputype (device.go:22) */
    const char *numa_bind;                  /* nvidia.com/numa-bind     (device.go:23) */
} vgpu_sched_annotations_t;

typedef struct vgpu_sched_assignment {      /* util.ContainerDevice incl. Idx (index into the node's sorted device list) */
    int32_t idx;
    int32_t container;                      /* container index this device was given to */
    vgpu_container_device_t dev;
} vgpu_sched_assignment_t;

/* sort.Sort(DeviceUsageList) (score.go:36-49): ascending (numa, count-used). Stable — identical to Go's result for the
 * <= 12 devices a node carries (Go's pdqsort is an insertion sort below 13 elements). */
void vgpu_sched_sort_devices(vgpu_device_usage_t *devs, int n);

/* checkType (score.go:72-85) for the NVIDIA vendor: *pass / *numa_assert as the reference returns them; returns 1 when
 * the request type is recognised, 0 otherwise (then both outputs are 0). */
int vgpu_sched_check_type(const vgpu_sched_annotations_t *annos, const vgpu_device_usage_t *d, const vgpu_device_request_t *req,
                          int *pass, int *numa_assert);

/* fitInCertainDevice (score.go:87-161): walk the (already sorted) list from the END. Returns 1 fit / 0 no fit;
 * out[0..*n_out) are the devices picked so far either way. Does not modify devs. */
int vgpu_sched_fit_in_certain_device(const vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *req,
                                     const vgpu_sched_annotations_t *annos, vgpu_sched_assignment_t *out, int cap, int *n_out);

/* fitInDevices (score.go:163-195) for ONE container carrying n_req requests (one per device vendor; one here): sorts
 * devs, fits, and on success charges the picked devices (used++, usedcores+=, usedmem+=) exactly like the reference.
 * Returns 1 fit / 0 no fit; *score = float32(total)/float32(free) + float32(len(devs) - sum(nums)). */
int vgpu_sched_fit_in_devices(vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *reqs, int n_req,
                              const vgpu_sched_annotations_t *annos, vgpu_sched_assignment_t *out, int cap, int *n_out, float *score);

/* calcScore's per-node body (score.go:197-226): reqs[c] is container c's request (nums == 0: the container asks for no
 * device). mode 0 = reference behaviour, including its container bookkeeping: a node is kept only when
 * len(score.devices) == len(nums) — the number of device VENDORS against the number of CONTAINERS — so only
 * single-container pods ever fit, and a device-less container after a fitted one indexes past the slice (the Go runtime
 * panics; here: return -5). mode 1 = per-container bookkeeping (what later upstream releases do): the node fits when
 * every requesting container fits, device-less containers get an empty slot.
 * Returns 1 node fits, 0 not, <0 error. out/n_out: every assignment with its container index. */
int vgpu_sched_score_node(vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *reqs, int n_ctrs,
                          const vgpu_sched_annotations_t *annos, int mode, vgpu_sched_assignment_t *out, int cap, int *n_out,
                          float *score);

/* getNodesUsage's inner accumulation (scheduler.go:277-296): charge one scheduled pod's devices to a node's usage list
 * (every device whose id matches: used++, usedmem+=, usedcores+=). */
void vgpu_sched_charge(vgpu_device_usage_t *devs, int n, const vgpu_container_device_t *pod_devs, int n_pod_devs);

#ifdef __cplusplus
}
#endif
#endif
