/*
 * vgpu_plugin.h — C ABI of the host-side (kubelet-facing) logic of the device plugin for the vGPU path.
 *
 * The reference implements this in Go (pkg/device-plugin/nvidiadevice/nvinternal/plugin/server.go, rm/devices.go,
 * pkg/util/util.go); this image has no Go toolchain, so the logic lives in C++ (csrc/plugin_core.cc) behind plain C
 * strings. A Go device plugin binds these with cgo and keeps only the gRPC/kubelet transport; here
 * the transport for tests is k8s-device-plugin_b200/plugin/ (Python grpc speaking the kubelet v1beta1 API).
 *
 * Wire formats are the reference's, byte for byte:
 *   node annotation  4pd.io/node-nvidia-register : "<UUID>,<count>,<MiB>,<cores>,<type>,<numa>,<health>:" per GPU
 *                                                  (EncodeNodeDevices util.go:111-118, DecodeNodeDevices util.go:78-109)
 *   pod annotation   hami.sh/vgpu-devices-to-allocate : "<UUID>,<Type>,<MiB>,<cores>:" per device, then ONE ";" for
 *                                                  the whole pod (EncodePodSingleDevice util.go:142-150 — reproduced
 *                                                  as is, including the single trailing ';')
 * All functions return 0 on success, <0 on error (-1 format error, -2 buffer too small, -3 not found).
 */
#ifndef VGPU_PLUGIN_H
#define VGPU_PLUGIN_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VGPU_PLUGIN_MAX_STR 128

typedef struct vgpu_container_device {      /* util.ContainerDevice, pkg/util/types.go:85-91 */
    char uuid[VGPU_PLUGIN_MAX_STR];
    char type[VGPU_PLUGIN_MAX_STR];
    int32_t usedmem;                        /* MiB */
    int32_t usedcores;                      /* percent */
} vgpu_container_device_t;

typedef struct vgpu_node_device {           /* api.DeviceInfo, pkg/api/device_register.go:13-22 */
    char id[VGPU_PLUGIN_MAX_STR];
    int32_t count, devmem, devcore;
    char type[VGPU_PLUGIN_MAX_STR];
    int32_t numa, health;
} vgpu_node_device_t;

/* ---- annotation codec */
int vgpu_codec_encode_node_devices(const vgpu_node_device_t *devs, int n, char *buf, size_t cap);
int vgpu_codec_decode_node_devices(const char *s, vgpu_node_device_t *out, int cap, int *n);
int vgpu_codec_encode_container_devices(const vgpu_container_device_t *devs, int n, char *buf, size_t cap);
int vgpu_codec_decode_container_devices(const char *s, vgpu_container_device_t *out, int cap, int *n);
/* pod level: counts[i] devices for container i, devices flattened in order. */
int vgpu_codec_encode_pod_single_device(const vgpu_container_device_t *devs, const int *counts, int n_ctrs, char *buf, size_t cap);
int vgpu_codec_decode_pod_single_device(const char *s, vgpu_container_device_t *out, int dev_cap, int *counts, int ctr_cap, int *n_ctrs);
/* GetNextDeviceRequest (util.go:216-236): first container with a non-empty device list */
int vgpu_codec_next_device_request(const char *annotation, int *ctr_index, vgpu_container_device_t *out, int cap, int *n);
/* EraseNextDeviceTypeFromAnnotation (util.go:244-271): the annotation with that container's list emptied */
int vgpu_codec_erase_next_device_request(const char *annotation, char *buf, size_t cap);

/* ---- device fan-out: GetPluginDevices (rm/devices.go:144-167): DeviceSplitCount ids "<UUID>-<i>" per GPU */
int vgpu_plugin_device_id(const char *uuid, unsigned index, char *buf, size_t cap);
/* registeredmem (register.go:128-131): int32(MiB) scaled by DeviceMemoryScaling; devcore = scaling*100 (register.go:153) */
int32_t vgpu_plugin_registered_mem(uint64_t total_bytes, double memory_scaling);
int32_t vgpu_plugin_registered_cores(double cores_scaling);

/* ---- Allocate (server.go:288-411): the env + mount contract consumed by the in-container hook */
typedef struct vgpu_kv { char key[VGPU_PLUGIN_MAX_STR]; char value[512]; } vgpu_kv_t;
typedef struct vgpu_mount { char container_path[512]; char host_path[512]; int32_t read_only; } vgpu_mount_t;
typedef struct vgpu_allocate_in {
    const vgpu_container_device_t *devices; /* devreq of the container being allocated */
    int n_devices;
    int n_requested_ids;                    /* len(req.DevicesIDs): must equal n_devices (server.go:328) */
    const char *host_hook_path;             /* HOOK_PATH */
    const char *pod_uid;
    const char *container_name;
    const char *cache_uuid;                 /* uuid.New() of the reference; caller supplies (NULL -> derived from pod/ctr) */
    double device_memory_scaling;           /* > 1 -> CUDA_OVERSUBSCRIBE=true (server.go:356-358) */
    int disable_core_limit;                 /* --disable-core-limit -> GPU_CORE_UTILIZATION_POLICY=disable */
    int container_sets_disable_control;     /* container env has CUDA_DISABLE_CONTROL -> no ld.so.preload mount */
    int license_present;                    /* <hook>/vgpu/license exists */
    const char *device_list_envvar;         /* normally NVIDIA_VISIBLE_DEVICES */
} vgpu_allocate_in_t;
typedef struct vgpu_allocate_out {
    vgpu_kv_t envs[32]; int n_envs;
    vgpu_mount_t mounts[8]; int n_mounts;
    char cache_host_dir[512];               /* directory the plugin must create 0777 before answering */
} vgpu_allocate_out_t;
/* returns 0, or -4 "device allocate number not matched" */
int vgpu_plugin_allocate(const vgpu_allocate_in_t *in, vgpu_allocate_out_t *out);

#ifdef __cplusplus
}
#endif
#endif
