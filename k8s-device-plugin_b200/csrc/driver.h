// driver.h — table of REAL CUDA-driver / NVML entry points the library calls underneath its hooks.
//
// Replaces the reference's cuda_library_entry[196] / nvml_library_entry[243] name tables
// (libvgpu.so .data@0x60460 / @0x5f520, filled by load_cuda_libraries@0x2db00 / load_nvml_libraries@0x23b63):
// typed function pointers resolved once through the REAL dlsym, so a hooked process never recurses into its own
// wrappers. Missing symbols are left null and checked at the call site (older drivers lack e.g. cuFuncGetParamInfo).
#pragma once
#include <cuda.h>
#include <nvml.h>

namespace vgpu {

#define VGPU_DRV_FUNCS(X)                                                                                         \
    X(cuInit) X(cuDriverGetVersion) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetAttribute)                    \
    X(cuDeviceTotalMem_v2) X(cuDeviceGetUuid_v2) X(cuDeviceGetName)                                               \
    X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease_v2)                                                   \
    X(cuCtxCreate_v2) X(cuCtxGetCurrent) X(cuCtxSetCurrent) X(cuCtxPushCurrent_v2) X(cuCtxPopCurrent_v2)          \
    X(cuCtxGetDevice) X(cuCtxSynchronize)                                                                         \
    X(cuMemAlloc_v2) X(cuMemAllocManaged) X(cuMemAllocPitch_v2) X(cuMemFree_v2) X(cuMemGetInfo_v2)                \
    X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostGetDevicePointer_v2) X(cuMemHostRegister_v2)                    \
    X(cuMemAllocHost_v2) X(cuMemGetAddressRange_v2) X(cuMemHostUnregister) X(cuMipmappedArrayCreate) X(cuMipmappedArrayDestroy) X(cuPointerGetAttribute) X(cuPointerGetAttributes)                                                            \
    X(cuMemcpy2D_v2) X(cuMemcpy2DUnaligned_v2) X(cuMemcpy2DAsync_v2) X(cuMemcpy3D_v2) X(cuMemcpy3DAsync_v2)            \
    X(cuMemsetD2D8_v2) X(cuMemsetD2D16_v2) X(cuMemsetD2D32_v2) X(cuMemsetD2D8Async) X(cuMemsetD2D16Async) X(cuMemsetD2D32Async) \
    X(cuMemcpyPeer) X(cuMemcpyPeerAsync) X(cuMemcpyBatchAsync) X(cuMemcpy3DBatchAsync)                 \
    X(cuMemAddressReserve) X(cuMemAddressFree) X(cuMemCreate) X(cuMemRelease) X(cuMemMap) X(cuMemUnmap)           \
    X(cuMemSetAccess) X(cuMemGetAllocationGranularity)                                                            \
    X(cuMemcpyHtoD_v2) X(cuMemcpyDtoH_v2) X(cuMemcpyDtoD_v2) X(cuMemcpyHtoDAsync_v2) X(cuMemcpyDtoHAsync_v2)      \
    X(cuMemcpyDtoDAsync_v2) X(cuMemcpy) X(cuMemcpyAsync)                                                          \
    X(cuMemsetD8_v2) X(cuMemsetD16_v2) X(cuMemsetD32_v2) X(cuMemsetD8Async) X(cuMemsetD16Async)                   \
    X(cuMemsetD32Async)                                                                                           \
    X(cuStreamCreate) X(cuStreamCreateWithPriority) X(cuCtxGetStreamPriorityRange) X(cuStreamDestroy_v2) X(cuStreamSynchronize) X(cuStreamWaitEvent) X(cuStreamQuery) X(cuStreamIsCapturing)          \
    X(cuEventCreate) X(cuEventRecord) X(cuEventSynchronize) X(cuEventQuery) X(cuEventElapsedTime)                 \
    X(cuEventDestroy_v2)                                                                                          \
    X(cuModuleLoadData) X(cuModuleGetFunction) X(cuModuleUnload) X(cuFuncSetAttribute) X(cuFuncGetParamInfo)      \
    X(cuLaunchKernel) X(cuLaunchKernelEx) X(cuLaunchCooperativeKernel) X(cuOccupancyMaxActiveBlocksPerMultiprocessor) \
    X(cuMemAllocAsync) X(cuMemAllocFromPoolAsync) X(cuMemFreeAsync) X(cuGraphLaunch)                              \
    X(cuMemAdvise) X(cuMemAdvise_v2) X(cuMemPrefetchAsync) X(cuMemPrefetchAsync_v2)                                                      \
    X(cuGraphAddKernelNode_v2) X(cuGraphKernelNodeSetParams_v2) X(cuGraphExecKernelNodeSetParams_v2)              \
    X(cuGraphAddMemcpyNode) X(cuGraphMemcpyNodeSetParams) X(cuGraphExecMemcpyNodeSetParams) X(cuGraphAddMemsetNode) \
    X(cuGraphAddNode) X(cuGraphAddNode_v2)                                                                        \
    X(cuGetProcAddress_v2) X(cuGetErrorString) X(cuGetErrorName)

// per-thread-default-stream twins (cuda.h hides their prototypes behind __CUDA_API_VERSION_INTERNAL; the signatures
// are those of the base name). cudart built with --default-stream per-thread asks cuGetProcAddress for these.
#define VGPU_DRV_PT_FUNCS(X)                                                                                      \
    X(cuLaunchKernel, _ptsz) X(cuLaunchKernelEx, _ptsz) X(cuLaunchCooperativeKernel, _ptsz) X(cuGraphLaunch, _ptsz) \
    X(cuMemAllocAsync, _ptsz) X(cuMemAllocFromPoolAsync, _ptsz) X(cuMemFreeAsync, _ptsz)                          \
    X(cuMemcpyHtoD_v2, _ptds) X(cuMemcpyDtoH_v2, _ptds) X(cuMemcpyDtoD_v2, _ptds) X(cuMemcpy, _ptds)              \
    X(cuMemcpyHtoDAsync_v2, _ptsz) X(cuMemcpyDtoHAsync_v2, _ptsz) X(cuMemcpyDtoDAsync_v2, _ptsz)                  \
    X(cuMemcpyAsync, _ptsz) X(cuMemcpyBatchAsync, _ptsz) X(cuMemcpy3DBatchAsync, _ptsz)                           \
    X(cuMemPrefetchAsync, _ptsz) X(cuMemPrefetchAsync_v2, _ptsz)                                                  \
    X(cuMemsetD8_v2, _ptds) X(cuMemsetD16_v2, _ptds) X(cuMemsetD32_v2, _ptds)                                     \
    X(cuMemsetD8Async, _ptsz) X(cuMemsetD16Async, _ptsz) X(cuMemsetD32Async, _ptsz)

#define VGPU_NVML_FUNCS(X)                                                                                        \
    X(nvmlInit_v2) X(nvmlShutdown) X(nvmlDeviceGetCount_v2) X(nvmlDeviceGetHandleByIndex_v2) X(nvmlDeviceGetUUID) \
    X(nvmlDeviceGetMemoryInfo) X(nvmlDeviceGetMemoryInfo_v2) X(nvmlDeviceGetComputeRunningProcesses_v3)           \
    X(nvmlDeviceGetProcessUtilization) X(nvmlDeviceGetIndex) X(nvmlErrorString) X(nvmlDeviceGetHandleByUUID)

struct DriverTable {
#define X(name) decltype(&::name) name = nullptr;
    VGPU_DRV_FUNCS(X)
#undef X
#define X(name, sfx) decltype(&::name) name##sfx = nullptr;
    VGPU_DRV_PT_FUNCS(X)
#undef X
    // cuGetProcAddress (v1 signature, CUDA 11.3-11.8 entry point) has no prototype in cuda.h 12.x
    CUresult (*cuGetProcAddress_v1)(const char *, void **, int, cuuint64_t) = nullptr;
    void *handle = nullptr;
    bool loaded = false;
};
struct NvmlTable {
#define X(name) decltype(&::name) name = nullptr;
    VGPU_NVML_FUNCS(X)
#undef X
    void *handle = nullptr;
    bool loaded = false;
    bool inited = false;
};

// dlsym of the C library itself, immune to this library's own exported dlsym override
void *real_dlsym(void *handle, const char *name);
// lazily dlopen("libcuda.so.1") and fill the table; returns table with loaded=false when there is no driver
const DriverTable &drv();
// lazily dlopen("libnvidia-ml.so.1"); nvml_ready() additionally runs nvmlInit_v2 once
const NvmlTable &nvml();
bool nvml_ready();
// any other real driver symbol by name (pass-through resolution for dlsym / cuGetProcAddress hooks)
void *real_cuda_symbol(const char *name);
void *real_nvml_symbol(const char *name);

const char *cu_err(CUresult r);

}  // namespace vgpu
