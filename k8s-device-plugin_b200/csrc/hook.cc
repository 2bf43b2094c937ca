// hook.cc — the LD_PRELOAD surface of libvgpu.so: every symbol an application (or libcudart, through dlsym /
// cuGetProcAddress) can reach and that carries vGPU semantics.
//
// Reference boundary: lib/nvidia/libvgpu.so is injected into every process of the container via
// /etc/ld.so.preload (pkg/device-plugin/nvidiadevice/nvinternal/plugin/server.go:386-391, lib/nvidia/ld.so.preload:1)
// and re-routes symbol resolution three ways — its own exported cu*/nvml* symbols, a dlsym override
// (libvgpu.so@0x11b36 -> __dlsym_hook_section@0x11e8e / _nvml@0x13ab8) and cuGetProcAddress{,_v2}
// (@0x2e4ad / @0x2e9c8 via find_symbols_in_table@0x2e172). The same three routes exist here. The reference wraps
// 205 cu* + 246 nvml* names, nearly all of them pure log-and-forward; only names with semantics are wrapped here,
// everything else resolves straight to the real library (nothing to forward, nothing to log on the hot path).
#include <cuda.h>
#include <dlfcn.h>
#include <nvml.h>
#undef cuGetProcAddress   // cuda.h maps the unsuffixed name onto _v2; both entry points are exported here

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "driver.h"
#include "log.h"
#include "runtime.h"

using vgpu::drv;
using vgpu::Runtime;

#define VGPU_EXPORT extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------------ hooked driver API
VGPU_EXPORT CUresult cuInit(unsigned int flags) { return Runtime::get().init(flags); }

VGPU_EXPORT CUresult cuMemAlloc_v2(CUdeviceptr *dptr, size_t bytesize) { return Runtime::get().mem_alloc(dptr, bytesize); }
VGPU_EXPORT CUresult cuMemAllocManaged(CUdeviceptr *dptr, size_t bytesize, unsigned int flags) {
    return Runtime::get().mem_alloc_managed(dptr, bytesize, flags);
}
VGPU_EXPORT CUresult cuMemAllocPitch_v2(CUdeviceptr *dptr, size_t *pPitch, size_t WidthInBytes, size_t Height,
                                        unsigned int ElementSizeBytes) {
    return Runtime::get().mem_alloc_pitch(dptr, pPitch, WidthInBytes, Height, ElementSizeBytes);
}
VGPU_EXPORT CUresult cuMemFree_v2(CUdeviceptr dptr) { return Runtime::get().mem_free(dptr); }
VGPU_EXPORT CUresult cuMemGetInfo_v2(size_t *free_b, size_t *total) { return Runtime::get().mem_get_info(free_b, total); }
VGPU_EXPORT CUresult cuDeviceTotalMem_v2(size_t *bytes, CUdevice dev) { return Runtime::get().device_total_mem(bytes, dev); }
VGPU_EXPORT CUresult cuDevicePrimaryCtxRetain(CUcontext *pctx, CUdevice dev) { return Runtime::get().primary_ctx_retain(pctx, dev); }
VGPU_EXPORT CUresult cuCtxCreate_v2(CUcontext *pctx, unsigned int flags, CUdevice dev) { return Runtime::get().ctx_create(pctx, flags, dev); }

// host allocations only run the quota check (reference: cuMemHostAlloc@0x32577 / cuMemAllocHost_v2@0x319f8 /
// cuMemHostRegister_v2@0x32842: real call, then check_oom(); on breach undo and return CUDA_ERROR_OUT_OF_MEMORY)
VGPU_EXPORT CUresult cuMemHostAlloc(void **pp, size_t bytesize, unsigned int flags) {
    Runtime &rt = Runtime::get();
    rt.ensure_initialized();
    CUresult r = drv().cuMemHostAlloc(pp, bytesize, flags);
    if (r == CUDA_SUCCESS && rt.check_oom()) { drv().cuMemFreeHost(*pp); *pp = nullptr; return CUDA_ERROR_OUT_OF_MEMORY; }
    return r;
}
VGPU_EXPORT CUresult cuMemAllocHost_v2(void **pp, size_t bytesize) {
    Runtime &rt = Runtime::get();
    rt.ensure_initialized();
    CUresult r = drv().cuMemAllocHost_v2(pp, bytesize);
    if (r == CUDA_SUCCESS && rt.check_oom()) { drv().cuMemFreeHost(*pp); return CUDA_ERROR_OUT_OF_MEMORY; }
    return r;
}

VGPU_EXPORT CUresult cuMemHostRegister_v2(void *p, size_t bytesize, unsigned int flags) {   // cuMemHostRegister_v2@0x32842
    Runtime &rt = Runtime::get();
    rt.ensure_initialized();
    if (!drv().cuMemHostRegister_v2) return CUDA_ERROR_NOT_SUPPORTED;
    CUresult r = drv().cuMemHostRegister_v2(p, bytesize, flags);
    if (r == CUDA_SUCCESS && rt.check_oom()) { if (drv().cuMemHostUnregister) drv().cuMemHostUnregister(p); return CUDA_ERROR_OUT_OF_MEMORY; }
    return r;
}
VGPU_EXPORT CUresult cuMipmappedArrayCreate(CUmipmappedArray *h, const CUDA_ARRAY3D_DESCRIPTOR *desc, unsigned int levels) {   // @0x36d29
    Runtime &rt = Runtime::get();
    rt.ensure_initialized();
    if (!drv().cuMipmappedArrayCreate) return CUDA_ERROR_NOT_SUPPORTED;
    CUresult r = drv().cuMipmappedArrayCreate(h, desc, levels);
    if (r == CUDA_SUCCESS && rt.check_oom()) { if (drv().cuMipmappedArrayDestroy) drv().cuMipmappedArrayDestroy(*h); return CUDA_ERROR_OUT_OF_MEMORY; }
    return r;
}

VGPU_EXPORT CUresult cuLaunchKernel(CUfunction f, unsigned int gridDimX, unsigned int gridDimY, unsigned int gridDimZ,
                                    unsigned int blockDimX, unsigned int blockDimY, unsigned int blockDimZ,
                                    unsigned int sharedMemBytes, CUstream hStream, void **kernelParams, void **extra) {
    return Runtime::get().launch_kernel(f, gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ, sharedMemBytes,
                                        hStream, kernelParams, extra);
}
VGPU_EXPORT CUresult cuLaunchKernelEx(const CUlaunchConfig *config, CUfunction f, void **kernelParams, void **extra) {
    return Runtime::get().launch_kernel_ex(config, f, kernelParams, extra);
}
VGPU_EXPORT CUresult cuLaunchCooperativeKernel(CUfunction f, unsigned int gridDimX, unsigned int gridDimY,
                                               unsigned int gridDimZ, unsigned int blockDimX, unsigned int blockDimY,
                                               unsigned int blockDimZ, unsigned int sharedMemBytes, CUstream hStream,
                                               void **kernelParams) {
    return Runtime::get().launch_cooperative(f, gridDimX, gridDimY, gridDimZ, blockDimX, blockDimY, blockDimZ,
                                             sharedMemBytes, hStream, kernelParams);
}

// per-thread-default-stream twins: same semantics, the driver's _ptsz entry underneath
VGPU_EXPORT CUresult cuLaunchKernel_ptsz(CUfunction f, unsigned int gx, unsigned int gy, unsigned int gz, unsigned int bx,
                                         unsigned int by, unsigned int bz, unsigned int smem, CUstream st, void **params, void **extra) {
    return Runtime::get().launch_kernel_ptsz(f, gx, gy, gz, bx, by, bz, smem, st, params, extra);
}
VGPU_EXPORT CUresult cuLaunchKernelEx_ptsz(const CUlaunchConfig *config, CUfunction f, void **params, void **extra) {
    return Runtime::get().launch_kernel_ex_ptsz(config, f, params, extra);
}
VGPU_EXPORT CUresult cuLaunchCooperativeKernel_ptsz(CUfunction f, unsigned int gx, unsigned int gy, unsigned int gz, unsigned int bx,
                                                    unsigned int by, unsigned int bz, unsigned int smem, CUstream st, void **params) {
    return Runtime::get().launch_cooperative_ptsz(f, gx, gy, gz, bx, by, bz, smem, st, params);
}

// beyond the reference's coverage: graph launches are rate-limited, stream-ordered and VMM
// allocations are charged to the quota
VGPU_EXPORT CUresult cuGraphLaunch(CUgraphExec g, CUstream st) { return Runtime::get().graph_launch(g, st, false); }
VGPU_EXPORT CUresult cuGraphLaunch_ptsz(CUgraphExec g, CUstream st) { return Runtime::get().graph_launch(g, st, true); }
VGPU_EXPORT CUresult cuMemAllocAsync(CUdeviceptr *dptr, size_t bytesize, CUstream st) {
    return Runtime::get().mem_alloc_async(dptr, bytesize, nullptr, false, st, false);
}
VGPU_EXPORT CUresult cuMemAllocAsync_ptsz(CUdeviceptr *dptr, size_t bytesize, CUstream st) {
    return Runtime::get().mem_alloc_async(dptr, bytesize, nullptr, false, st, true);
}
VGPU_EXPORT CUresult cuMemAllocFromPoolAsync(CUdeviceptr *dptr, size_t bytesize, CUmemoryPool pool, CUstream st) {
    return Runtime::get().mem_alloc_async(dptr, bytesize, pool, true, st, false);
}
VGPU_EXPORT CUresult cuMemAllocFromPoolAsync_ptsz(CUdeviceptr *dptr, size_t bytesize, CUmemoryPool pool, CUstream st) {
    return Runtime::get().mem_alloc_async(dptr, bytesize, pool, true, st, true);
}
VGPU_EXPORT CUresult cuMemFreeAsync(CUdeviceptr dptr, CUstream st) { return Runtime::get().mem_free_async(dptr, st, false); }
VGPU_EXPORT CUresult cuMemFreeAsync_ptsz(CUdeviceptr dptr, CUstream st) { return Runtime::get().mem_free_async(dptr, st, true); }
VGPU_EXPORT CUresult cuMemCreate(CUmemGenericAllocationHandle *handle, size_t size, const CUmemAllocationProp *prop, unsigned long long flags) {
    return Runtime::get().mem_create(handle, size, prop, flags);
}
VGPU_EXPORT CUresult cuMemRelease(CUmemGenericAllocationHandle handle) { return Runtime::get().mem_release(handle); }
VGPU_EXPORT CUresult cuMemMap(CUdeviceptr ptr, size_t size, size_t offset, CUmemGenericAllocationHandle handle, unsigned long long flags) {
    return Runtime::get().mem_map(ptr, size, offset, handle, flags);
}
VGPU_EXPORT CUresult cuMemUnmap(CUdeviceptr ptr, size_t size) { return Runtime::get().mem_unmap(ptr, size); }

VGPU_EXPORT CUresult cuModuleUnload(CUmodule hmod) {
    Runtime::get().forget_function_layouts();
    return drv().cuModuleUnload(hmod);
}

// memcpy / memset: pass-through, except that in swap mode the device ranges they touch are paged in first (a DMA
// engine cannot fault on an unmapped VMM range the way UVM-managed memory does in the reference)
#define TOUCH1(p, n, st) Runtime::get().touch_range((p), (n), (st))
#define TOUCH1R(p, n, st) Runtime::get().touch_range((p), (n), (st), /*writes=*/false)   /* the range is only read */
#define TOUCH2(a, an, b, bn, st) Runtime::get().touch_range2((a), (an), (b), (bn), (st))
#define TOUCH_DONE(st) Runtime::get().touch_done(st)
VGPU_EXPORT CUresult cuMemcpyHtoD_v2(CUdeviceptr dst, const void *src, size_t n) {
    TOUCH1(dst, n, nullptr); CUresult r = drv().cuMemcpyHtoD_v2(dst, src, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoH_v2(void *dst, CUdeviceptr src, size_t n) {
    TOUCH1R(src, n, nullptr); CUresult r = drv().cuMemcpyDtoH_v2(dst, src, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoD_v2(CUdeviceptr dst, CUdeviceptr src, size_t n) {
    TOUCH2(dst, n, src, n, nullptr); CUresult r = drv().cuMemcpyDtoD_v2(dst, src, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr dst, const void *src, size_t n, CUstream st) {
    TOUCH1(dst, n, st); CUresult r = drv().cuMemcpyHtoDAsync_v2(dst, src, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoHAsync_v2(void *dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH1R(src, n, st); CUresult r = drv().cuMemcpyDtoHAsync_v2(dst, src, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoDAsync_v2(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH2(dst, n, src, n, st); CUresult r = drv().cuMemcpyDtoDAsync_v2(dst, src, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemcpy(CUdeviceptr dst, CUdeviceptr src, size_t n) {
    TOUCH2(dst, n, src, n, nullptr); CUresult r = drv().cuMemcpy(dst, src, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpyAsync(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH2(dst, n, src, n, st); CUresult r = drv().cuMemcpyAsync(dst, src, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD8_v2(CUdeviceptr dst, unsigned char v, size_t n) {
    TOUCH1(dst, n, nullptr); CUresult r = drv().cuMemsetD8_v2(dst, v, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD16_v2(CUdeviceptr dst, unsigned short v, size_t n) {
    TOUCH1(dst, n * 2, nullptr); CUresult r = drv().cuMemsetD16_v2(dst, v, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD32_v2(CUdeviceptr dst, unsigned int v, size_t n) {
    TOUCH1(dst, n * 4, nullptr); CUresult r = drv().cuMemsetD32_v2(dst, v, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD8Async(CUdeviceptr dst, unsigned char v, size_t n, CUstream st) {
    TOUCH1(dst, n, st); CUresult r = drv().cuMemsetD8Async(dst, v, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD16Async(CUdeviceptr dst, unsigned short v, size_t n, CUstream st) {
    TOUCH1(dst, n * 2, st); CUresult r = drv().cuMemsetD16Async(dst, v, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD32Async(CUdeviceptr dst, unsigned int v, size_t n, CUstream st) {
    TOUCH1(dst, n * 4, st); CUresult r = drv().cuMemsetD32Async(dst, v, n, st); TOUCH_DONE(st); return r;
}

// per-thread-default-stream twins of the copy / fill family (a null stream is the calling thread's own stream)
#define PTS(st) ((st) ? (st) : CU_STREAM_PER_THREAD)
VGPU_EXPORT CUresult cuMemcpyHtoD_v2_ptds(CUdeviceptr dst, const void *src, size_t n) {
    TOUCH1(dst, n, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemcpyHtoD_v2_ptds(dst, src, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoH_v2_ptds(void *dst, CUdeviceptr src, size_t n) {
    TOUCH1R(src, n, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemcpyDtoH_v2_ptds(dst, src, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoD_v2_ptds(CUdeviceptr dst, CUdeviceptr src, size_t n) {
    TOUCH2(dst, n, src, n, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemcpyDtoD_v2_ptds(dst, src, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemcpy_ptds(CUdeviceptr dst, CUdeviceptr src, size_t n) {
    TOUCH2(dst, n, src, n, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemcpy_ptds(dst, src, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemcpyHtoDAsync_v2_ptsz(CUdeviceptr dst, const void *src, size_t n, CUstream st) {
    TOUCH1(dst, n, PTS(st)); CUresult r = drv().cuMemcpyHtoDAsync_v2_ptsz(dst, src, n, st); TOUCH_DONE(PTS(st)); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoHAsync_v2_ptsz(void *dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH1R(src, n, PTS(st)); CUresult r = drv().cuMemcpyDtoHAsync_v2_ptsz(dst, src, n, st); TOUCH_DONE(PTS(st)); return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoDAsync_v2_ptsz(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH2(dst, n, src, n, PTS(st)); CUresult r = drv().cuMemcpyDtoDAsync_v2_ptsz(dst, src, n, st); TOUCH_DONE(PTS(st)); return r;
}
VGPU_EXPORT CUresult cuMemcpyAsync_ptsz(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream st) {
    TOUCH2(dst, n, src, n, PTS(st)); CUresult r = drv().cuMemcpyAsync_ptsz(dst, src, n, st); TOUCH_DONE(PTS(st)); return r;
}
// Batched copies (CUDA 12.8). The reference predates them: its managed buffers simply fault in under the copy engines. Here
// every swappable operand of the batch is admitted at once; a batch whose operands do not fit the resident quota together
// is issued as single copies, each with its own admission (the batch's attributes are placement / ordering hints).
template <typename Batch, typename Single>
static CUresult batch_copy(CUdeviceptr *dsts, CUdeviceptr *srcs, size_t *sizes, size_t count, size_t *failIdx, CUstream eff, Batch batch, Single single) {
    if (!dsts || !srcs || !sizes || count == 0) return batch();
    if (Runtime::get().touch_batch(dsts, count, srcs, count, eff)) { CUresult r = batch(); TOUCH_DONE(eff); return r; }
    for (size_t i = 0; i < count; i++) {
        TOUCH2(dsts[i], sizes[i], srcs[i], sizes[i], eff);
        CUresult r = single(dsts[i], srcs[i], sizes[i]);
        TOUCH_DONE(eff);
        if (r != CUDA_SUCCESS) { if (failIdx) *failIdx = i; return r; }
    }
    return CUDA_SUCCESS;
}
VGPU_EXPORT CUresult cuMemcpyBatchAsync(CUdeviceptr *dsts, CUdeviceptr *srcs, size_t *sizes, size_t count, CUmemcpyAttributes *attrs, size_t *attrsIdxs,
                                        size_t numAttrs, size_t *failIdx, CUstream st) {
    if (!drv().cuMemcpyBatchAsync) return CUDA_ERROR_NOT_SUPPORTED;
    return batch_copy(dsts, srcs, sizes, count, failIdx, st,
                      [&] { return drv().cuMemcpyBatchAsync(dsts, srcs, sizes, count, attrs, attrsIdxs, numAttrs, failIdx, st); },
                      [&](CUdeviceptr d, CUdeviceptr s, size_t n) { return drv().cuMemcpyAsync(d, s, n, st); });
}
VGPU_EXPORT CUresult cuMemcpyBatchAsync_ptsz(CUdeviceptr *dsts, CUdeviceptr *srcs, size_t *sizes, size_t count, CUmemcpyAttributes *attrs, size_t *attrsIdxs,
                                             size_t numAttrs, size_t *failIdx, CUstream st) {
    if (!drv().cuMemcpyBatchAsync_ptsz) return CUDA_ERROR_NOT_SUPPORTED;
    return batch_copy(dsts, srcs, sizes, count, failIdx, PTS(st),
                      [&] { return drv().cuMemcpyBatchAsync_ptsz(dsts, srcs, sizes, count, attrs, attrsIdxs, numAttrs, failIdx, st); },
                      [&](CUdeviceptr d, CUdeviceptr s, size_t n) { return drv().cuMemcpyAsync_ptsz(d, s, n, st); });
}
// 3D batches: pointer operands are admitted together (array operands are not swappable); a batch that cannot be resident
// at once is refused — there is no one-by-one form of a 3D batch op that keeps its meaning.
template <typename Real>
static CUresult batch_copy_3d(size_t numOps, CUDA_MEMCPY3D_BATCH_OP *ops, CUstream eff, Real real) {
    if (!ops || numOps == 0) return real();
    std::vector<CUdeviceptr> w, r;
    for (size_t i = 0; i < numOps; i++) {
        if (ops[i].dst.type == CU_MEMCPY_OPERAND_TYPE_POINTER) w.push_back(ops[i].dst.op.ptr.ptr);
        if (ops[i].src.type == CU_MEMCPY_OPERAND_TYPE_POINTER) r.push_back(ops[i].src.op.ptr.ptr);
    }
    if (!Runtime::get().touch_batch(w.data(), w.size(), r.data(), r.size(), eff)) return CUDA_ERROR_OUT_OF_MEMORY;
    CUresult rc = real();
    TOUCH_DONE(eff);
    return rc;
}
VGPU_EXPORT CUresult cuMemcpy3DBatchAsync(size_t numOps, CUDA_MEMCPY3D_BATCH_OP *opList, size_t *failIdx, unsigned long long flags, CUstream st) {
    if (!drv().cuMemcpy3DBatchAsync) return CUDA_ERROR_NOT_SUPPORTED;
    return batch_copy_3d(numOps, opList, st, [&] { return drv().cuMemcpy3DBatchAsync(numOps, opList, failIdx, flags, st); });
}
VGPU_EXPORT CUresult cuMemcpy3DBatchAsync_ptsz(size_t numOps, CUDA_MEMCPY3D_BATCH_OP *opList, size_t *failIdx, unsigned long long flags, CUstream st) {
    if (!drv().cuMemcpy3DBatchAsync_ptsz) return CUDA_ERROR_NOT_SUPPORTED;
    return batch_copy_3d(numOps, opList, PTS(st), [&] { return drv().cuMemcpy3DBatchAsync_ptsz(numOps, opList, failIdx, flags, st); });
}
VGPU_EXPORT CUresult cuMemsetD8_v2_ptds(CUdeviceptr dst, unsigned char v, size_t n) {
    TOUCH1(dst, n, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemsetD8_v2_ptds(dst, v, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemsetD16_v2_ptds(CUdeviceptr dst, unsigned short v, size_t n) {
    TOUCH1(dst, n * 2, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemsetD16_v2_ptds(dst, v, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemsetD32_v2_ptds(CUdeviceptr dst, unsigned int v, size_t n) {
    TOUCH1(dst, n * 4, CU_STREAM_PER_THREAD); CUresult r = drv().cuMemsetD32_v2_ptds(dst, v, n); TOUCH_DONE(CU_STREAM_PER_THREAD); return r;
}
VGPU_EXPORT CUresult cuMemsetD8Async_ptsz(CUdeviceptr dst, unsigned char v, size_t n, CUstream st) {
    TOUCH1(dst, n, PTS(st)); CUresult r = drv().cuMemsetD8Async_ptsz(dst, v, n, st); TOUCH_DONE(PTS(st)); return r;
}
VGPU_EXPORT CUresult cuMemsetD16Async_ptsz(CUdeviceptr dst, unsigned short v, size_t n, CUstream st) {
    TOUCH1(dst, n * 2, PTS(st)); CUresult r = drv().cuMemsetD16Async_ptsz(dst, v, n, st); TOUCH_DONE(PTS(st)); return r;
}
VGPU_EXPORT CUresult cuMemsetD32Async_ptsz(CUdeviceptr dst, unsigned int v, size_t n, CUstream st) {
    TOUCH1(dst, n * 4, PTS(st)); CUresult r = drv().cuMemsetD32Async_ptsz(dst, v, n, st); TOUCH_DONE(PTS(st)); return r;
}
#undef PTS

// 2-D / 3-D / peer copies and 2-D fills: same rule as the linear family — device operands living in the swap arena are
// paged in before the real call (the reference needs nothing here: UVM faults its managed memory in).
#define DEVPTR_IF(type, ptr) ((type) == CU_MEMORYTYPE_DEVICE || (type) == CU_MEMORYTYPE_UNIFIED ? (ptr) : (CUdeviceptr)0)
VGPU_EXPORT CUresult cuMemcpy2D_v2(const CUDA_MEMCPY2D *c) {
    if (c) TOUCH2(DEVPTR_IF(c->dstMemoryType, c->dstDevice), 1, DEVPTR_IF(c->srcMemoryType, c->srcDevice), 1, nullptr);
    CUresult r = drv().cuMemcpy2D_v2(c); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpy2DUnaligned_v2(const CUDA_MEMCPY2D *c) {
    if (c) TOUCH2(DEVPTR_IF(c->dstMemoryType, c->dstDevice), 1, DEVPTR_IF(c->srcMemoryType, c->srcDevice), 1, nullptr);
    CUresult r = drv().cuMemcpy2DUnaligned_v2(c); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpy2DAsync_v2(const CUDA_MEMCPY2D *c, CUstream st) {
    if (c) TOUCH2(DEVPTR_IF(c->dstMemoryType, c->dstDevice), 1, DEVPTR_IF(c->srcMemoryType, c->srcDevice), 1, st);
    CUresult r = drv().cuMemcpy2DAsync_v2(c, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemcpy3D_v2(const CUDA_MEMCPY3D *c) {
    if (c) TOUCH2(DEVPTR_IF(c->dstMemoryType, c->dstDevice), 1, DEVPTR_IF(c->srcMemoryType, c->srcDevice), 1, nullptr);
    CUresult r = drv().cuMemcpy3D_v2(c); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpy3DAsync_v2(const CUDA_MEMCPY3D *c, CUstream st) {
    if (c) TOUCH2(DEVPTR_IF(c->dstMemoryType, c->dstDevice), 1, DEVPTR_IF(c->srcMemoryType, c->srcDevice), 1, st);
    CUresult r = drv().cuMemcpy3DAsync_v2(c, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemcpyPeer(CUdeviceptr dst, CUcontext dctx, CUdeviceptr src, CUcontext sctx, size_t n) {
    TOUCH2(dst, n, src, n, nullptr); CUresult r = drv().cuMemcpyPeer(dst, dctx, src, sctx, n); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemcpyPeerAsync(CUdeviceptr dst, CUcontext dctx, CUdeviceptr src, CUcontext sctx, size_t n, CUstream st) {
    TOUCH2(dst, n, src, n, st); CUresult r = drv().cuMemcpyPeerAsync(dst, dctx, src, sctx, n, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D8_v2(CUdeviceptr dst, size_t pitch, unsigned char v, size_t w, size_t h) {
    TOUCH1(dst, pitch * h, nullptr); CUresult r = drv().cuMemsetD2D8_v2(dst, pitch, v, w, h); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D16_v2(CUdeviceptr dst, size_t pitch, unsigned short v, size_t w, size_t h) {
    TOUCH1(dst, pitch * h, nullptr); CUresult r = drv().cuMemsetD2D16_v2(dst, pitch, v, w, h); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D32_v2(CUdeviceptr dst, size_t pitch, unsigned int v, size_t w, size_t h) {
    TOUCH1(dst, pitch * h, nullptr); CUresult r = drv().cuMemsetD2D32_v2(dst, pitch, v, w, h); TOUCH_DONE(nullptr); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D8Async(CUdeviceptr dst, size_t pitch, unsigned char v, size_t w, size_t h, CUstream st) {
    TOUCH1(dst, pitch * h, st); CUresult r = drv().cuMemsetD2D8Async(dst, pitch, v, w, h, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D16Async(CUdeviceptr dst, size_t pitch, unsigned short v, size_t w, size_t h, CUstream st) {
    TOUCH1(dst, pitch * h, st); CUresult r = drv().cuMemsetD2D16Async(dst, pitch, v, w, h, st); TOUCH_DONE(st); return r;
}
VGPU_EXPORT CUresult cuMemsetD2D32Async(CUdeviceptr dst, size_t pitch, unsigned int v, size_t w, size_t h, CUstream st) {
    TOUCH1(dst, pitch * h, st); CUresult r = drv().cuMemsetD2D32Async(dst, pitch, v, w, h, st); TOUCH_DONE(st); return r;
}

// Pointer queries. A PAGED-OUT swappable buffer has no mapping at all and the driver would answer INVALID_VALUE, so the
// buffer is made resident first; resident or not, it is a VMM mapping — plain device memory to the driver — and the
// driver's answers are passed on as they are.
// The reference post-processes the plural call only (cuPointerGetAttributes@0x33187; the singular @0x32f8c forwards):
// because its swap switch turns large allocations into managed memory behind the application's back, it writes 0 over
// every CU_POINTER_ATTRIBUTE_IS_MANAGED answer — for every pointer, the application's own cuMemAllocManaged memory
// included (run against the binary: type=3 managed=0) — and leaves MEMORY_TYPE alone (check_memory_type@0x407f2 is only
// called for its debug log, @0x333aa-0x33447). Nothing is hidden behind managed memory here, so the blanket override
// is not needed and would only mislead cudaMemPrefetchAsync-style callers; VGPU_REFERENCE_COVERAGE=1 restores it.
VGPU_EXPORT CUresult cuPointerGetAttribute(void *data, CUpointer_attribute attribute, CUdeviceptr ptr) {
    if (!drv().cuPointerGetAttribute) return CUDA_ERROR_NOT_SUPPORTED;
    TOUCH1R(ptr, 1, nullptr);
    CUresult r = drv().cuPointerGetAttribute(data, attribute, ptr);
    TOUCH_DONE(nullptr);
    return r;
}
VGPU_EXPORT CUresult cuPointerGetAttributes(unsigned int numAttributes, CUpointer_attribute *attributes, void **data, CUdeviceptr ptr) {
    if (!drv().cuPointerGetAttributes) return CUDA_ERROR_NOT_SUPPORTED;
    TOUCH1R(ptr, 1, nullptr);
    CUresult r = drv().cuPointerGetAttributes(numAttributes, attributes, data, ptr);
    TOUCH_DONE(nullptr);
    if (attributes && data && Runtime::get().reference_coverage_mode())
        for (unsigned int i = 0; i < numAttributes; i++)
            if (attributes[i] == CU_POINTER_ATTRIBUTE_IS_MANAGED && data[i]) *static_cast<unsigned int *>(data[i]) = 0;
    return r;
}

// Explicitly built graphs. Stream capture is handled at the captured call (guarded_launch / touch_range2); a graph built node
// by node never passes there, and its replay (cuGraphLaunch) cannot be admitted operand by operand. So the swappable
// buffers a kernel / memcpy / memset node names are made resident when the node is defined or re-parameterised and stay
// pinned from then on. The reference needs nothing here: UVM faults under a replay as under any launch.
static CUresult pin_kernel_node(const CUDA_KERNEL_NODE_PARAMS *p) {
    return p ? Runtime::get().pin_graph_kernel(p->func, p->kernelParams, p->extra) : CUDA_SUCCESS;
}
static CUresult pin_copy_node(const CUDA_MEMCPY3D *c) {
    if (!c) return CUDA_SUCCESS;
    CUdeviceptr p[2] = {DEVPTR_IF(c->dstMemoryType, c->dstDevice), DEVPTR_IF(c->srcMemoryType, c->srcDevice)};
    return Runtime::get().pin_graph_ptrs(p, 2);
}
VGPU_EXPORT CUresult cuGraphAddKernelNode_v2(CUgraphNode *node, CUgraph g, const CUgraphNode *deps, size_t ndeps, const CUDA_KERNEL_NODE_PARAMS *p) {
    if (!drv().cuGraphAddKernelNode_v2) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_kernel_node(p)) return r;
    return drv().cuGraphAddKernelNode_v2(node, g, deps, ndeps, p);
}
VGPU_EXPORT CUresult cuGraphKernelNodeSetParams_v2(CUgraphNode node, const CUDA_KERNEL_NODE_PARAMS *p) {
    if (!drv().cuGraphKernelNodeSetParams_v2) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_kernel_node(p)) return r;
    return drv().cuGraphKernelNodeSetParams_v2(node, p);
}
VGPU_EXPORT CUresult cuGraphExecKernelNodeSetParams_v2(CUgraphExec ge, CUgraphNode node, const CUDA_KERNEL_NODE_PARAMS *p) {
    if (!drv().cuGraphExecKernelNodeSetParams_v2) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_kernel_node(p)) return r;
    return drv().cuGraphExecKernelNodeSetParams_v2(ge, node, p);
}
VGPU_EXPORT CUresult cuGraphAddMemcpyNode(CUgraphNode *node, CUgraph g, const CUgraphNode *deps, size_t ndeps, const CUDA_MEMCPY3D *c, CUcontext ctx) {
    if (!drv().cuGraphAddMemcpyNode) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_copy_node(c)) return r;
    return drv().cuGraphAddMemcpyNode(node, g, deps, ndeps, c, ctx);
}
VGPU_EXPORT CUresult cuGraphMemcpyNodeSetParams(CUgraphNode node, const CUDA_MEMCPY3D *c) {
    if (!drv().cuGraphMemcpyNodeSetParams) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_copy_node(c)) return r;
    return drv().cuGraphMemcpyNodeSetParams(node, c);
}
VGPU_EXPORT CUresult cuGraphExecMemcpyNodeSetParams(CUgraphExec ge, CUgraphNode node, const CUDA_MEMCPY3D *c, CUcontext ctx) {
    if (!drv().cuGraphExecMemcpyNodeSetParams) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_copy_node(c)) return r;
    return drv().cuGraphExecMemcpyNodeSetParams(ge, node, c, ctx);
}
VGPU_EXPORT CUresult cuGraphAddMemsetNode(CUgraphNode *node, CUgraph g, const CUgraphNode *deps, size_t ndeps, const CUDA_MEMSET_NODE_PARAMS *m, CUcontext ctx) {
    if (!drv().cuGraphAddMemsetNode) return CUDA_ERROR_NOT_SUPPORTED;
    if (m) if (CUresult r = Runtime::get().pin_graph_ptrs(&m->dst, 1)) return r;
    return drv().cuGraphAddMemsetNode(node, g, deps, ndeps, m, ctx);
}
static CUresult pin_generic_node(const CUgraphNodeParams *np) {
    if (!np) return CUDA_SUCCESS;
    switch (np->type) {
    case CU_GRAPH_NODE_TYPE_KERNEL: return Runtime::get().pin_graph_kernel(np->kernel.func, np->kernel.kernelParams, np->kernel.extra);
    case CU_GRAPH_NODE_TYPE_MEMCPY: return pin_copy_node(&np->memcpy.copyParams);
    case CU_GRAPH_NODE_TYPE_MEMSET: return Runtime::get().pin_graph_ptrs(&np->memset.dst, 1);
    default: return CUDA_SUCCESS;
    }
}
VGPU_EXPORT CUresult cuGraphAddNode(CUgraphNode *node, CUgraph g, const CUgraphNode *deps, size_t ndeps, CUgraphNodeParams *np) {
    if (!drv().cuGraphAddNode) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_generic_node(np)) return r;
    return drv().cuGraphAddNode(node, g, deps, ndeps, np);
}
VGPU_EXPORT CUresult cuGraphAddNode_v2(CUgraphNode *node, CUgraph g, const CUgraphNode *deps, const CUgraphEdgeData *edges, size_t ndeps, CUgraphNodeParams *np) {
    if (!drv().cuGraphAddNode_v2) return CUDA_ERROR_NOT_SUPPORTED;
    if (CUresult r = pin_generic_node(np)) return r;
    return drv().cuGraphAddNode_v2(node, g, deps, edges, ndeps, np);
}

// cuMemGetAddressRange: a swappable buffer is answered from the engine's table — base and size as the application allocated
// it, resident or not (the driver knows a resident one as a VMM mapping rounded to the 2 MiB granule, and a paged-out one
// not at all). The reference forwards: its managed allocation is an ordinary allocation to the driver.
VGPU_EXPORT CUresult cuMemGetAddressRange_v2(CUdeviceptr *pbase, size_t *psize, CUdeviceptr dptr) {
    if (Runtime::get().swap_address_range(dptr, pbase, psize)) return CUDA_SUCCESS;
    return drv().cuMemGetAddressRange_v2 ? drv().cuMemGetAddressRange_v2(pbase, psize, dptr) : CUDA_ERROR_NOT_SUPPORTED;
}

// cuMemAdvise / cuMemPrefetchAsync. In the reference a large allocation under CUDA_OVERSUBSCRIBE IS managed memory
// (cuMemoryAllocate@0x315da -> cuMemAllocManaged), so both calls work on it and steer UVM. Here the same pointers are the
// swap engine's: SET_READ_MOSTLY keeps kernel launches from dirtying the range (eviction without write-back), a prefetch
// to the device queues a page-in with the pager; everything else about them is accepted and ignored. Other pointers go to
// the driver unchanged.
VGPU_EXPORT CUresult cuMemAdvise(CUdeviceptr devPtr, size_t count, CUmem_advise advice, CUdevice device) {
    if (Runtime::get().swap_advise(devPtr, advice)) return CUDA_SUCCESS;
    return drv().cuMemAdvise ? drv().cuMemAdvise(devPtr, count, advice, device) : CUDA_ERROR_NOT_SUPPORTED;
}
VGPU_EXPORT CUresult cuMemAdvise_v2(CUdeviceptr devPtr, size_t count, CUmem_advise advice, CUmemLocation location) {
    if (Runtime::get().swap_advise(devPtr, advice)) return CUDA_SUCCESS;
    return drv().cuMemAdvise_v2 ? drv().cuMemAdvise_v2(devPtr, count, advice, location) : CUDA_ERROR_NOT_SUPPORTED;
}
VGPU_EXPORT CUresult cuMemPrefetchAsync(CUdeviceptr devPtr, size_t count, CUdevice dstDevice, CUstream hStream) {
    if (Runtime::get().swap_prefetch(devPtr, dstDevice >= 0)) return CUDA_SUCCESS;
    return drv().cuMemPrefetchAsync ? drv().cuMemPrefetchAsync(devPtr, count, dstDevice, hStream) : CUDA_ERROR_NOT_SUPPORTED;
}
VGPU_EXPORT CUresult cuMemPrefetchAsync_ptsz(CUdeviceptr devPtr, size_t count, CUdevice dstDevice, CUstream hStream) {
    if (Runtime::get().swap_prefetch(devPtr, dstDevice >= 0)) return CUDA_SUCCESS;
    return drv().cuMemPrefetchAsync_ptsz ? drv().cuMemPrefetchAsync_ptsz(devPtr, count, dstDevice, hStream) : CUDA_ERROR_NOT_SUPPORTED;
}
VGPU_EXPORT CUresult cuMemPrefetchAsync_v2(CUdeviceptr devPtr, size_t count, CUmemLocation location, unsigned int flags, CUstream hStream) {
    if (Runtime::get().swap_prefetch(devPtr, location.type == CU_MEM_LOCATION_TYPE_DEVICE)) return CUDA_SUCCESS;
    return drv().cuMemPrefetchAsync_v2 ? drv().cuMemPrefetchAsync_v2(devPtr, count, location, flags, hStream) : CUDA_ERROR_NOT_SUPPORTED;
}
VGPU_EXPORT CUresult cuMemPrefetchAsync_v2_ptsz(CUdeviceptr devPtr, size_t count, CUmemLocation location, unsigned int flags, CUstream hStream) {
    if (Runtime::get().swap_prefetch(devPtr, location.type == CU_MEM_LOCATION_TYPE_DEVICE)) return CUDA_SUCCESS;
    return drv().cuMemPrefetchAsync_v2_ptsz ? drv().cuMemPrefetchAsync_v2_ptsz(devPtr, count, location, flags, hStream) : CUDA_ERROR_NOT_SUPPORTED;
}

// extras the reference exports for its own tooling
VGPU_EXPORT CUresult cuMemoryAllocate(CUdeviceptr *dptr, size_t bytesize, size_t *bytesallocated, void *data) {
    (void)data;                                   // cuMemoryAllocate@0x315da: the allocmode switch
    if (bytesallocated) *bytesallocated = bytesize;
    return Runtime::get().mem_alloc(dptr, bytesize);
}
VGPU_EXPORT CUresult cuMemoryFree(CUdeviceptr dptr) { return Runtime::get().mem_free(dptr); }   // cuMemoryFree@0x3792f
VGPU_EXPORT int cuVGPUViewAllocator(void) {       // view_vgpu_allocator@0x3fc34
    Runtime &rt = Runtime::get();
    std::fprintf(stderr, "[vgpu-b200] allocation table: %zu entries, context_size=%lu\n", rt.table_size(),
                 (unsigned long)rt.context_size());
    return 0;
}

// ------------------------------------------------------------------------------------------------ hooked NVML
VGPU_EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t device, nvmlMemory_t *memory) {
    // nvmlDeviceGetMemoryInfo@0x24069 (nvml/hook.c:L327-334): under a quota, nvidia-smi inside the container sees
    // total = limit, used = container usage
    const vgpu::NvmlTable &n = vgpu::nvml();
    if (!n.nvmlDeviceGetMemoryInfo) return NVML_ERROR_FUNCTION_NOT_FOUND;
    nvmlReturn_t r = n.nvmlDeviceGetMemoryInfo(device, memory);
    if (r != NVML_SUCCESS || !memory) return r;
    unsigned idx = 0;
    if (n.nvmlDeviceGetIndex && n.nvmlDeviceGetIndex(device, &idx) == NVML_SUCCESS) {
        unsigned long long t = 0, f = 0, u = 0;
        if (Runtime::get().nvml_memory_view((int)idx, &t, &f, &u)) { memory->used = u; if (t) { memory->total = t; memory->free = f; } }
    }
    return r;
}
VGPU_EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo_v2(nvmlDevice_t device, nvmlMemory_v2_t *memory) {
    const vgpu::NvmlTable &n = vgpu::nvml();
    if (!n.nvmlDeviceGetMemoryInfo_v2) return NVML_ERROR_FUNCTION_NOT_FOUND;
    nvmlReturn_t r = n.nvmlDeviceGetMemoryInfo_v2(device, memory);
    if (r != NVML_SUCCESS || !memory) return r;
    unsigned idx = 0;
    if (n.nvmlDeviceGetIndex && n.nvmlDeviceGetIndex(device, &idx) == NVML_SUCCESS) {
        unsigned long long t = 0, f = 0, u = 0;
        if (Runtime::get().nvml_memory_view((int)idx, &t, &f, &u)) { memory->used = u; if (t) { memory->total = t; memory->free = f; memory->reserved = 0; } }
    }
    return r;
}

// ------------------------------------------------------------------------------------------------ symbol routing
VGPU_EXPORT CUresult cuGetProcAddress_v2(const char *symbol, void **pfn, int cudaVersion, cuuint64_t flags,
                                         CUdriverProcAddressQueryResult *symbolStatus);
VGPU_EXPORT CUresult cuGetProcAddress(const char *symbol, void **pfn, int cudaVersion, cuuint64_t flags);

namespace {
// name -> wrapper, plus the address of the REAL entry point the wrapper forwards to. cuGetProcAddress requests are
// resolved by the real driver first and the wrapper is substituted only when the driver's answer IS that real entry
// point: the ABI version the caller gets (cudaVersion / flags dependent: _v2 vs _v3, _ptsz ...) is then exactly
// the one the wrapper implements. The reference substitutes by NAME (find_symbols_in_table@0x2e172 tries name_v3,
// name_v2, name and ignores cudaVersion/flags), which hands e.g. a cuCtxCreate_v2-shaped wrapper to a caller that
// asked for the 4-argument cuCtxCreate_v4 of CUDA 12.5+.
struct HookEntry { const char *name; void *fn; void *const *real; };
#define H(sym) {#sym, reinterpret_cast<void *>(&sym), reinterpret_cast<void *const *>(&drv().sym)}
#define HN(sym) {#sym, reinterpret_cast<void *>(&sym), nullptr}
const std::vector<HookEntry> &hooks() {
    static const std::vector<HookEntry> *t = new std::vector<HookEntry>{
        H(cuInit),
        {"cuGetProcAddress", reinterpret_cast<void *>(&cuGetProcAddress), reinterpret_cast<void *const *>(&drv().cuGetProcAddress_v1)},
        H(cuGetProcAddress_v2),
        H(cuMemAlloc_v2), H(cuMemAllocManaged), H(cuMemAllocPitch_v2), H(cuMemFree_v2), H(cuMemGetInfo_v2),
        H(cuDeviceTotalMem_v2), H(cuDevicePrimaryCtxRetain), H(cuCtxCreate_v2), H(cuMemHostAlloc), H(cuMemAllocHost_v2),
        H(cuMemHostRegister_v2), H(cuMipmappedArrayCreate), H(cuPointerGetAttribute), H(cuPointerGetAttributes),
        H(cuMemcpy2D_v2), H(cuMemcpy2DUnaligned_v2), H(cuMemcpy2DAsync_v2), H(cuMemcpy3D_v2), H(cuMemcpy3DAsync_v2), H(cuMemcpyPeer), H(cuMemcpyPeerAsync),
        H(cuMemsetD2D8_v2), H(cuMemsetD2D16_v2), H(cuMemsetD2D32_v2), H(cuMemsetD2D8Async), H(cuMemsetD2D16Async), H(cuMemsetD2D32Async),
        H(cuLaunchKernel), H(cuLaunchKernelEx), H(cuLaunchCooperativeKernel), H(cuModuleUnload),
        H(cuMemcpyHtoD_v2), H(cuMemcpyDtoH_v2), H(cuMemcpyDtoD_v2), H(cuMemcpyHtoDAsync_v2), H(cuMemcpyDtoHAsync_v2),
        H(cuMemcpyDtoDAsync_v2), H(cuMemcpy), H(cuMemcpyAsync),
        H(cuMemsetD8_v2), H(cuMemsetD16_v2), H(cuMemsetD32_v2), H(cuMemsetD8Async), H(cuMemsetD16Async), H(cuMemsetD32Async),
        H(cuLaunchKernel_ptsz), H(cuLaunchKernelEx_ptsz), H(cuLaunchCooperativeKernel_ptsz), H(cuGraphLaunch), H(cuGraphLaunch_ptsz),
        H(cuMemAllocAsync), H(cuMemAllocAsync_ptsz), H(cuMemAllocFromPoolAsync), H(cuMemAllocFromPoolAsync_ptsz), H(cuMemFreeAsync),
        H(cuMemFreeAsync_ptsz), H(cuMemCreate), H(cuMemRelease), H(cuMemMap), H(cuMemUnmap),
        H(cuMemcpyHtoD_v2_ptds), H(cuMemcpyDtoH_v2_ptds), H(cuMemcpyDtoD_v2_ptds), H(cuMemcpy_ptds), H(cuMemcpyHtoDAsync_v2_ptsz),
        H(cuMemcpyDtoHAsync_v2_ptsz), H(cuMemcpyDtoDAsync_v2_ptsz), H(cuMemcpyAsync_ptsz), H(cuMemsetD8_v2_ptds), H(cuMemsetD16_v2_ptds),
        H(cuMemsetD32_v2_ptds), H(cuMemsetD8Async_ptsz), H(cuMemsetD16Async_ptsz), H(cuMemsetD32Async_ptsz),
        H(cuMemAdvise), H(cuMemAdvise_v2), H(cuMemPrefetchAsync),
        H(cuMemGetAddressRange_v2),
        H(cuGraphAddKernelNode_v2), H(cuGraphKernelNodeSetParams_v2), H(cuGraphExecKernelNodeSetParams_v2), H(cuGraphAddMemcpyNode),
        H(cuGraphMemcpyNodeSetParams), H(cuGraphExecMemcpyNodeSetParams), H(cuGraphAddMemsetNode), H(cuGraphAddNode), H(cuGraphAddNode_v2),
        H(cuMemPrefetchAsync_ptsz), H(cuMemPrefetchAsync_v2), H(cuMemPrefetchAsync_v2_ptsz),
        H(cuMemcpyBatchAsync), H(cuMemcpyBatchAsync_ptsz), H(cuMemcpy3DBatchAsync), H(cuMemcpy3DBatchAsync_ptsz),
        HN(cuMemoryAllocate), HN(cuMemoryFree), HN(cuVGPUViewAllocator),
        HN(nvmlDeviceGetMemoryInfo), HN(nvmlDeviceGetMemoryInfo_v2),
    };
    return *t;
}
#undef H
#undef HN

void *find_hook_exact(const char *name) {
    for (const HookEntry &e : hooks())
        if (!std::strcmp(e.name, name)) return e.fn;
    return nullptr;
}
void *find_hook_for_real(void *real_fn) {
    if (!real_fn) return nullptr;
    for (const HookEntry &e : hooks())
        if (e.real && *e.real == real_fn) return e.fn;
    return nullptr;
}
// VGPU_TRACE_GPA=1: log every cuGetProcAddress request and whether a wrapper was substituted (diagnostics)
bool trace_gpa() {
    static bool on = std::getenv("VGPU_TRACE_GPA") != nullptr;
    return on;
}
bool control_disabled() {
    static bool off = std::getenv("CUDA_DISABLE_CONTROL") != nullptr;   // container opt-out (server.go:380-385)
    return off;
}
}  // namespace

VGPU_EXPORT CUresult cuGetProcAddress_v2(const char *symbol, void **pfn, int cudaVersion, cuuint64_t flags,
                                         CUdriverProcAddressQueryResult *symbolStatus) {
    CUresult r;
    if (drv().cuGetProcAddress_v2) {
        r = drv().cuGetProcAddress_v2(symbol, pfn, cudaVersion, flags, symbolStatus);
    } else if (drv().cuGetProcAddress_v1) {
        r = drv().cuGetProcAddress_v1(symbol, pfn, cudaVersion, flags);
        if (symbolStatus) *symbolStatus = r == CUDA_SUCCESS ? CU_GET_PROC_ADDRESS_SUCCESS : CU_GET_PROC_ADDRESS_SYMBOL_NOT_FOUND;
    } else {
        return CUDA_ERROR_NOT_INITIALIZED;
    }
    if (r == CUDA_SUCCESS && pfn && !control_disabled()) {
        void *h = find_hook_for_real(*pfn);
        if (trace_gpa()) std::fprintf(stderr, "[vgpu-b200 gpa] %s ver=%d flags=%llu -> %p %s\n", symbol, cudaVersion, (unsigned long long)flags, *pfn, h ? "HOOKED" : "");
        if (h) *pfn = h;
    }
    return r;
}

VGPU_EXPORT CUresult cuGetProcAddress(const char *symbol, void **pfn, int cudaVersion, cuuint64_t flags) {
    CUresult r;
    if (drv().cuGetProcAddress_v1) r = drv().cuGetProcAddress_v1(symbol, pfn, cudaVersion, flags);
    else if (drv().cuGetProcAddress_v2) r = drv().cuGetProcAddress_v2(symbol, pfn, cudaVersion, flags, nullptr);
    else return CUDA_ERROR_NOT_INITIALIZED;
    if (r == CUDA_SUCCESS && pfn && !control_disabled())
        if (void *h = find_hook_for_real(*pfn)) *pfn = h;
    return r;
}

// dlsym override (libvgpu.so@0x11b36): libcudart and frameworks resolve the driver with dlopen("libcuda.so.1") +
// dlsym(handle, "cu..."), which never consults LD_PRELOAD order — so the lookup itself is intercepted.
#ifndef VGPU_NO_DLSYM_OVERRIDE   // (sanitizer builds leave it out: the sanitizer runtime calls dlsym before its shadow memory exists)
VGPU_EXPORT void *dlsym(void *handle, const char *symbol) {
    // glibc declares the name parameter nonnull, which lets the compiler drop a plain NULL test; callers do pass NULL
    // (and expect the real dlsym's error), so the test goes through a volatile copy
    const char *volatile seen = symbol;
    const char *name = seen;
    if (name && !control_disabled() && ((name[0] == 'c' && name[1] == 'u') || !std::strncmp(name, "nvml", 4))) {
        if (void *h = find_hook_exact(name)) return h;
    }
    return vgpu::real_dlsym(handle, name);
}
#endif
