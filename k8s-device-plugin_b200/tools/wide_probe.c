/*
 * wide_probe.c — the entry points the reference hook forwards untouched exercised on a real GPU:
 * stream-ordered allocation, VMM physical handles, and a captured graph replayed in a loop. Run bare and under
 * LD_PRELOAD=libvgpu.so with a gpumem quota / gpucores limit; prints one JSON object.
 *   wide_probe kernels.cubin [seconds]
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "wide_probe: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); printf("{\"error\": \"%s rc=%d\"}\n", #x, (int)_r); exit(3); } } while (0)
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
static size_t free_now(void) { size_t f = 0, t = 0; cuMemGetInfo(&f, &t); return f; }
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s kernels.cubin [seconds]\n", argv[0]); return 2; }
    double seconds = argc > 2 ? atof(argv[2]) : 4.0;
    const size_t MiB = 1 << 20;
    CUdevice dev; CUcontext ctx; CUmodule mod; CUfunction f; CUstream st;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    CK(cuModuleLoad(&mod, argv[1])); CK(cuModuleGetFunction(&f, mod, "vgpu_wl_touch"));
    CK(cuStreamCreate(&st, CU_STREAM_NON_BLOCKING));

    /* ---- stream-ordered allocations */
    size_t f0 = free_now();
    CUdeviceptr a = 0, big = 0;
    CUresult r_async = cuMemAllocAsync(&a, 512 * MiB, st);
    CK(cuStreamSynchronize(st));
    size_t f1 = free_now();
    CUresult r_async_big = cuMemAllocAsync(&big, 4096 * MiB, st);
    if (r_async_big == CUDA_SUCCESS) { cuMemFreeAsync(big, st); }
    CUresult r_free_async = a ? cuMemFreeAsync(a, st) : CUDA_ERROR_INVALID_VALUE;
    CK(cuStreamSynchronize(st));
    size_t f2 = free_now();

    /* ---- VMM physical handle, mapped, written through the hooked memset, unmapped, released */
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    size_t vsz = ((512 * MiB + gran - 1) / gran) * gran;
    CUmemGenericAllocationHandle h = 0, hbig = 0;
    CUresult r_create = cuMemCreate(&h, vsz, &prop, 0);
    size_t f3 = free_now();
    CUresult r_create_big = cuMemCreate(&hbig, ((4096 * MiB + gran - 1) / gran) * gran, &prop, 0);
    if (r_create_big == CUDA_SUCCESS) cuMemRelease(hbig);
    int vmm_ok = 0;
    if (r_create == CUDA_SUCCESS) {
        CUdeviceptr va = 0; CK(cuMemAddressReserve(&va, vsz, 0, 0, 0)); CK(cuMemMap(va, vsz, 0, h, 0));
        CUmemAccessDesc ad; memset(&ad, 0, sizeof ad); ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CK(cuMemSetAccess(va, vsz, &ad, 1)); CK(cuMemsetD8(va, 0x5a, vsz));
        unsigned char probe = 0; CK(cuMemcpyDtoH(&probe, va + vsz - 1, 1)); vmm_ok = probe == 0x5a;
        CK(cuMemUnmap(va, vsz)); CK(cuMemRelease(h)); CK(cuMemAddressFree(va, vsz));
    }
    size_t f4 = free_now();

    /* ---- a captured graph of 8 read-modify-write kernels replayed as fast as the intercept allows */
    CUdeviceptr buf; size_t bytes = 1024 * MiB; CK(cuMemAlloc(&buf, bytes)); CK(cuMemsetD8(buf, 0, bytes));
    int sm = 148; cuDeviceGetAttribute(&sm, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    unsigned long long nwords = bytes / 8; void *args[] = {&buf, &nwords};
    CUgraph g; CUgraphExec ge;
    CK(cuStreamBeginCapture(st, CU_STREAM_CAPTURE_MODE_THREAD_LOCAL));
    for (int i = 0; i < 8; i++) CK(cuLaunchKernel(f, sm * 16, 1, 1, 256, 1, 1, 0, st, args, 0));
    CK(cuStreamEndCapture(st, &g));
    CK(cuGraphInstantiate(&ge, g, 0));
    CK(cuGraphLaunch(ge, st)); CK(cuStreamSynchronize(st));
    struct timespec nap = {0, 300000000}; nanosleep(&nap, 0);
    double t0 = now_s(); long n = 0;
    while (now_s() - t0 < seconds) { CK(cuGraphLaunch(ge, st)); n++; if ((n & 3) == 0) CK(cuStreamSynchronize(st)); }
    CK(cuStreamSynchronize(st));
    double wall = now_s() - t0;
    /* every word was incremented once per kernel: 8 per graph launch, +8 for the warm-up launch */
    unsigned long long w0 = 0; CK(cuMemcpyDtoH(&w0, buf, 8));

    printf("{\"async\": {\"rc\": %d, \"rc_big\": %d, \"rc_free\": %d, \"charged\": %ld, \"returned\": %ld}, "
           "\"vmm\": {\"rc\": %d, \"rc_big\": %d, \"charged\": %ld, \"returned\": %ld, \"data_ok\": %d}, "
           "\"graph\": {\"launches\": %ld, \"wall_s\": %.3f, \"word0\": %llu, \"expect_word0\": %ld}}\n",
           (int)r_async, (int)r_async_big, (int)r_free_async, (long)(f0 - f1), (long)(f2 - f1),
           (int)r_create, (int)r_create_big, (long)(f2 - f3), (long)(f4 - f3), vmm_ok, n, wall, w0, (n + 1) * 8);
    return 0;
}
