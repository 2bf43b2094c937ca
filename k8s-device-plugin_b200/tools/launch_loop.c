/*
 * launch_loop.c — BASELINE.json configs[3] as a DRIVER-API program (directly linked cuLaunchKernel, so both the new
 * hook and the reference hook binary interpose it): launches the read-modify-write kernel over a fixed buffer as
 * fast as the intercept lets it for T seconds. Reports the launch count, the un-throttled kernel duration (CUDA
 * events over a back-to-back burst before the loop) and the achieved duty cycle = launches x kernel_ms / wall.
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "launch_loop: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); printf("{\"error\": \"%s rc=%d\"}\n", #x, (int)_r); exit(3); } } while (0)
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s kernels.cubin [mib] [seconds]\n", argv[0]); return 2; }
    size_t mib = argc > 2 ? (size_t)atol(argv[2]) : 2048; double seconds = argc > 3 ? atof(argv[3]) : 6.0;
    CUdevice dev; CUcontext ctx; CUmodule mod; CUfunction f; CUdeviceptr buf; CUevent a, b;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    CK(cuModuleLoad(&mod, argv[1])); CK(cuModuleGetFunction(&f, mod, "vgpu_wl_touch"));
    CK(cuMemAlloc(&buf, mib << 20)); CK(cuMemsetD8(buf, 0, mib << 20));
    CK(cuEventCreate(&a, 0)); CK(cuEventCreate(&b, 0));
    int sm = 148; cuDeviceGetAttribute(&sm, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    unsigned long long nwords = (mib << 20) / 8; void *args[] = {&buf, &nwords};
    /* kernel duration inside a window the limiter cannot cut: its burst allowance is 5 ms */
    CK(cuLaunchKernel(f, sm * 16, 1, 1, 256, 1, 1, 0, 0, args, 0)); CK(cuCtxSynchronize());
    struct timespec nap = {0, 300000000}; nanosleep(&nap, 0);
    CK(cuEventRecord(a, 0)); CK(cuLaunchKernel(f, sm * 16, 1, 1, 256, 1, 1, 0, 0, args, 0)); CK(cuEventRecord(b, 0)); CK(cuCtxSynchronize());
    float kms; CK(cuEventElapsedTime(&kms, a, b));
    nanosleep(&nap, 0);
    double t0 = now_s(); long n = 0;
    while (now_s() - t0 < seconds) { CK(cuLaunchKernel(f, sm * 16, 1, 1, 256, 1, 1, 0, 0, args, 0)); n++; if ((n & 15) == 0) CK(cuCtxSynchronize()); }
    CK(cuCtxSynchronize());
    double wall = now_s() - t0;
    printf("{\"mib\": %zu, \"launches\": %ld, \"wall_s\": %.3f, \"kernel_ms\": %.3f, \"duty\": %.4f}\n", mib, n, wall, kms, n * kms / 1e3 / wall);
    return 0;
}
