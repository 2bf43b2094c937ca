/*
 * intercept_bench.c — "intercept overhead % vs bare CUDA" (BASELINE.json metric, SURVEY.md §8d cfg 2): ns per
 * cuMemAlloc_v2+cuMemFree_v2 pair of 1 MiB and ns per cuLaunchKernel of an empty 1x1x1 kernel. Run it bare, under
 * the new hook and under the reference hook; the caller computes (hooked - bare) / bare.
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "intercept_bench: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); printf("{\"error\": \"%s rc=%d\"}\n", #x, (int)_r); exit(3); } } while (0)
static double now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e9 + ts.tv_nsec; }
int main(int argc, char **argv) {
    const char *cubin = argc > 1 ? argv[1] : NULL;
    long n_alloc = argc > 2 ? atol(argv[2]) : 20000, n_launch = argc > 3 ? atol(argv[3]) : 200000;
    if (!cubin) { fprintf(stderr, "usage: %s kernels.cubin [n_alloc] [n_launch]\n", argv[0]); return 2; }
    CUdevice dev; CUcontext ctx; CUmodule mod; CUfunction f;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    CK(cuModuleLoad(&mod, cubin)); CK(cuModuleGetFunction(&f, mod, "vgpu_wl_empty"));
    CUdeviceptr p;
    for (int i = 0; i < 100; i++) { CK(cuMemAlloc(&p, 1 << 20)); CK(cuMemFree(p)); }
    double best_alloc = 1e30, best_launch = 1e30;
    for (int rep = 0; rep < 5; rep++) {
        double t0 = now_ns();
        for (long i = 0; i < n_alloc; i++) { CK(cuMemAlloc(&p, 1 << 20)); CK(cuMemFree(p)); }
        double t1 = now_ns();
        if ((t1 - t0) / n_alloc < best_alloc) best_alloc = (t1 - t0) / n_alloc;
    }
    for (int i = 0; i < 1000; i++) CK(cuLaunchKernel(f, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0));
    CK(cuCtxSynchronize());
    for (int rep = 0; rep < 5; rep++) {
        double t0 = now_ns();
        for (long i = 0; i < n_launch; i++) CK(cuLaunchKernel(f, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0));
        double t1 = now_ns();            /* enqueue cost only: what the intercept adds sits on this thread */
        CK(cuCtxSynchronize());
        if ((t1 - t0) / n_launch < best_launch) best_launch = (t1 - t0) / n_launch;
    }
    printf("{\"alloc_free_1mib_ns\": %.1f, \"launch_empty_ns\": %.1f, \"n_alloc\": %ld, \"n_launch\": %ld}\n", best_alloc, best_launch, n_alloc, n_launch);
    return 0;
}
