/*
 * vmm_contend.c — does the cost of a VMM remap (cuMemUnmap + cuMemMap + cuMemSetAccess of one 64 MiB handle, what the
 * swap engine pays per paged buffer) grow when OTHER processes remap on OTHER GPUs of the same box? Run k copies, one
 * per GPU, and compare the per-call latency: a driver-global lock shows up as latency scaling with k.
 *   vmm_contend <device> [seconds] [MiB]
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "vmm_contend: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); exit(3); } } while (0)
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }
int main(int argc, char **argv) {
    int devi = argc > 1 ? atoi(argv[1]) : 0; double seconds = argc > 2 ? atof(argv[2]) : 3.0; size_t mib = argc > 3 ? (size_t)atol(argv[3]) : 64;
    CUdevice dev; CUcontext ctx;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, devi)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = devi;
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t bytes = mib << 20; CUdeviceptr va; CUmemGenericAllocationHandle h;
    CK(cuMemAddressReserve(&va, bytes, 0, 0, 0)); CK(cuMemCreate(&h, bytes, &prop, 0));
    CK(cuMemMap(va, bytes, 0, h, 0)); CK(cuMemSetAccess(va, bytes, &acc, 1));
    double t_un = 0, t_map = 0, t_acc = 0, worst = 0; long n = 0;
    double t0 = now_us();
    while (now_us() - t0 < seconds * 1e6) {
        double a = now_us(); CK(cuMemUnmap(va, bytes));
        double b = now_us(); CK(cuMemMap(va, bytes, 0, h, 0));
        double c = now_us(); CK(cuMemSetAccess(va, bytes, &acc, 1));
        double d = now_us();
        t_un += b - a; t_map += c - b; t_acc += d - c; if (d - a > worst) worst = d - a; n++;
    }
    printf("{\"device\": %d, \"mib\": %zu, \"remaps\": %ld, \"unmap_us\": %.1f, \"map_us\": %.1f, \"setaccess_us\": %.1f, \"remap_us\": %.1f, \"worst_remap_us\": %.1f}\n",
           devi, mib, n, t_un / n, t_map / n, t_acc / n, (t_un + t_map + t_acc) / n, worst);
    return 0;
}
