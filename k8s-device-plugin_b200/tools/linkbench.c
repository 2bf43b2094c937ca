/*
 * linkbench.c — measures the two denominators the swap path is judged against on THIS box (BASELINE.md §3 "host
 * link: not yet measured"): pinned-memory cuMemcpyHtoDAsync / DtoHAsync bandwidth, one direction at a time and both
 * at once, plus the cost of the VMM calls the swap engine issues per eviction (create/map/setaccess/unmap/release).
 * Driver API only. Output: one JSON object.
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "linkbench: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); exit(3); } } while (0)
static void *h2p_buf(void *p) { return p; }
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }

int main(int argc, char **argv) {
    size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 1024;
    size_t bytes = mib << 20;
    CUdevice dev; CUcontext ctx;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    void *h1, *h2; void *h2_host; CUdeviceptr d1, d2; CUstream s1, s2; CUevent a, b, c, d;
    CK(cuMemHostAlloc(&h1, bytes, CU_MEMHOSTALLOC_PORTABLE)); CK(cuMemHostAlloc(&h2, bytes, CU_MEMHOSTALLOC_PORTABLE));
    memset(h1, 1, bytes); memset(h2, 2, bytes); h2_host = h2;
    CK(cuMemAlloc(&d1, bytes)); CK(cuMemAlloc(&d2, bytes));
    CK(cuStreamCreate(&s1, CU_STREAM_NON_BLOCKING)); CK(cuStreamCreate(&s2, CU_STREAM_NON_BLOCKING));
    CK(cuEventCreate(&a, 0)); CK(cuEventCreate(&b, 0)); CK(cuEventCreate(&c, 0)); CK(cuEventCreate(&d, 0));
    double best_h2d = 0, best_d2h = 0, best_bidir = 0;
    for (int it = 0; it < 10; it++) {
        float ms;
        CK(cuEventRecord(a, s1)); CK(cuMemcpyHtoDAsync(d1, h1, bytes, s1)); CK(cuEventRecord(b, s1)); CK(cuStreamSynchronize(s1));
        CK(cuEventElapsedTime(&ms, a, b)); if (bytes / ms / 1e6 > best_h2d) best_h2d = bytes / ms / 1e6;
        CK(cuEventRecord(a, s1)); CK(cuMemcpyDtoHAsync(h1, d1, bytes, s1)); CK(cuEventRecord(b, s1)); CK(cuStreamSynchronize(s1));
        CK(cuEventElapsedTime(&ms, a, b)); if (bytes / ms / 1e6 > best_d2h) best_d2h = bytes / ms / 1e6;
        /* both directions at once: wall-clock bracket around two streams */
        CK(cuCtxSynchronize());
        double t0 = now_us();
        CK(cuMemcpyHtoDAsync(d1, h1, bytes, s1)); CK(cuMemcpyDtoHAsync(h2, d2, bytes, s2));
        CK(cuStreamSynchronize(s1)); CK(cuStreamSynchronize(s2));
        double t1 = now_us();
        double gbs = 2.0 * bytes / (t1 - t0) / 1e3; if (gbs > best_bidir) best_bidir = gbs;
    }
    /* chunked pipeline as the engine issues it: 32 MiB DMAs back to back */
    size_t chunk = 32u << 20; float ms;
    CK(cuEventRecord(a, s1));
    for (size_t o = 0; o < bytes; o += chunk) CK(cuMemcpyDtoHAsync((char *)h1 + o, d1 + o, chunk, s1));
    CK(cuEventRecord(b, s1)); CK(cuStreamSynchronize(s1)); CK(cuEventElapsedTime(&ms, a, b));
    double d2h_chunked = bytes / ms / 1e6;

    /* VMM op costs */
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t sizes[] = {2u << 20, 64u << 20, 1024u << 20};
    printf("{\"mib\": %zu, \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"bidir_gbs\": %.2f, \"d2h_32m_chunks_gbs\": %.2f, \"vmm_gran\": %zu, \"vmm_us\": {",
           mib, best_h2d, best_d2h, best_bidir, d2h_chunked, gran);
    for (int k = 0; k < 3; k++) {
        size_t sz = sizes[k]; CUdeviceptr va; CUmemGenericAllocationHandle h;
        CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
        double tc = 0, tm = 0, ta = 0, tu = 0, tr = 0; int reps = 20;
        for (int i = 0; i < reps; i++) {
            double t0 = now_us(); CK(cuMemCreate(&h, sz, &prop, 0));
            double t1 = now_us(); CK(cuMemMap(va, sz, 0, h, 0));
            double t2 = now_us(); CK(cuMemSetAccess(va, sz, &acc, 1));
            double t3 = now_us(); CK(cuMemUnmap(va, sz));
            double t4 = now_us(); CK(cuMemRelease(h));
            double t5 = now_us();
            tc += t1 - t0; tm += t2 - t1; ta += t3 - t2; tu += t4 - t3; tr += t5 - t4;
        }
        /* remap of a cached handle (what the engine does on the steady-state path) */
        CK(cuMemCreate(&h, sz, &prop, 0));
        double t0 = now_us();
        for (int i = 0; i < reps; i++) { CK(cuMemMap(va, sz, 0, h, 0)); CK(cuMemSetAccess(va, sz, &acc, 1)); CK(cuMemUnmap(va, sz)); }
        double remap = (now_us() - t0) / reps;
        CK(cuMemRelease(h)); CK(cuMemAddressFree(va, sz));
        printf("%s\"%zuMiB\": {\"create\": %.1f, \"map\": %.1f, \"setaccess\": %.1f, \"unmap\": %.1f, \"release\": %.1f, \"remap_cycle\": %.1f}",
               k ? ", " : "", sz >> 20, tc / reps, tm / reps, ta / reps, tu / reps, tr / reps, remap);
    }
    printf("}");
    /* the same VMM calls while a long DMA is in flight on another stream: a call that takes about as long as the
     * copy is implicitly synchronising the device, which would serialise the swap pipeline */
    {
        size_t sz = 64u << 20; CUdeviceptr va; CUmemGenericAllocationHandle h, h2;
        CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
        CK(cuMemCreate(&h2, sz, &prop, 0));
        double t[6] = {0, 0, 0, 0, 0, 0}; int reps = 5;
        for (int i = 0; i < reps; i++) {
            double a0, a1;
            #define UNDER_LOAD(idx, stmt) CK(cuCtxSynchronize()); CK(cuMemcpyDtoHAsync(h1, d1, bytes, s1)); CK(cuMemcpyHtoDAsync(d2, h2p, bytes, s2)); a0 = now_us(); stmt; a1 = now_us(); t[idx] += a1 - a0;
            void *h2p = h2p_buf(h2_host);
            (void)h2p;
            UNDER_LOAD(0, CK(cuMemCreate(&h, sz, &prop, 0)))
            UNDER_LOAD(1, CK(cuMemMap(va, sz, 0, h, 0)))
            UNDER_LOAD(2, CK(cuMemSetAccess(va, sz, &acc, 1)))
            UNDER_LOAD(3, CK(cuMemUnmap(va, sz)))
            UNDER_LOAD(4, CK(cuMemRelease(h)))
            CK(cuCtxSynchronize()); CK(cuMemcpyDtoHAsync(h1, d1, bytes, s1)); a0 = now_us(); CK(cuStreamSynchronize(s1)); a1 = now_us(); t[5] += a1 - a0;
        }
        CK(cuCtxSynchronize());
        printf(", \"vmm_us_under_dma_load\": {\"create\": %.1f, \"map\": %.1f, \"setaccess\": %.1f, \"unmap\": %.1f, \"release\": %.1f, \"dma_itself\": %.1f}",
               t[0] / reps, t[1] / reps, t[2] / reps, t[3] / reps, t[4] / reps, t[5] / reps);
        CK(cuMemRelease(h2)); CK(cuMemAddressFree(va, sz));
    }
    printf("}\n");
    return 0;
}
