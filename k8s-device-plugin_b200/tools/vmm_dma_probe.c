/*
 * vmm_dma_probe.c — do VMM remaps (cuMemUnmap / cuMemMap / cuMemSetAccess) of an unrelated range slow down, or get
 * slowed down by, pinned DMA running in both directions? The swap engine issues exactly that mix once per paged
 * buffer; this isolates the interaction. Output: one JSON object.
 */
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "probe: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); exit(3); } } while (0)
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }

static double run(int with_vmm, int with_dma, size_t chunk, int nchunks, double *vmm_us_avg, int *vmm_ops, size_t remap_bytes) {
    static void *h_in, *h_out; static CUdeviceptr d_in, d_out, va; static CUstream s_in, s_out; static CUevent a, b, c, d;
    static CUmemGenericAllocationHandle hnd[2]; static int init;
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (!init) {
        CK(cuMemHostAlloc(&h_in, chunk, CU_MEMHOSTALLOC_PORTABLE)); CK(cuMemHostAlloc(&h_out, chunk, CU_MEMHOSTALLOC_PORTABLE));
        memset(h_in, 1, chunk); memset(h_out, 2, chunk);
        CK(cuMemAlloc(&d_in, chunk)); CK(cuMemAlloc(&d_out, chunk));
        CK(cuStreamCreate(&s_in, CU_STREAM_NON_BLOCKING)); CK(cuStreamCreate(&s_out, CU_STREAM_NON_BLOCKING));
        CK(cuEventCreate(&a, 0)); CK(cuEventCreate(&b, 0)); CK(cuEventCreate(&c, 0)); CK(cuEventCreate(&d, 0));
        CK(cuMemAddressReserve(&va, remap_bytes, 0, 0, 0));
        CK(cuMemCreate(&hnd[0], remap_bytes, &prop, 0)); CK(cuMemCreate(&hnd[1], remap_bytes, &prop, 0));
        CK(cuMemMap(va, remap_bytes, 0, hnd[0], 0)); CK(cuMemSetAccess(va, remap_bytes, &acc, 1));
        init = 1;
    }
    CK(cuCtxSynchronize());
    double t0 = now_us();
    if (with_dma) {
        CK(cuEventRecord(a, s_in)); CK(cuEventRecord(c, s_out));
        for (int i = 0; i < nchunks; i++) { CK(cuMemcpyHtoDAsync(d_in, h_in, chunk, s_in)); CK(cuMemcpyDtoHAsync(h_out, d_out, chunk, s_out)); }
        CK(cuEventRecord(b, s_in)); CK(cuEventRecord(d, s_out));
    }
    int ops = 0; double vmm_total = 0; int cur = 0;
    if (with_vmm) {
        double limit = with_dma ? 1e18 : 400e3;   /* without DMA: spin for 0.4 s */
        while (1) {
            if (with_dma && cuEventQuery(b) == CUDA_SUCCESS && cuEventQuery(d) == CUDA_SUCCESS) break;
            if (!with_dma && now_us() - t0 > limit) break;
            double v0 = now_us();
            CK(cuMemUnmap(va, remap_bytes)); cur ^= 1;
            CK(cuMemMap(va, remap_bytes, 0, hnd[cur], 0)); CK(cuMemSetAccess(va, remap_bytes, &acc, 1));
            vmm_total += now_us() - v0; ops++;
        }
    }
    CK(cuCtxSynchronize());
    double gbs = 0;
    if (with_dma) { float m1, m2; CK(cuEventElapsedTime(&m1, a, b)); CK(cuEventElapsedTime(&m2, c, d)); double ms = m1 > m2 ? m1 : m2; gbs = 2.0 * chunk * nchunks / ms / 1e6; }
    *vmm_us_avg = ops ? vmm_total / ops : 0; *vmm_ops = ops;
    return gbs;
}

/* do VMM calls from several threads scale, or does the driver serialise them? each thread remaps its own range */
static CUcontext g_ctx;
struct targ { int reps; size_t bytes; double us_per_cycle; int split; };
static void *remap_thread(void *p) {
    struct targ *a = p;
    CK(cuCtxSetCurrent(g_ctx));
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUdeviceptr va; CUmemGenericAllocationHandle h;
    CK(cuMemAddressReserve(&va, a->bytes, 0, 0, 0)); CK(cuMemCreate(&h, a->bytes, &prop, 0));
    CK(cuMemMap(va, a->bytes, 0, h, 0)); CK(cuMemSetAccess(va, a->bytes, &acc, 1));
    double t0 = now_us();
    for (int i = 0; i < a->reps; i++) {
        if (a->split == 1) { CK(cuMemUnmap(va, a->bytes)); CK(cuMemMap(va, a->bytes, 0, h, 0)); }            /* unmap+map only */
        else if (a->split == 2) { CK(cuMemSetAccess(va, a->bytes, &acc, 1)); }                                /* setaccess only */
        else { CK(cuMemUnmap(va, a->bytes)); CK(cuMemMap(va, a->bytes, 0, h, 0)); CK(cuMemSetAccess(va, a->bytes, &acc, 1)); }
    }
    a->us_per_cycle = (now_us() - t0) / a->reps;
    CK(cuMemUnmap(va, a->bytes)); CK(cuMemRelease(h)); CK(cuMemAddressFree(va, a->bytes));
    return NULL;
}
static double threads_cycle_us(int nthreads, size_t bytes, int split) {
    pthread_t th[8]; struct targ a[8]; double worst = 0;
    for (int i = 0; i < nthreads; i++) { a[i].reps = 200; a[i].bytes = bytes; a[i].split = split; pthread_create(&th[i], NULL, remap_thread, &a[i]); }
    for (int i = 0; i < nthreads; i++) { pthread_join(th[i], NULL); if (a[i].us_per_cycle > worst) worst = a[i].us_per_cycle; }
    return worst;
}

/* what does a remap cycle cost as a function of where the range sits in the page-table hierarchy, and of whether the
 * same or another physical handle comes back? */
static void offset_study(void) {
    CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t span = 4ull << 30, sz = 64u << 20; CUdeviceptr base; CUmemGenericAllocationHandle h[2];
    CK(cuMemAddressReserve(&base, span, span, 0, 0));
    CK(cuMemCreate(&h[0], sz, &prop, 0)); CK(cuMemCreate(&h[1], sz, &prop, 0));
    size_t offs[] = {0, 2u << 20, 32u << 20, 480u << 20, (1024u - 32u) << 20, (2048u - 2u) << 20};
    printf("\"offset_study_us\": {\"base_mod_4GiB\": %llu", (unsigned long long)(base & (span - 1)));
    for (int alt = 0; alt < 2; alt++)
        for (unsigned k = 0; k < sizeof offs / sizeof offs[0]; k++) {
            CUdeviceptr va = base + offs[k]; int cur = 0;
            CK(cuMemMap(va, sz, 0, h[0], 0)); CK(cuMemSetAccess(va, sz, &acc, 1));
            for (int i = 0; i < 20; i++) { CK(cuMemUnmap(va, sz)); if (alt) cur ^= 1; CK(cuMemMap(va, sz, 0, h[cur], 0)); CK(cuMemSetAccess(va, sz, &acc, 1)); }
            double t0 = now_us(); int reps = 100;
            for (int i = 0; i < reps; i++) { CK(cuMemUnmap(va, sz)); if (alt) cur ^= 1; CK(cuMemMap(va, sz, 0, h[cur], 0)); CK(cuMemSetAccess(va, sz, &acc, 1)); }
            printf(", \"%s@%zuMiB\": %.1f", alt ? "alt" : "same", offs[k] >> 20, (now_us() - t0) / reps);
            CK(cuMemUnmap(va, sz));
        }
    /* two DIFFERENT ranges alternately (the engine's pattern: unmap victim range, map+setaccess incoming range) */
    {
        CUdeviceptr va1 = base + (256u << 20), va2 = base + (1280u << 20);
        CK(cuMemMap(va1, sz, 0, h[0], 0)); CK(cuMemSetAccess(va1, sz, &acc, 1));
        double t0 = now_us(); int reps = 100;
        for (int i = 0; i < reps; i++) {
            CK(cuMemUnmap(va1, sz)); CK(cuMemMap(va2, sz, 0, h[0], 0)); CK(cuMemSetAccess(va2, sz, &acc, 1));
            CK(cuMemUnmap(va2, sz)); CK(cuMemMap(va1, sz, 0, h[0], 0)); CK(cuMemSetAccess(va1, sz, &acc, 1));
        }
        printf(", \"pingpong_two_ranges_per_remap\": %.1f", (now_us() - t0) / reps / 2);
        CK(cuMemUnmap(va1, sz));
    }
    printf("}, ");
    CK(cuMemRelease(h[0])); CK(cuMemRelease(h[1])); CK(cuMemAddressFree(base, span));
}

int main(void) {
    CUdevice dev; CUcontext ctx;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    g_ctx = ctx;
    double t1 = threads_cycle_us(1, 64u << 20, 0), t2 = threads_cycle_us(2, 64u << 20, 0), t4 = threads_cycle_us(4, 64u << 20, 0);
    double um = threads_cycle_us(1, 64u << 20, 1), sa = threads_cycle_us(1, 64u << 20, 2);
    double small = threads_cycle_us(1, 2u << 20, 0), big = threads_cycle_us(1, 1024u << 20, 0);
    printf("{");
    offset_study();
    size_t chunk = 32u << 20; int n = 1200; double v; int ops;
    double base = run(0, 1, chunk, n, &v, &ops, 64u << 20);
    double v_idle; int ops_idle; run(1, 0, chunk, n, &v_idle, &ops_idle, 64u << 20);
    double v_load; int ops_load; double loaded = run(1, 1, chunk, n, &v_load, &ops_load, 64u << 20);
    printf("\"dma_bidir_gbs_alone\": %.1f, \"dma_bidir_gbs_with_vmm_remaps\": %.1f, \"remap_us_idle\": %.1f, \"remap_us_under_dma\": %.1f, "
           "\"remaps_during_dma\": %d, \"chunk_mib\": %zu, \"remap_mib\": 64, "
           "\"remap_cycle_us_1_2_4_threads\": [%.1f, %.1f, %.1f], \"unmap_map_us\": %.1f, \"setaccess_us\": %.1f, \"cycle_us_2MiB\": %.1f, \"cycle_us_1GiB\": %.1f}\n",
           base, loaded, v_idle, v_load, ops_load, chunk >> 20, t1, t2, t4, um, sa, small, big);
    return 0;
}
