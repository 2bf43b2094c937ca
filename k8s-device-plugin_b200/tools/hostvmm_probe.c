/*
 * hostvmm_probe.c — hardware facts the round-2 pager is designed around, measured on the box it runs on:
 *   A. host-located VMM (cuMemCreate with CU_MEM_LOCATION_TYPE_HOST_NUMA): can a device VA be backed by HOST memory, can
 *      a kernel dereference it, and how fast do the copy engines move data to / from it compared with cuMemHostAlloc?
 *      (the no-fault fallback for paged-out buffers: an evicted range is re-mapped onto host memory instead of left a hole)
 *   B. the same physical handle mapped at two VAs (alias) — lets an unmap be deferred and batched.
 *   C. batched VMM calls: ONE cuMemSetAccess / cuMemUnmap over a run of adjacent mappings vs one call per mapping,
 *      idle and under bidirectional DMA with 32 MiB and 8 MiB copies in flight (do VMM calls wait for the copy in flight?)
 *   D. a 64 MiB range assembled from 32 x 2 MiB handles (size-agnostic frame pool): map / setaccess / unmap cost.
 * Driver API only. Usage: hostvmm_probe <vgpu_kernels.cubin>. Output: JSON lines, one per experiment; a failing
 * experiment prints its error code and the probe goes on.
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { fprintf(stderr, "probe: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); exit(3); } } while (0)
#define TRY(x) ({ CUresult _r = (x); if (_r != CUDA_SUCCESS) fprintf(stderr, "probe: %s -> %d (line %d)\n", #x, (int)_r, __LINE__); _r; })
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }

static CUcontext g_ctx;
static CUfunction f_fill, f_verify, f_touch;
static CUmemAllocationProp dev_prop;
static CUmemAccessDesc dev_acc;

static int gpu_numa_node(CUdevice dev) {
    char bus[32] = {0}, path[128];
    if (cuDeviceGetPCIBusId(bus, sizeof bus, dev) != CUDA_SUCCESS) return 0;
    for (char *p = bus; *p; p++) if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    const char *bdf = strlen(bus) > 12 ? bus + strlen(bus) - 12 : bus;
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE *f = fopen(path, "r"); int node = 0;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = 0; fclose(f); }
    return node < 0 ? 0 : node;
}

/* ---- background bidirectional DMA load with a given copy size ---- */
struct load { volatile int stop; size_t chunk; double gbs; };
static void *load_thread(void *p) {
    struct load *L = p;
    CK(cuCtxSetCurrent(g_ctx));
    void *h1, *h2; CUdeviceptr d1, d2; CUstream s1, s2; CUevent e1[4], e2[4];
    CK(cuMemHostAlloc(&h1, L->chunk, 0)); CK(cuMemHostAlloc(&h2, L->chunk, 0));
    CK(cuMemAlloc(&d1, L->chunk)); CK(cuMemAlloc(&d2, L->chunk));
    CK(cuStreamCreate(&s1, CU_STREAM_NON_BLOCKING)); CK(cuStreamCreate(&s2, CU_STREAM_NON_BLOCKING));
    for (int i = 0; i < 4; i++) { CK(cuEventCreate(&e1[i], CU_EVENT_DISABLE_TIMING)); CK(cuEventCreate(&e2[i], CU_EVENT_DISABLE_TIMING)); }
    double t0 = now_us(); long n = 0;
    while (!L->stop) {   /* keep 4 copies queued per direction */
        int k = (int)(n & 3);
        if (n >= 4) { CK(cuEventSynchronize(e1[k])); CK(cuEventSynchronize(e2[k])); }
        CK(cuMemcpyHtoDAsync(d1, h1, L->chunk, s1)); CK(cuEventRecord(e1[k], s1));
        CK(cuMemcpyDtoHAsync(h2, d2, L->chunk, s2)); CK(cuEventRecord(e2[k], s2));
        n++;
    }
    CK(cuStreamSynchronize(s1)); CK(cuStreamSynchronize(s2));
    L->gbs = 2.0 * (double)L->chunk * (double)n / (now_us() - t0) / 1e3;
    cuMemFreeHost(h1); cuMemFreeHost(h2); cuMemFree(d1); cuMemFree(d2);
    return NULL;
}

/* ---- C: batched VMM ops over K adjacent 64 MiB mappings ---- */
static void batch_study(const char *label, int K, int reps) {
    size_t sz = 64u << 20; CUdeviceptr va; CUmemGenericAllocationHandle h[16];
    CK(cuMemAddressReserve(&va, sz * K, 0, 0, 0));
    for (int i = 0; i < K; i++) CK(cuMemCreate(&h[i], sz, &dev_prop, 0));
    double t_map = 0, t_acc1 = 0, t_accK = 0, t_un1 = 0, t_unK = 0, worst = 0; int unK_ok = 1;
    for (int r = 0; r < reps; r++) {
        double a, b;
        /* one call per mapping */
        a = now_us(); for (int i = 0; i < K; i++) CK(cuMemMap(va + i * sz, sz, 0, h[i], 0)); t_map += now_us() - a;
        a = now_us(); for (int i = 0; i < K; i++) CK(cuMemSetAccess(va + i * sz, sz, &dev_acc, 1)); b = now_us() - a; t_acc1 += b; if (b > worst) worst = b;
        a = now_us(); for (int i = 0; i < K; i++) CK(cuMemUnmap(va + i * sz, sz)); b = now_us() - a; t_un1 += b; if (b > worst) worst = b;
        /* one call for the run */
        for (int i = 0; i < K; i++) CK(cuMemMap(va + i * sz, sz, 0, h[i], 0));
        a = now_us(); CK(cuMemSetAccess(va, sz * K, &dev_acc, 1)); b = now_us() - a; t_accK += b; if (b > worst) worst = b;
        a = now_us();
        if (unK_ok && cuMemUnmap(va, sz * K) != CUDA_SUCCESS) { unK_ok = 0; }
        if (!unK_ok) for (int i = 0; i < K; i++) cuMemUnmap(va + i * sz, sz);
        b = now_us() - a; t_unK += b; if (b > worst) worst = b;
    }
    printf("{\"exp\": \"batch\", \"cond\": \"%s\", \"K\": %d, \"map_each_us\": %.1f, \"setaccess_each_us\": %.1f, \"setaccess_run_us\": %.1f, "
           "\"unmap_each_us\": %.1f, \"unmap_run_us\": %.1f, \"unmap_run_supported\": %s, \"worst_us\": %.1f}\n",
           label, K, t_map / reps / K, t_acc1 / reps / K, t_accK / reps, t_un1 / reps / K, t_unK / reps, unK_ok ? "true" : "false", worst);
    fflush(stdout);
    for (int i = 0; i < K; i++) cuMemRelease(h[i]);
    cuMemAddressFree(va, sz * K);
}

/* ---- D: 64 MiB range from 32 x 2 MiB frames ---- */
static void frames_study(const char *label, int reps) {
    size_t fsz = 2u << 20; int F = 32; CUdeviceptr va; CUmemGenericAllocationHandle h[32];
    CK(cuMemAddressReserve(&va, fsz * F, 0, 0, 0));
    for (int i = 0; i < F; i++) CK(cuMemCreate(&h[i], fsz, &dev_prop, 0));
    double t_map = 0, t_acc = 0, t_un = 0; int run_ok = 1;
    for (int r = 0; r < reps; r++) {
        double a = now_us(); for (int i = 0; i < F; i++) CK(cuMemMap(va + i * fsz, fsz, 0, h[i], 0)); t_map += now_us() - a;
        a = now_us(); CK(cuMemSetAccess(va, fsz * F, &dev_acc, 1)); t_acc += now_us() - a;
        a = now_us();
        if (run_ok && cuMemUnmap(va, fsz * F) != CUDA_SUCCESS) run_ok = 0;
        if (!run_ok) for (int i = 0; i < F; i++) cuMemUnmap(va + i * fsz, fsz);
        t_un += now_us() - a;
    }
    printf("{\"exp\": \"frames_2MiB_x32\", \"cond\": \"%s\", \"map_all_us\": %.1f, \"setaccess_us\": %.1f, \"unmap_us\": %.1f, \"unmap_run_supported\": %s}\n",
           label, t_map / reps, t_acc / reps, t_un / reps, run_ok ? "true" : "false");
    fflush(stdout);
    for (int i = 0; i < F; i++) cuMemRelease(h[i]);
    cuMemAddressFree(va, fsz * F);
}

static double copy_gbs(CUdeviceptr dst, CUdeviceptr src, size_t bytes, size_t chunk, CUstream s, int reps) {
    CUevent a, b; CK(cuEventCreate(&a, 0)); CK(cuEventCreate(&b, 0));
    double best = 0;
    for (int r = 0; r < reps; r++) {
        CK(cuEventRecord(a, s));
        for (size_t o = 0; o < bytes; o += chunk) CK(cuMemcpyDtoDAsync(dst + o, src + o, chunk, s));
        CK(cuEventRecord(b, s)); CK(cuStreamSynchronize(s));
        float ms; CK(cuEventElapsedTime(&ms, a, b));
        double g = bytes / ms / 1e6; if (g > best) best = g;
    }
    return best;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: hostvmm_probe <cubin>\n"); return 2; }
    CUdevice dev; CUmodule mod;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&g_ctx, dev)); CK(cuCtxSetCurrent(g_ctx));
    CK(cuModuleLoad(&mod, argv[1]));
    CK(cuModuleGetFunction(&f_fill, mod, "vgpu_wl_fill")); CK(cuModuleGetFunction(&f_verify, mod, "vgpu_wl_verify")); CK(cuModuleGetFunction(&f_touch, mod, "vgpu_wl_touch"));
    memset(&dev_prop, 0, sizeof dev_prop);
    dev_prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; dev_prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; dev_prop.location.id = 0;
    memset(&dev_acc, 0, sizeof dev_acc); dev_acc.location = dev_prop.location; dev_acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    int node = gpu_numa_node(dev);
    CUstream s; CK(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING));

    /* ------------------------------------------------------------------ A: host-located VMM */
    {
        CUmemAllocationProp hp; memset(&hp, 0, sizeof hp);
        hp.type = CU_MEM_ALLOCATION_TYPE_PINNED; hp.location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA; hp.location.id = node;
        size_t hgran = 0; CUresult rg = cuMemGetAllocationGranularity(&hgran, &hp, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
        size_t bytes = 1ull << 30, row = 64u << 20;
        CUmemGenericAllocationHandle hh = 0; CUdeviceptr hva = 0, dva = 0; CUmemGenericAllocationHandle dh = 0;
        double t0 = now_us();
        CUresult rc = cuMemCreate(&hh, bytes, &hp, 0);
        double t_create = now_us() - t0;
        printf("{\"exp\": \"host_vmm_create\", \"numa_node\": %d, \"granularity_rc\": %d, \"granularity\": %zu, \"create_rc\": %d, \"create_1GiB_us\": %.0f}\n", node, (int)rg, hgran, (int)rc, t_create);
        fflush(stdout);
        if (rc == CUDA_SUCCESS) {
            CK(cuMemAddressReserve(&hva, bytes, 0, 0, 0));
            CUresult rm = TRY(cuMemMap(hva, bytes, 0, hh, 0));
            CUmemAccessDesc acc2[2]; memset(acc2, 0, sizeof acc2);
            acc2[0] = dev_acc;
            acc2[1].location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA; acc2[1].location.id = node; acc2[1].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
            t0 = now_us();
            CUresult ra = rm == CUDA_SUCCESS ? TRY(cuMemSetAccess(hva, bytes, acc2, 2)) : rm;
            double t_acc = now_us() - t0;
            int host_access = ra == CUDA_SUCCESS;
            if (ra != CUDA_SUCCESS && rm == CUDA_SUCCESS) ra = TRY(cuMemSetAccess(hva, bytes, acc2, 1));   /* device only */
            printf("{\"exp\": \"host_vmm_map\", \"map_rc\": %d, \"setaccess_rc\": %d, \"with_host_access\": %s, \"setaccess_1GiB_us\": %.0f}\n", (int)rm, (int)ra, host_access ? "true" : "false", t_acc);
            fflush(stdout);
            if (ra == CUDA_SUCCESS) {
                /* device buffer as a VMM mapping too */
                CK(cuMemCreate(&dh, bytes, &dev_prop, 0)); CK(cuMemAddressReserve(&dva, bytes, 0, 0, 0));
                CK(cuMemMap(dva, bytes, 0, dh, 0)); CK(cuMemSetAccess(dva, bytes, &dev_acc, 1));
                void *pinned; CK(cuMemHostAlloc(&pinned, bytes, 0)); memset(pinned, 3, bytes);
                uint64_t nw = bytes / 8, idx = 7; void *fa[] = {&dva, &nw, &idx};
                CK(cuLaunchKernel(f_fill, 148 * 16, 1, 1, 256, 1, 1, 0, s, fa, 0)); CK(cuStreamSynchronize(s));
                /* copy engine: device -> host-VMM and back, 32 MiB pieces, vs cuMemHostAlloc */
                double out_v = copy_gbs(hva, dva, bytes, 32u << 20, s, 3);
                double in_v = copy_gbs(dva, hva, bytes, 32u << 20, s, 3);
                CUevent a, b; CK(cuEventCreate(&a, 0)); CK(cuEventCreate(&b, 0)); float ms;
                CK(cuEventRecord(a, s)); for (size_t o = 0; o < bytes; o += 32u << 20) CK(cuMemcpyDtoHAsync((char *)pinned + o, dva + o, 32u << 20, s));
                CK(cuEventRecord(b, s)); CK(cuStreamSynchronize(s)); CK(cuEventElapsedTime(&ms, a, b)); double out_p = bytes / ms / 1e6;
                CK(cuEventRecord(a, s)); for (size_t o = 0; o < bytes; o += 32u << 20) CK(cuMemcpyHtoDAsync(dva + o, (char *)pinned + o, 32u << 20, s));
                CK(cuEventRecord(b, s)); CK(cuStreamSynchronize(s)); CK(cuEventElapsedTime(&ms, a, b)); double in_p = bytes / ms / 1e6;
                /* refill (the HtoD above overwrote it), copy out to host-VMM, then let a KERNEL verify straight from the host-backed VA */
                CK(cuLaunchKernel(f_fill, 148 * 16, 1, 1, 256, 1, 1, 0, s, fa, 0));
                CK(cuMemcpyDtoDAsync(hva, dva, bytes, s));
                CUdeviceptr cnt; CK(cuMemAlloc(&cnt, 8)); CK(cuMemsetD8(cnt, 0, 8));
                uint64_t added = 0; void *va_[] = {&hva, &nw, &idx, &added, &cnt};
                CK(cuEventRecord(a, s));
                CUresult rk = TRY(cuLaunchKernel(f_verify, 148 * 16, 1, 1, 256, 1, 1, 0, s, va_, 0));
                CK(cuEventRecord(b, s));
                CUresult rs = TRY(cuStreamSynchronize(s));
                unsigned long long bad = ~0ull; float kms = 0;
                if (rk == CUDA_SUCCESS && rs == CUDA_SUCCESS) { CK(cuMemcpyDtoH(&bad, cnt, 8)); CK(cuEventElapsedTime(&kms, a, b)); }
                /* kernel WRITES through the host-backed VA (RMW), verified again with added = 1 */
                unsigned long long bad2 = ~0ull; float tms = 0;
                if (rs == CUDA_SUCCESS) {
                    void *ta[] = {&hva, &nw};
                    CK(cuMemsetD8(cnt, 0, 8));
                    CK(cuEventRecord(a, s));
                    CK(cuLaunchKernel(f_touch, 148 * 16, 1, 1, 256, 1, 1, 0, s, ta, 0));
                    CK(cuEventRecord(b, s));
                    added = 1;
                    CK(cuLaunchKernel(f_verify, 148 * 16, 1, 1, 256, 1, 1, 0, s, va_, 0));
                    if (TRY(cuStreamSynchronize(s)) == CUDA_SUCCESS) { CK(cuMemcpyDtoH(&bad2, cnt, 8)); CK(cuEventElapsedTime(&tms, a, b)); }
                }
                /* CPU view of the same memory (only when host access was granted) */
                unsigned long long cpu_word = 0;
                if (host_access) cpu_word = ((volatile unsigned long long *)(uintptr_t)hva)[5];
                printf("{\"exp\": \"host_vmm_data\", \"d2hostvmm_gbs\": %.1f, \"hostvmm2d_gbs\": %.1f, \"d2pinned_gbs\": %.1f, \"pinned2d_gbs\": %.1f, "
                       "\"kernel_read_rc\": %d, \"kernel_read_mismatches\": %llu, \"kernel_read_1GiB_ms\": %.2f, \"kernel_rmw_mismatches\": %llu, \"kernel_rmw_1GiB_ms\": %.2f, \"cpu_word5\": %llu}\n",
                       out_v, in_v, out_p, in_p, (int)(rk != CUDA_SUCCESS ? rk : rs), bad, kms, bad2, tms, cpu_word);
                fflush(stdout);
                /* the remap sequence of an eviction with host backing: row VA (device handle) -> host handle, and back */
                {
                    CUmemAllocationProp hp2 = hp; CUmemGenericAllocationHandle rh, rd; CUdeviceptr rva;
                    CK(cuMemAddressReserve(&rva, row, 0, 0, 0));
                    if (TRY(cuMemCreate(&rh, row, &hp2, 0)) == CUDA_SUCCESS) {
                        CK(cuMemCreate(&rd, row, &dev_prop, 0));
                        CK(cuMemMap(rva, row, 0, rd, 0)); CK(cuMemSetAccess(rva, row, &dev_acc, 1));
                        double t_to_host = 0, t_to_dev = 0; int reps = 50;
                        for (int r = 0; r < reps; r++) {
                            double t = now_us();
                            CK(cuMemUnmap(rva, row)); CK(cuMemMap(rva, row, 0, rh, 0)); CK(cuMemSetAccess(rva, row, &dev_acc, 1));
                            t_to_host += now_us() - t; t = now_us();
                            CK(cuMemUnmap(rva, row)); CK(cuMemMap(rva, row, 0, rd, 0)); CK(cuMemSetAccess(rva, row, &dev_acc, 1));
                            t_to_dev += now_us() - t;
                        }
                        printf("{\"exp\": \"remap_row_64MiB_idle\", \"dev_to_host_backing_us\": %.1f, \"host_to_dev_backing_us\": %.1f}\n", t_to_host / reps, t_to_dev / reps);
                        fflush(stdout);
                        cuMemUnmap(rva, row); cuMemRelease(rh); cuMemRelease(rd);
                    }
                    cuMemAddressFree(rva, row);
                }
                cuMemUnmap(dva, bytes); cuMemRelease(dh); cuMemAddressFree(dva, bytes); cuMemFreeHost(pinned);
            }
            if (rm == CUDA_SUCCESS) cuMemUnmap(hva, bytes);
            cuMemAddressFree(hva, bytes);
            cuMemRelease(hh);
        }
    }

    /* ------------------------------------------------------------------ B: alias (one handle, two VAs) */
    {
        size_t sz = 64u << 20; CUmemGenericAllocationHandle h; CUdeviceptr v1, v2;
        CK(cuMemCreate(&h, sz, &dev_prop, 0)); CK(cuMemAddressReserve(&v1, sz, 0, 0, 0)); CK(cuMemAddressReserve(&v2, sz, 0, 0, 0));
        CUresult r1 = TRY(cuMemMap(v1, sz, 0, h, 0)), r2 = TRY(cuMemMap(v2, sz, 0, h, 0));
        CUresult a1 = TRY(cuMemSetAccess(v1, sz, &dev_acc, 1)), a2 = r2 == CUDA_SUCCESS ? TRY(cuMemSetAccess(v2, sz, &dev_acc, 1)) : r2;
        unsigned long long bad = ~0ull;
        if (a1 == CUDA_SUCCESS && a2 == CUDA_SUCCESS) {
            uint64_t nw = sz / 8, idx = 3, added = 0; void *fa[] = {&v1, &nw, &idx};
            CUdeviceptr cnt; CK(cuMemAlloc(&cnt, 8)); CK(cuMemsetD8(cnt, 0, 8));
            void *va_[] = {&v2, &nw, &idx, &added, &cnt};
            CK(cuLaunchKernel(f_fill, 148 * 16, 1, 1, 256, 1, 1, 0, s, fa, 0));
            CK(cuLaunchKernel(f_verify, 148 * 16, 1, 1, 256, 1, 1, 0, s, va_, 0));
            CK(cuStreamSynchronize(s)); CK(cuMemcpyDtoH(&bad, cnt, 8));
        }
        printf("{\"exp\": \"alias\", \"map1_rc\": %d, \"map2_rc\": %d, \"access_rc\": [%d, %d], \"write_v1_read_v2_mismatches\": %llu}\n", (int)r1, (int)r2, (int)a1, (int)a2, bad);
        fflush(stdout);
        if (r1 == CUDA_SUCCESS) cuMemUnmap(v1, sz);
        if (r2 == CUDA_SUCCESS) cuMemUnmap(v2, sz);
        cuMemRelease(h); cuMemAddressFree(v1, sz); cuMemAddressFree(v2, sz);
    }

    /* ------------------------------------------------------------------ C + D: batched ops, idle and under DMA load */
    batch_study("idle", 4, 30);
    batch_study("idle", 8, 30);
    frames_study("idle", 30);
    size_t chunks[] = {32u << 20, 8u << 20, 2u << 20};
    for (int c = 0; c < 3; c++) {
        struct load L = {0, chunks[c], 0}; pthread_t th; char label[64];
        pthread_create(&th, NULL, load_thread, &L);
        struct timespec ts = {0, 200000000}; nanosleep(&ts, NULL);
        snprintf(label, sizeof label, "bidir DMA, %zu MiB copies in flight", chunks[c] >> 20);
        batch_study(label, 4, 30);
        if (c == 0) frames_study(label, 20);
        L.stop = 1; pthread_join(th, NULL);
        printf("{\"exp\": \"load\", \"chunk_mib\": %zu, \"bidir_gbs_while_remapping\": %.1f}\n", chunks[c] >> 20, L.gbs);
        fflush(stdout);
    }
    /* kernels (not copies) in flight: does a VMM call wait for a running kernel? */
    {
        size_t sz = 1ull << 30; CUdeviceptr d; CK(cuMemAlloc(&d, sz)); uint64_t nw = sz / 8; void *ta[] = {&d, &nw};
        CUstream ks; CK(cuStreamCreate(&ks, CU_STREAM_NON_BLOCKING));
        for (int i = 0; i < 400; i++) CK(cuLaunchKernel(f_touch, 148 * 16, 1, 1, 256, 1, 1, 0, ks, ta, 0));   /* ~0.35 ms each */
        batch_study("touch kernels (1 GiB RMW each) in flight", 4, 20);
        CK(cuStreamSynchronize(ks)); cuMemFree(d);
    }
    return 0;
}
