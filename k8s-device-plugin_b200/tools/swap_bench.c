/*
 * swap_bench.c — the oversubscribed alloc+touch loop of SURVEY.md §8d cfg 3 / BASELINE.json configs[2], written as a
 * plain CUDA-driver-API application (no knowledge of the hook): it calls cuMemAlloc_v2 / cuLaunchKernel and nothing
 * else, so the same binary measures
 *    - the new library            (LD_PRELOAD=libvgpu.so, CUDA_OVERSUBSCRIBE=true, CUDA_DEVICE_MEMORY_LIMIT_0=<quota>)
 *    - the reference hook         (LD_PRELOAD=dlsym_shim.so:oracle/_ref/libvgpu.so, same env; its swap = UVM)
 *    - the reference's swap path without the hook (--managed 1: cuMemAllocManaged directly, what
 *      cuMemoryAllocate libvgpu.so@0x315da does in allocmode 0), with --ballast-mib forcing physical pressure.
 * Output: one JSON object on stdout.
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) { const char *s = 0; cuGetErrorName(_r, &s); \
    fprintf(stderr, "swap_bench: %s -> %d %s (line %d)\n", #x, (int)_r, s ? s : "?", __LINE__); printf("{\"error\": \"%s rc=%d\"}\n", #x, (int)_r); exit(3); } } while (0)

/* mirrors vgpu_swap_stats_t (include/vgpu.h) */
typedef struct {
    uint64_t v[17]; double pack_ms, unpack_ms; uint64_t scan_cache_hits, host_admit_ns, host_wait_ns, host_vmm_ns;
    uint64_t pager_vmm_ns, pager_scan_ns, pager_packsync_ns, pager_ring_ns, pager_busy_ns, vmm_calls;
    double pack_span_ms, unpack_span_ms;
    uint64_t direct_out_bytes, direct_in_bytes, prefetch_issued, prefetch_hits, prefetch_wasted, demand_waits, clean_evictions, host_slabs, host_slabs_local;
    uint64_t pager_unmap_ns, pager_setaccess_ns, pager_issue_ns, pager_poll_ns, pager_lock_ns, pager_step_ns[5];
    uint64_t vmm_slow_calls, vmm_slow_ns, vmm_max_ns;
    uint64_t inplace_uses;
} swap_stats_t;
typedef int (*stats_fn)(int, swap_stats_t *);
typedef int (*prof_fn)(int, int);

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6; }
static uint64_t rng_state = 0x5EED;
static uint64_t rng(void) { uint64_t x = (rng_state += 0x9E3779B97F4A7C15ull); x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

int main(int argc, char **argv) {
    const char *cubin = NULL, *order = "cyclic";
    long nbuf = 64, mib = 64, steps = 64, warmup = 8, managed = 0, ballast_mib = 0, verify = 1, profile = 0, seed = 0x5EED, wait_stdin = 0;
    long ro_every = 0;                   /* > 0: every ro_every-th buffer is advised read-mostly (cuMemAdvise) and only ever READ after the fill */
    long ragged_lo = 0, ragged_hi = 0;   /* cfg 3 variant B: sizes log-uniform in [lo, hi] MiB, same total as --buffers x --mib */
    double zipf_s = 1.1;
    for (int i = 1; i + 1 < argc; i += 2) {
        const char *k = argv[i], *v = argv[i + 1];
        if (!strcmp(k, "--cubin")) cubin = v;
        else if (!strcmp(k, "--buffers")) nbuf = atol(v);
        else if (!strcmp(k, "--mib")) mib = atol(v);
        else if (!strcmp(k, "--steps")) steps = atol(v);
        else if (!strcmp(k, "--warmup")) warmup = atol(v);
        else if (!strcmp(k, "--order")) order = v;
        else if (!strcmp(k, "--managed")) managed = atol(v);
        else if (!strcmp(k, "--ballast-mib")) ballast_mib = atol(v);
        else if (!strcmp(k, "--verify")) verify = atol(v);
        else if (!strcmp(k, "--profile")) profile = atol(v);
        else if (!strcmp(k, "--seed")) seed = strtol(v, 0, 0);
        else if (!strcmp(k, "--zipf")) zipf_s = atof(v);
        else if (!strcmp(k, "--wait-stdin")) wait_stdin = atol(v);
        else if (!strcmp(k, "--ro-every")) ro_every = atol(v);
        else if (!strcmp(k, "--ragged-lo")) ragged_lo = atol(v);
        else if (!strcmp(k, "--ragged-hi")) ragged_hi = atol(v);
        else { fprintf(stderr, "unknown option %s\n", k); return 2; }
    }
    if (getenv("SWAP_BENCH_HOLD_MIB")) {
        /* ballast holder: an UNHOOKED helper process that pins physical memory with plain cuMemAlloc so that the
         * reference hook's UVM allocations in another process see the same physical budget as the quota (the reference
         * turns every large cuMemAlloc of its own process into managed memory, so the ballast cannot live there) */
        CUdevice hd; CUcontext hc; CUdeviceptr hp; char line[8];
        CK(cuInit(0)); CK(cuDeviceGet(&hd, 0)); CK(cuDevicePrimaryCtxRetain(&hc, hd)); CK(cuCtxSetCurrent(hc));
        CK(cuMemAlloc(&hp, (size_t)atol(getenv("SWAP_BENCH_HOLD_MIB")) << 20));
        CK(cuMemsetD8(hp, 0, 1 << 20));
        fprintf(stderr, "READY\n"); fflush(stderr);
        if (!fgets(line, sizeof line, stdin)) return 0;
        return 0;
    }
    if (!cubin) { fprintf(stderr, "--cubin required\n"); return 2; }
    rng_state = (uint64_t)seed;
    /* buffer sizes: uniform (variant A), or log-uniform in [lo, hi] MiB rounded to 256 B until the same total is
     * reached (variant B, SURVEY.md §8d cfg 3) */
    size_t *sz;
    if (ragged_lo > 0 && ragged_hi >= ragged_lo) {
        const size_t want = (size_t)nbuf * ((size_t)mib << 20);
        size_t cap = (size_t)(want / ((size_t)ragged_lo << 20)) + 2, sum = 0;
        long n = 0;
        sz = calloc(cap, sizeof *sz);
        while (sum < want && (size_t)n < cap) {
            double u = (double)(rng() >> 11) / 9007199254740992.0;
            double m = (double)ragged_lo * pow((double)ragged_hi / (double)ragged_lo, u);
            size_t b = ((size_t)(m * 1048576.0) + 255) & ~(size_t)255;
            if (sum + b > want && want - sum >= ((size_t)ragged_lo << 20)) b = want - sum;
            sz[n++] = b; sum += b;
        }
        nbuf = n;
    } else {
        sz = calloc((size_t)nbuf, sizeof *sz);
        for (long i = 0; i < nbuf; i++) sz[i] = (size_t)mib << 20;
    }

    CUdevice dev; CUcontext ctx; CUmodule mod; CUfunction f_fill, f_touch, f_verify;
    CK(cuInit(0));
    CK(cuDeviceGet(&dev, 0));
    CK(cuDevicePrimaryCtxRetain(&ctx, dev));
    CK(cuCtxSetCurrent(ctx));
    CK(cuModuleLoad(&mod, cubin));
    CK(cuModuleGetFunction(&f_fill, mod, "vgpu_wl_fill"));
    CK(cuModuleGetFunction(&f_touch, mod, "vgpu_wl_touch"));
    CK(cuModuleGetFunction(&f_verify, mod, "vgpu_wl_verify"));
    int sm = 148; cuDeviceGetAttribute(&sm, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    unsigned grid = (unsigned)sm * 16;

    CUdeviceptr ballast = 0;
    if (ballast_mib > 0) CK(cuMemAlloc(&ballast, (size_t)ballast_mib << 20));

    stats_fn get_stats = (stats_fn)dlsym(RTLD_DEFAULT, "vgpu_runtime_swap_stats");
    prof_fn set_prof = (prof_fn)dlsym(RTLD_DEFAULT, "vgpu_runtime_set_swap_profile");

    CUdeviceptr *buf = calloc((size_t)nbuf, sizeof *buf);
    uint64_t *touches = calloc((size_t)nbuf, sizeof *touches);
    double t_alloc0 = now_ms();
    for (long i = 0; i < nbuf; i++) {
        if (managed) CK(cuMemAllocManaged(&buf[i], sz[i], CU_MEM_ATTACH_GLOBAL));
        else CK(cuMemAlloc(&buf[i], sz[i]));
        uint64_t idx = (uint64_t)i, nw = sz[i] / 8;
        void *a[] = {&buf[i], &nw, &idx};
        CK(cuLaunchKernel(f_fill, grid, 1, 1, 256, 1, 1, 0, 0, a, 0));
    }
    CK(cuCtxSynchronize());
    /* read-mostly buffers: the hint a UVM application gives (the reference's swappable memory IS managed memory); a plain
     * cuMemAlloc pointer makes the bare driver refuse it, which is fine */
    if (ro_every > 0) for (long i = ro_every - 1; i < nbuf; i += ro_every) (void)cuMemAdvise(buf[i], sz[i], CU_MEM_ADVISE_SET_READ_MOSTLY, dev);
    CUdeviceptr d_ro_bad = 0; CK(cuMemAlloc(&d_ro_bad, 8)); CK(cuMemsetD8(d_ro_bad, 0, 8));
    double t_alloc1 = now_ms();
    if (profile && set_prof) set_prof(0, 1);

    /* touch order */
    long total = warmup + steps;
    long *seq = malloc((size_t)total * sizeof *seq);
    if (!strcmp(order, "zipf")) {
        double *cdf = malloc((size_t)nbuf * sizeof *cdf); double z = 0;
        for (long i = 0; i < nbuf; i++) { z += 1.0 / pow((double)(i + 1), zipf_s); cdf[i] = z; }
        for (long t = 0; t < total; t++) {
            double u = (double)(rng() >> 11) / 9007199254740992.0 * z;
            long lo = 0, hi = nbuf - 1;
            while (lo < hi) { long m = (lo + hi) / 2; if (cdf[m] < u) lo = m + 1; else hi = m; }
            seq[t] = lo;
        }
        free(cdf);
    } else {
        /* cyclic continues after the fill order: LRU worst case, every touch misses once the set exceeds the quota */
        for (long t = 0; t < total; t++) seq[t] = t % nbuf;
    }

    CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
    swap_stats_t s0, s1; memset(&s0, 0, sizeof s0); memset(&s1, 0, sizeof s1);
#define IS_RO(i) (ro_every > 0 && ((i) + 1) % ro_every == 0)
#define TOUCH(i) do { uint64_t nw_ = sz[i] / 8; \
        if (IS_RO(i)) { uint64_t idx_ = (uint64_t)(i), add_ = 0; void *a_[] = {&buf[i], &nw_, &idx_, &add_, &d_ro_bad}; CK(cuLaunchKernel(f_verify, grid, 1, 1, 256, 1, 1, 0, 0, a_, 0)); } \
        else { void *a_[] = {&buf[i], &nw_}; CK(cuLaunchKernel(f_touch, grid, 1, 1, 256, 1, 1, 0, 0, a_, 0)); touches[i]++; } } while (0)
    for (long t = 0; t < warmup; t++) TOUCH(seq[t]);
    CK(cuCtxSynchronize());
    if (wait_stdin) {   /* multi-GPU runs: the launcher releases every rank's timed region together */
        char line[16];
        fprintf(stderr, "READY\n"); fflush(stderr);
        if (!fgets(line, sizeof line, stdin)) return 5;
    }
    if (get_stats) get_stats(0, &s0);
    double w0 = now_ms();
    CK(cuEventRecord(e0, 0));
    unsigned long long touched = 0;
    for (long t = warmup; t < total; t++) { TOUCH(seq[t]); touched += sz[seq[t]]; }
    double w_enq = now_ms();
    CK(cuEventRecord(e1, 0));
    CK(cuCtxSynchronize());
    double w1 = now_ms();
    float ev_ms = 0; CK(cuEventElapsedTime(&ev_ms, e0, e1));
    if (get_stats) get_stats(0, &s1);

    unsigned long long mism = 0;
    double t_ver0 = now_ms();
    if (verify) {
        CUdeviceptr dcnt; CK(cuMemAlloc(&dcnt, 8)); CK(cuMemsetD8(dcnt, 0, 8));
        for (long i = 0; i < nbuf; i++) {
            uint64_t nw = sz[i] / 8, idx = (uint64_t)i, add = touches[i];
            void *a[] = {&buf[i], &nw, &idx, &add, &dcnt};
            CK(cuLaunchKernel(f_verify, grid, 1, 1, 256, 1, 1, 0, 0, a, 0));
        }
        CK(cuCtxSynchronize());
        CK(cuMemcpyDtoH(&mism, dcnt, 8));
    }
    { unsigned long long ro_bad = 0; CK(cuMemcpyDtoH(&ro_bad, d_ro_bad, 8)); mism += ro_bad; }   /* what the read-only touches saw */
    double t_ver1 = now_ms();

    uint64_t pin = s1.v[1] - s0.v[1], pout = s1.v[0] - s0.v[0];
    printf("{\"ro_every\": %ld, \"ragged_mib\": [%ld, %ld], \"buffers\": %ld, \"mib\": %ld, \"steps\": %ld, \"warmup\": %ld, \"order\": \"%s\", \"managed\": %ld, \"ballast_mib\": %ld, "
           "\"event_ms\": %.3f, \"wall_ms\": %.3f, \"enqueue_ms\": %.3f, \"alloc_fill_ms\": %.1f, \"verify_ms\": %.1f, "
           "\"hooked_stats\": %s, \"page_in_bytes\": %llu, \"page_out_bytes\": %llu, \"touched_bytes\": %llu, "
           "\"mismatches\": %llu, \"verified\": %ld, "
           "\"pack_ms\": %.3f, \"unpack_ms\": %.3f, \"pack_bytes\": %llu, \"unpack_bytes\": %llu, "
           "\"pack_launches\": %llu, \"unpack_launches\": %llu, \"scan_launches\": %llu, \"faults\": %llu, \"evictions\": %llu, "
           "\"phys_creates\": %llu, \"phys_reuses\": %llu, \"scans\": %llu, \"scan_cache_hits\": %llu, "
           "\"host_ms\": {\"admit\": %.1f, \"wait\": %.1f, \"vmm\": %.1f}, "
           "\"pager_ms\": {\"busy\": %.1f, \"vmm\": %.1f, \"scan\": %.1f, \"packsync\": %.1f, \"ringwait\": %.1f, \"unmap\": %.1f, \"setaccess\": %.1f, \"issue\": %.1f, \"poll\": %.1f, \"lock\": %.1f, "
           "\"steps\": [%.1f, %.1f, %.1f, %.1f, %.1f]}, \"vmm_slow\": {\"calls\": %llu, \"ms\": %.1f, \"max_ms\": %.1f}, \"vmm_calls\": %llu, "
           "\"direct_in_bytes\": %llu, \"direct_out_bytes\": %llu, \"prefetch\": {\"issued\": %llu, \"hits\": %llu, \"wasted\": %llu}, "
           "\"demand_waits\": %llu, \"clean_evictions\": %llu, \"host_slabs\": [%llu, %llu], "
           "\"pack_span_ms\": %.3f, \"unpack_span_ms\": %.3f}\n",
           ro_every, ragged_lo, ragged_hi, nbuf, mib, steps, warmup, order, managed, ballast_mib, ev_ms, w1 - w0, w_enq - w0, t_alloc1 - t_alloc0, t_ver1 - t_ver0,
           get_stats ? "true" : "false", (unsigned long long)pin, (unsigned long long)pout,
           touched, mism, verify,
           s1.pack_ms - s0.pack_ms, s1.unpack_ms - s0.unpack_ms,
           (unsigned long long)(s1.v[15] - s0.v[15]), (unsigned long long)(s1.v[16] - s0.v[16]),
           (unsigned long long)(s1.v[5] - s0.v[5]), (unsigned long long)(s1.v[6] - s0.v[6]), (unsigned long long)(s1.v[7] - s0.v[7]),
           (unsigned long long)(s1.v[3] - s0.v[3]), (unsigned long long)(s1.v[2] - s0.v[2]),
           (unsigned long long)(s1.v[13] - s0.v[13]), (unsigned long long)(s1.v[14] - s0.v[14]),
           (unsigned long long)(s1.v[8] - s0.v[8]), (unsigned long long)(s1.scan_cache_hits - s0.scan_cache_hits),
           (s1.host_admit_ns - s0.host_admit_ns) / 1e6, (s1.host_wait_ns - s0.host_wait_ns) / 1e6, (s1.host_vmm_ns - s0.host_vmm_ns) / 1e6,
           (s1.pager_busy_ns - s0.pager_busy_ns) / 1e6, (s1.pager_vmm_ns - s0.pager_vmm_ns) / 1e6, (s1.pager_scan_ns - s0.pager_scan_ns) / 1e6,
           (s1.pager_packsync_ns - s0.pager_packsync_ns) / 1e6, (s1.pager_ring_ns - s0.pager_ring_ns) / 1e6,
           (s1.pager_unmap_ns - s0.pager_unmap_ns) / 1e6, (s1.pager_setaccess_ns - s0.pager_setaccess_ns) / 1e6, (s1.pager_issue_ns - s0.pager_issue_ns) / 1e6,
           (s1.pager_poll_ns - s0.pager_poll_ns) / 1e6, (s1.pager_lock_ns - s0.pager_lock_ns) / 1e6,
           (s1.pager_step_ns[0] - s0.pager_step_ns[0]) / 1e6, (s1.pager_step_ns[1] - s0.pager_step_ns[1]) / 1e6, (s1.pager_step_ns[2] - s0.pager_step_ns[2]) / 1e6,
           (s1.pager_step_ns[3] - s0.pager_step_ns[3]) / 1e6, (s1.pager_step_ns[4] - s0.pager_step_ns[4]) / 1e6,
           (unsigned long long)(s1.vmm_slow_calls - s0.vmm_slow_calls), (s1.vmm_slow_ns - s0.vmm_slow_ns) / 1e6, s1.vmm_max_ns / 1e6, (unsigned long long)(s1.vmm_calls - s0.vmm_calls),
           (unsigned long long)(s1.direct_in_bytes - s0.direct_in_bytes), (unsigned long long)(s1.direct_out_bytes - s0.direct_out_bytes),
           (unsigned long long)(s1.prefetch_issued - s0.prefetch_issued), (unsigned long long)(s1.prefetch_hits - s0.prefetch_hits),
           (unsigned long long)(s1.prefetch_wasted - s0.prefetch_wasted), (unsigned long long)(s1.demand_waits - s0.demand_waits),
           (unsigned long long)(s1.clean_evictions - s0.clean_evictions), (unsigned long long)s1.host_slabs, (unsigned long long)s1.host_slabs_local,
           s1.pack_span_ms - s0.pack_span_ms, s1.unpack_span_ms - s0.unpack_span_ms);
    fflush(stdout);
    return mism ? 4 : 0;
}
