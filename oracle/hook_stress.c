/*
 * hook_stress.c — TEST INFRASTRUCTURE (oracle/). Threading and fork behaviour of a preloaded hook (SURVEY.md §8b
 * "Threading": any application thread may call any hook concurrently; the hook must tolerate fork). Runs on the fake
 * driver. Modes:
 *   threads <nthreads> <iters>  : every thread allocates and frees through cuMemAlloc_v2 / cuMemAllocAsync /
 *                                 cuMemFree_v2 / cuMemFreeAsync with random sizes; at the end nothing may be left
 *                                 charged (region buffer lane == 0) or allocated in the driver.
 *   swap <nthreads> <iters>     : (functional fake, CUDA_OVERSUBSCRIBE=true) every thread owns four swappable buffers and
 *                                 fills / touches / verifies them with kernel launches while the other threads force
 *                                 evictions: a buffer must stay put from admission until its kernel has run, and every
 *                                 word must survive its page-outs.
 *   fork                        : parent allocates, forks; the child (new pid, own slot) allocates and exits WITHOUT
 *                                 freeing; the parent then fills its quota — the dead child's bytes must be reclaimed
 *                                 (rm_quitted_process) before the request is refused.
 * Prints one JSON object.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int CUresult; typedef int CUdevice; typedef unsigned long long CUdeviceptr; typedef void *CUcontext; typedef void *CUstream;
extern CUresult cuInit(unsigned);
extern CUresult cuDeviceGet(CUdevice *, int);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, CUdevice);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemAllocAsync(CUdeviceptr *, size_t, CUstream);
extern CUresult cuMemFreeAsync(CUdeviceptr, CUstream);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);
extern unsigned long long fake_gpu_used_bytes(int) __attribute__((weak));
extern CUresult cuModuleLoadData(void **, const void *);
extern CUresult cuModuleGetFunction(void **, void *, const char *);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
extern CUresult cuCtxSynchronize(void);

#define REGION_SIZE 0xC4748
#define OFF_PROCS 0x738
#define SLOT_STRIDE 0x310
static const unsigned char *g_region;
static void map_region(void) {
    const char *path = getenv("CUDA_DEVICE_MEMORY_SHARED_CACHE");
    int fd = open(path ? path : "/tmp/cudevshr.cache", O_RDONLY);
    if (fd < 0) return;
    g_region = mmap(NULL, REGION_SIZE, PROT_READ, MAP_SHARED, fd, 0);
    if (g_region == MAP_FAILED) g_region = NULL;
    close(fd);
}
static void lanes(uint64_t *ctx, uint64_t *buf, uint64_t *tot, int *nproc) {
    *ctx = *buf = *tot = 0; *nproc = 0;
    if (!g_region) return;
    int32_t procnum; memcpy(&procnum, g_region + 0xC4738, 4);      /* only the first proc_num slots are live: the reference
                                                                      compacts without clearing what it vacates */
    for (int s = 0; s < procnum && s < 1024; s++) {
        const unsigned char *slot = g_region + OFF_PROCS + (size_t)s * SLOT_STRIDE;
        int32_t pid; memcpy(&pid, slot, 4);
        if (!pid) continue;
        (*nproc)++;
        uint64_t v[5]; memcpy(v, slot + 8, 40);
        *ctx += v[0]; *buf += v[2]; *tot += v[4];
    }
}

static CUcontext g_ctx; static int g_iters; static long g_fail;
static void *worker(void *arg) {
    unsigned seed = (unsigned)(uintptr_t)arg * 2654435761u + 1;
    cuCtxSetCurrent(g_ctx);
    CUdeviceptr held[16] = {0}; int kind[16] = {0};
    for (int i = 0; i < g_iters; i++) {
        int k = rand_r(&seed) & 15;
        if (held[k]) {
            CUresult r = ((rand_r(&seed) & 1) ? cuMemFree_v2(held[k]) : cuMemFreeAsync(held[k], NULL));   /* frees cross allocator families */
            if (r) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
            held[k] = 0;
        } else {
            size_t n = 256 + (rand_r(&seed) % (1 << 20));
            kind[k] = rand_r(&seed) & 1;
            CUresult r = kind[k] ? cuMemAllocAsync(&held[k], n, NULL) : cuMemAlloc_v2(&held[k], n);
            if (r) held[k] = 0;      /* quota breach is legal under contention */
        }
    }
    for (int k = 0; k < 16; k++) if (held[k] && cuMemFree_v2(held[k])) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
    return NULL;
}

static void *g_fill, *g_touch, *g_verify;
static unsigned long long g_bad;
static void *swap_worker(void *arg) {
    unsigned id = (unsigned)(uintptr_t)arg, seed = id * 2654435761u + 7;
    cuCtxSetCurrent(g_ctx);
    enum { NB = 4 };
    CUdeviceptr buf[NB]; unsigned long long words[NB], touches[NB], idx[NB];
    for (int k = 0; k < NB; k++) {
        size_t n = ((size_t)(6 + rand_r(&seed) % 5) << 20) + (size_t)(rand_r(&seed) % 64) * 4096;
        if (cuMemAlloc_v2(&buf[k], n)) { __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED); return NULL; }
        words[k] = n / 8; touches[k] = 0; idx[k] = (unsigned long long)id * 100 + (unsigned)k;
        void *a[] = {&buf[k], &words[k], &idx[k]};
        if (cuLaunchKernel(g_fill, 64, 1, 1, 256, 1, 1, 0, NULL, a, NULL)) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
    }
    for (int i = 0; i < g_iters; i++) {
        int k = rand_r(&seed) % NB;
        if (rand_r(&seed) % 4) {
            void *a[] = {&buf[k], &words[k]};
            if (cuLaunchKernel(g_touch, 64, 1, 1, 256, 1, 1, 0, NULL, a, NULL)) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
            touches[k]++;
        } else {
            unsigned long long local = 0, *lp = &local;
            void *a[] = {&buf[k], &words[k], &idx[k], &touches[k], &lp};
            if (cuLaunchKernel(g_verify, 64, 1, 1, 256, 1, 1, 0, NULL, a, NULL)) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
            if (local) __atomic_add_fetch(&g_bad, local, __ATOMIC_RELAXED);
        }
    }
    for (int k = 0; k < NB; k++) {
        unsigned long long local = 0, *lp = &local;
        void *a[] = {&buf[k], &words[k], &idx[k], &touches[k], &lp};
        cuLaunchKernel(g_verify, 64, 1, 1, 256, 1, 1, 0, NULL, a, NULL);
        if (local) __atomic_add_fetch(&g_bad, local, __ATOMIC_RELAXED);
        if (cuMemFree_v2(buf[k])) __atomic_add_fetch(&g_fail, 1, __ATOMIC_RELAXED);
    }
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    CUdevice dev;
    if (cuInit(0) || cuDeviceGet(&dev, 0) || cuDevicePrimaryCtxRetain(&g_ctx, dev) || cuCtxSetCurrent(g_ctx)) { printf("{\"error\": \"init\"}\n"); return 1; }
    map_region();
    uint64_t c, b, t; int np;
    if (!strcmp(argv[1], "threads")) {
        int nt = argc > 2 ? atoi(argv[2]) : 8; g_iters = argc > 3 ? atoi(argv[3]) : 2000;
        pthread_t th[64];
        for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, worker, (void *)(uintptr_t)(i + 1));
        for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
        lanes(&c, &b, &t, &np);
        printf("{\"mode\": \"threads\", \"free_errors\": %ld, \"ctx\": %lu, \"buf\": %lu, \"tot\": %lu, \"procs\": %d, \"driver_bytes\": %llu}\n",
               g_fail, c, b, t, np, fake_gpu_used_bytes ? fake_gpu_used_bytes(0) : 0ull);
        return 0;
    }
    if (!strcmp(argv[1], "swap")) {
        int nt = argc > 2 ? atoi(argv[2]) : 4; g_iters = argc > 3 ? atoi(argv[3]) : 200;
        void *mod;
        if (cuModuleLoadData(&mod, "x") || cuModuleGetFunction(&g_fill, mod, "vgpu_wl_fill") || cuModuleGetFunction(&g_touch, mod, "vgpu_wl_touch") ||
            cuModuleGetFunction(&g_verify, mod, "vgpu_wl_verify")) { printf("{\"error\": \"module\"}\n"); return 1; }
        pthread_t th[64];
        for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, swap_worker, (void *)(uintptr_t)(i + 1));
        for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
        cuCtxSynchronize();
        lanes(&c, &b, &t, &np);
        printf("{\"mode\": \"swap\", \"errors\": %ld, \"bad_words\": %llu, \"buf\": %lu, \"driver_bytes\": %llu}\n", g_fail, g_bad, b,
               fake_gpu_used_bytes ? fake_gpu_used_bytes(0) : 0ull);
        return 0;
    }
    if (!strcmp(argv[1], "fork")) {
        const size_t M = 1 << 20;
        CUdeviceptr p0 = 0, p1 = 0, p2 = 0;
        CUresult r0 = cuMemAlloc_v2(&p0, 8 * M);
        fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) {
            cuCtxSetCurrent(g_ctx);
            CUdeviceptr q = 0; CUresult rc = cuMemAlloc_v2(&q, 16 * M);
            lanes(&c, &b, &t, &np);
            printf("{\"mode\": \"child\", \"rc\": %d, \"buf\": %lu, \"procs\": %d}\n", rc, b, np);
            fflush(stdout);
            _exit(0);                                  /* no atexit handlers: the slot stays behind, like a kill -9 */
        }
        int st; waitpid(pid, &st, 0);
        lanes(&c, &b, &t, &np);
        uint64_t buf_after_child = b; int procs_after_child = np;
        /* 64 MiB quota, 16 ctx, 8 own, 16 orphaned: 32 MiB only fits once the orphan is reclaimed */
        CUresult r1 = cuMemAlloc_v2(&p1, 32 * M);
        lanes(&c, &b, &t, &np);
        CUresult r2 = cuMemAlloc_v2(&p2, 16 * M);     /* now over: 16 + 8 + 32 + 16 > 64 */
        printf("{\"mode\": \"parent\", \"r0\": %d, \"buf_after_child\": %lu, \"procs_after_child\": %d, \"r1\": %d, \"buf_after_reclaim\": %lu, \"procs_after_reclaim\": %d, \"r2\": %d}\n",
               r0, buf_after_child, procs_after_child, r1, b, np, r2);
        return 0;
    }
    return 2;
}
