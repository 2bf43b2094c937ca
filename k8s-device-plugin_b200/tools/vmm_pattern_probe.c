/*
 * vmm_pattern_probe.c — what makes cuMemUnmap / cuMemSetAccess cost 0.3-3 ms inside the swap engine when the same
 * calls cost 0.05-0.13 ms in isolation (profiles/r02_hostvmm_probe.jsonl)? Rebuilds the engine's situation piece by
 * piece inside ONE process and times a remap cycle (unmap + map + setaccess of one 64 MiB row) under each:
 *   rows      N rows of 64 MiB mapped in one address reservation of R GiB (the arena)
 *   dma       bidirectional pinned DMA running on two side streams, in pieces of P MiB, either between plain
 *             cuMemAlloc buffers and pinned memory (like hostvmm_probe) or between the arena's OWN rows and pinned memory
 *             (like the engine's direct path), 4 pieces queued per direction
 *   pinned    G GiB of pinned host memory allocated (the engine's pool is tens of GiB)
 * Usage: vmm_pattern_probe   (no arguments; prints one JSON line per condition)
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
static double now_us(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3; }
static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }

static CUcontext g_ctx;
static CUmemAllocationProp prop;
static CUmemAccessDesc acc;
#define ROW ((size_t)64 << 20)

struct load { volatile int stop; size_t piece; int use_rows; CUdeviceptr rows; int nrows; double gbs; int depth; };
static void *load_thread(void *p) {
    struct load *L = p;
    CK(cuCtxSetCurrent(g_ctx));
    size_t span = (size_t)16 * ROW;                 /* each direction cycles over 1 GiB */
    void *h1, *h2; CUdeviceptr d1 = 0, d2 = 0; CUstream s1, s2; CUevent e1[16], e2[16];
    CK(cuMemHostAlloc(&h1, span, 0)); CK(cuMemHostAlloc(&h2, span, 0));
    if (L->use_rows) { d1 = L->rows; d2 = L->rows + span; }            /* rows [0,16) receive, rows [16,32) are drained */
    else { CK(cuMemAlloc(&d1, span)); CK(cuMemAlloc(&d2, span)); }
    CK(cuStreamCreate(&s1, CU_STREAM_NON_BLOCKING)); CK(cuStreamCreate(&s2, CU_STREAM_NON_BLOCKING));
    for (int i = 0; i < L->depth; i++) { CK(cuEventCreate(&e1[i], CU_EVENT_DISABLE_TIMING)); CK(cuEventCreate(&e2[i], CU_EVENT_DISABLE_TIMING)); }
    double t0 = now_us(); long n = 0; size_t off = 0;
    while (!L->stop) {
        int k = (int)(n % L->depth);
        if (n >= L->depth) { CK(cuEventSynchronize(e1[k])); CK(cuEventSynchronize(e2[k])); }
        CK(cuMemcpyHtoDAsync(d1 + off, (char *)h1 + off, L->piece, s1)); CK(cuEventRecord(e1[k], s1));
        CK(cuMemcpyDtoHAsync((char *)h2 + off, d2 + off, L->piece, s2)); CK(cuEventRecord(e2[k], s2));
        off += L->piece; if (off + L->piece > span) off = 0;
        n++;
    }
    CK(cuStreamSynchronize(s1)); CK(cuStreamSynchronize(s2));
    L->gbs = 2.0 * (double)L->piece * (double)n / (now_us() - t0) / 1e3;
    cuMemFreeHost(h1); cuMemFreeHost(h2);
    if (!L->use_rows) { cuMemFree(d1); cuMemFree(d2); }
    cuStreamDestroy(s1); cuStreamDestroy(s2);
    return NULL;
}

static void measure(const char *label, CUdeviceptr arena, int nrows, CUmemGenericAllocationHandle *h, int reps, struct load *L) {
    /* remap target: the LAST 8 rows (never touched by the DMA load) */
    double *un = malloc(sizeof(double) * reps), *sa = malloc(sizeof(double) * reps), *mp = malloc(sizeof(double) * reps);
    for (int r = 0; r < reps; r++) {
        int j = nrows - 1 - (r % 8);
        CUdeviceptr va = arena + (size_t)j * ROW;
        double a = now_us(); CK(cuMemUnmap(va, ROW));
        double b = now_us(); CK(cuMemMap(va, ROW, 0, h[j], 0));
        double c = now_us(); CK(cuMemSetAccess(va, ROW, &acc, 1));
        double d = now_us();
        un[r] = b - a; mp[r] = c - b; sa[r] = d - c;
        struct timespec ts = {0, 300000}; nanosleep(&ts, NULL);      /* the pager does other things between calls */
    }
    double su = 0, ss = 0, sm = 0;
    for (int r = 0; r < reps; r++) { su += un[r]; ss += sa[r]; sm += mp[r]; }
    qsort(un, reps, sizeof(double), cmp_d); qsort(sa, reps, sizeof(double), cmp_d);
    printf("{\"cond\": \"%s\", \"unmap_us\": {\"mean\": %.0f, \"p50\": %.0f, \"p95\": %.0f, \"max\": %.0f}, \"setaccess_us\": {\"mean\": %.0f, \"p50\": %.0f, \"p95\": %.0f, \"max\": %.0f}, \"map_us\": %.1f",
           label, su / reps, un[reps / 2], un[reps * 95 / 100], un[reps - 1], ss / reps, sa[reps / 2], sa[reps * 95 / 100], sa[reps - 1], sm / reps);
    if (L) printf(", \"piece_mib\": %zu, \"dma_on_rows\": %d", L->piece >> 20, L->use_rows);
    printf("}\n"); fflush(stdout);
    free(un); free(sa); free(mp);
}

static void with_load(const char *label, CUdeviceptr arena, int nrows, CUmemGenericAllocationHandle *h, size_t piece, int use_rows, int depth) {
    struct load L = {0, piece, use_rows, arena, nrows, 0, depth}; pthread_t th;
    pthread_create(&th, NULL, load_thread, &L);
    struct timespec ts = {0, 300000000}; nanosleep(&ts, NULL);
    measure(label, arena, nrows, h, 150, &L);
    L.stop = 1; pthread_join(th, NULL);
    printf("{\"cond\": \"%s\", \"bidir_gbs\": %.1f}\n", label, L.gbs); fflush(stdout);
}

int main(void) {
    CUdevice dev;
    CK(cuInit(0)); CK(cuDeviceGet(&dev, 0)); CK(cuDevicePrimaryCtxRetain(&g_ctx, dev)); CK(cuCtxSetCurrent(g_ctx));
    memset(&prop, 0, sizeof prop); prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
    memset(&acc, 0, sizeof acc); acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    size_t arena_gib[] = {16, 1024};
    for (int ai = 0; ai < 2; ai++) {
        int nrows = 128;                                               /* 8 GiB resident, like the bench quota */
        size_t span = arena_gib[ai] << 30; CUdeviceptr arena;
        CK(cuMemAddressReserve(&arena, span, (size_t)4 << 30, 0, 0));
        CUmemGenericAllocationHandle *h = malloc(sizeof *h * nrows);
        for (int i = 0; i < nrows; i++) { CK(cuMemCreate(&h[i], ROW, &prop, 0)); CK(cuMemMap(arena + (size_t)i * ROW, ROW, 0, h[i], 0)); }
        CK(cuMemSetAccess(arena, (size_t)nrows * ROW, &acc, 1));
        char label[128];
        snprintf(label, sizeof label, "arena %zu GiB, 128 rows, idle", arena_gib[ai]); measure(label, arena, nrows, h, 150, NULL);
        snprintf(label, sizeof label, "arena %zu GiB, DMA plain buffers 32 MiB pieces", arena_gib[ai]); with_load(label, arena, nrows, h, 32u << 20, 0, 4);
        snprintf(label, sizeof label, "arena %zu GiB, DMA on arena rows 32 MiB pieces", arena_gib[ai]); with_load(label, arena, nrows, h, 32u << 20, 1, 4);
        if (ai == 1) {
            with_load("arena 1024 GiB, DMA on arena rows 64 MiB pieces", arena, nrows, h, 64u << 20, 1, 4);
            with_load("arena 1024 GiB, DMA on arena rows 16 MiB pieces", arena, nrows, h, 16u << 20, 1, 4);
            with_load("arena 1024 GiB, DMA on arena rows 4 MiB pieces depth 16", arena, nrows, h, 4u << 20, 1, 16);
            with_load("arena 1024 GiB, DMA on arena rows 2 MiB pieces depth 16", arena, nrows, h, 2u << 20, 1, 16);
            with_load("arena 1024 GiB, DMA on arena rows 32 MiB pieces depth 16", arena, nrows, h, 32u << 20, 1, 16);
            /* a large pinned pool next to it, like the engine's */
            void *big[24]; int nb = 0;
            for (; nb < 24; nb++) if (cuMemHostAlloc(&big[nb], (size_t)1 << 30, CU_MEMHOSTALLOC_PORTABLE) != CUDA_SUCCESS) break;
            snprintf(label, sizeof label, "arena 1024 GiB, %d GiB pinned pool, DMA on arena rows 32 MiB pieces", nb); with_load(label, arena, nrows, h, 32u << 20, 1, 4);
            for (int i = 0; i < nb; i++) cuMemFreeHost(big[i]);
        }
        for (int i = 0; i < nrows; i++) { cuMemUnmap(arena + (size_t)i * ROW, ROW); cuMemRelease(h[i]); }
        cuMemAddressFree(arena, span); free(h);
    }
    return 0;
}
