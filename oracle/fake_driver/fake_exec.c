/*
 * fake_exec.c — the FUNCTIONAL half of the fake driver (FAKE_GPU_EXEC=1). TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * fake_gpu.c lets the hooks' ACCOUNTING run without a GPU; this file additionally makes device memory real host memory
 * and executes work synchronously, so that the whole product — swap engine (VMM remaps, staging rings, pinned pool),
 * limiter stamps, the hook's launch path — runs end to end on a CPU box and data integrity can be checked there:
 *   - cuMemAlloc / VMM (cuMemAddressReserve, cuMemCreate, cuMemMap, cuMemUnmap, cuMemRelease) on mmap + memfd: a mapped
 *     range is readable, an unmapped one faults (like a GPU page fault kills the context), physical bytes are counted
 *     against the device total;
 *   - cuMemcpy and cuMemset families: memmove/memset at call time; streams and events complete immediately;
 *   - cuLaunchKernel of the product's own kernels, recognised BY NAME and emulated from their documented contract
 *     (k8s-device-plugin_b200/csrc/kernels.h holds the parameter layouts — included for the structs only): pack /
 *     unpack = segment copies, victim scan = the exact-LRU prefix rule, stamp = a clock read, wl_* = the synthetic
 *     workload. Unknown kernels only count as launches.
 * Everything executes in program order on the calling thread, so ordering bugs between streams are NOT exposed here;
 * logic, bookkeeping and byte movement are. The GPU tests (-m gpu) remain the parity tests for the real kernels.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include "../../k8s-device-plugin_b200/csrc/kernels.h"
#include "fake_internal.h"

#define GRAN (2ull << 20)

static uint64_t now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec; }

/* ------------------------------------------------------------------ plain allocations */
CUresult fx_alloc(CUdeviceptr *p, size_t n, int dev) {
    size_t len = (n + 4095) & ~(size_t)4095;
    void *m = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) return CUDA_ERROR_OUT_OF_MEMORY;
    CUresult r = fake_track((uint64_t)(uintptr_t)m, n, dev);
    if (r) { munmap(m, len); return r; }
    *p = (CUdeviceptr)(uintptr_t)m;
    return CUDA_SUCCESS;
}
CUresult fx_free(CUdeviceptr p) {
    uint64_t size = 0;
    CUresult r = fake_untrack(p, &size);
    if (r) return r;
    munmap((void *)(uintptr_t)p, (size + 4095) & ~(size_t)4095);
    return CUDA_SUCCESS;
}

/* ------------------------------------------------------------------ VMM */
typedef struct { int fd; size_t size; int dev; } fx_handle;

CUresult fx_address_reserve(CUdeviceptr *p, size_t size, size_t align, CUdeviceptr addr, unsigned long long flags) {
    (void)addr; (void)flags;
    if (!align) align = GRAN;
    size_t len = size + align;
    char *m = mmap(NULL, len, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) return CUDA_ERROR_OUT_OF_MEMORY;
    uintptr_t a = ((uintptr_t)m + align - 1) & ~(uintptr_t)(align - 1);
    if (a > (uintptr_t)m) munmap(m, a - (uintptr_t)m);
    uintptr_t end = (uintptr_t)m + len;
    if (end > a + size) munmap((void *)(a + size), end - (a + size));
    *p = (CUdeviceptr)a;
    return CUDA_SUCCESS;
}
CUresult fx_address_free(CUdeviceptr p, size_t n) { munmap((void *)(uintptr_t)p, n); return CUDA_SUCCESS; }
CUresult fx_mem_create(unsigned long long *h, size_t n, const void *prop, unsigned long long flags) {
    (void)flags;
    if (!h || !n || (n % GRAN)) return CUDA_ERROR_INVALID_VALUE;
    int dev = 0, host = 0;
    if (prop) { const int *pi = (const int *)prop; if (pi[2] == 1) dev = pi[3]; else if (pi[2] == 2 || pi[2] == 3) host = 1; }   /* location {type, id}: 2/3 = host memory */
    if (!host && fake_charge(dev, (int64_t)n)) return CUDA_ERROR_OUT_OF_MEMORY;
    fx_handle *x = calloc(1, sizeof *x);
    x->fd = memfd_create("fake-gpu-phys", 0);
    if (x->fd < 0 || ftruncate(x->fd, (off_t)n) != 0) { if (!host) fake_charge(dev, -(int64_t)n); if (x->fd >= 0) close(x->fd); free(x); return CUDA_ERROR_OUT_OF_MEMORY; }
    x->size = n; x->dev = host ? -1 : dev;
    *h = (unsigned long long)(uintptr_t)x;
    return CUDA_SUCCESS;
}
CUresult fx_mem_release(unsigned long long h) {
    fx_handle *x = (fx_handle *)(uintptr_t)h;
    if (!x) return CUDA_ERROR_INVALID_VALUE;
    if (x->dev >= 0) fake_charge(x->dev, -(int64_t)x->size);
    close(x->fd);           /* existing mappings keep the pages alive until unmapped, like a retained CUDA allocation */
    free(x);
    return CUDA_SUCCESS;
}
/* currently mapped VMM ranges (pointer queries on an unmapped address fail, like on the real driver) */
static struct { uint64_t va, n; } g_maps[65536];
static int g_nmaps;
static pthread_mutex_t g_maps_mu = PTHREAD_MUTEX_INITIALIZER;
int fx_is_mapped(CUdeviceptr p) {
    int hit = 0;
    pthread_mutex_lock(&g_maps_mu);
    for (int i = 0; i < g_nmaps && !hit; i++) hit = p >= g_maps[i].va && p < g_maps[i].va + g_maps[i].n;
    pthread_mutex_unlock(&g_maps_mu);
    return hit;
}
CUresult fx_mem_map(CUdeviceptr va, size_t n, size_t off, unsigned long long h, unsigned long long flags) {
    (void)flags;
    fx_handle *x = (fx_handle *)(uintptr_t)h;
    if (!x || off + n > x->size || (va % GRAN) || (n % GRAN)) return CUDA_ERROR_INVALID_VALUE;
    void *m = mmap((void *)(uintptr_t)va, n, PROT_NONE, MAP_SHARED | MAP_FIXED, x->fd, (off_t)off);   /* no access until cuMemSetAccess */
    if (m == MAP_FAILED) return CUDA_ERROR_INVALID_VALUE;
    pthread_mutex_lock(&g_maps_mu);
    if (g_nmaps < 65536) { g_maps[g_nmaps].va = va; g_maps[g_nmaps].n = n; g_nmaps++; }
    pthread_mutex_unlock(&g_maps_mu);
    return CUDA_SUCCESS;
}
CUresult fx_mem_set_access(CUdeviceptr va, size_t n, const void *desc, size_t cnt) {
    (void)desc; (void)cnt;
    return mprotect((void *)(uintptr_t)va, n, PROT_READ | PROT_WRITE) ? CUDA_ERROR_INVALID_VALUE : CUDA_SUCCESS;
}
CUresult fx_mem_unmap(CUdeviceptr va, size_t n) {
    pthread_mutex_lock(&g_maps_mu);
    for (int i = 0; i < g_nmaps; i++) if (g_maps[i].va == va) { g_maps[i] = g_maps[--g_nmaps]; break; }
    pthread_mutex_unlock(&g_maps_mu);
    void *m = mmap((void *)(uintptr_t)va, n, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_FIXED, -1, 0);
    return m == MAP_FAILED ? CUDA_ERROR_INVALID_VALUE : CUDA_SUCCESS;
}

/* ------------------------------------------------------------------ events */
typedef struct { uint64_t t; } fx_event;
CUresult fx_event_create(CUevent *e) { fx_event *x = calloc(1, sizeof *x); *e = x; return CUDA_SUCCESS; }
CUresult fx_event_record(CUevent e) { ((fx_event *)e)->t = now_ns(); return CUDA_SUCCESS; }
CUresult fx_event_elapsed(float *ms, CUevent a, CUevent b) { *ms = (float)((double)(((fx_event *)b)->t - ((fx_event *)a)->t) / 1e6); return CUDA_SUCCESS; }
CUresult fx_event_destroy(CUevent e) { free(e); return CUDA_SUCCESS; }

/* ------------------------------------------------------------------ functions by name */
enum { K_OTHER, K_PACK, K_VINIT, K_VHIST, K_VCOUNT, K_VEMIT, K_VSMALL, K_STAMP, K_FILL, K_TOUCH, K_VERIFY, K_EMPTY, K_COPY16, K_TOUCH_IND, K_VPERSIST };
typedef struct { const char *name; int kind; int nparam; size_t psize[8]; } fx_func;
static fx_func g_funcs[] = {
    {"vgpu_pack_tma", K_PACK, 1, {sizeof(VgpuPackParams)}}, {"vgpu_pack_generic", K_PACK, 1, {sizeof(VgpuPackParams)}},
    {"vgpu_victim_init", K_VINIT, 2, {8, 8}}, {"vgpu_victim_hist", K_VHIST, 6, {8, 4, 8, 4, 4, 4}},
    {"vgpu_victim_count", K_VCOUNT, 5, {8, 4, 8, 4, 4}}, {"vgpu_victim_emit", K_VEMIT, 7, {8, 4, 8, 4, 4, 8, 4}},
    {"vgpu_victim_small", K_VSMALL, 8, {8, 4, 8, 8, 4, 4, 8, 4}}, {"vgpu_victim_persist", K_VPERSIST, 8, {8, 4, 8, 8, 4, 4, 8, 4}}, {"vgpu_stamp", K_STAMP, 1, {8}},
    {"vgpu_wl_fill", K_FILL, 3, {8, 8, 8}}, {"vgpu_wl_touch", K_TOUCH, 2, {8, 8}}, {"vgpu_wl_verify", K_VERIFY, 5, {8, 8, 8, 8, 8}},
    {"vgpu_wl_empty", K_EMPTY, 0, {0}}, {"vgpu_empty", K_EMPTY, 0, {0}}, {"vgpu_copy16", K_COPY16, 3, {8, 8, 8}}, {"vgpu_wl_touch_indirect", K_TOUCH_IND, 3, {8, 4, 8}},
};
static fx_func g_other = {"?", K_OTHER, 0, {0}};
CUresult fx_get_function(CUfunction *f, const char *name) {
    for (size_t i = 0; i < sizeof g_funcs / sizeof g_funcs[0]; i++)
        if (!strcmp(g_funcs[i].name, name)) { *f = &g_funcs[i]; return CUDA_SUCCESS; }
    *f = &g_other;
    return CUDA_SUCCESS;
}
CUresult fx_param_info(CUfunction f, size_t idx, size_t *off, size_t *size) {
    fx_func *x = f;
    if (!x || (int)idx >= x->nparam) return CUDA_ERROR_INVALID_VALUE;
    size_t o = 0;
    for (size_t i = 0; i <= idx; i++) { size_t a = x->psize[i] >= 8 ? 8 : 4; o = (o + a - 1) & ~(a - 1); if (i < idx) o += x->psize[i]; }
    if (off) *off = o;
    if (size) *size = x->psize[idx];
    return CUDA_SUCCESS;
}

/* ------------------------------------------------------------------ kernel emulation */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

typedef struct { uint64_t touch; uint32_t idx; uint64_t size; } cand_t;
static int cand_cmp(const void *a, const void *b) {
    const cand_t *x = a, *y = b;
    if (x->touch != y->touch) return x->touch < y->touch ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
static int u32_cmp(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }

/* exact LRU: resident rows in (last_touch, index) order, shortest prefix whose sizes reach `need`, ascending index out */
static void victim_select(const VgpuEntry *tbl, uint32_t n, VgpuScanState *st, uint64_t need, uint32_t *out, uint32_t cap) {
    cand_t *c = malloc(sizeof(cand_t) * (n ? n : 1));
    uint32_t nc = 0;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++)
        if (tbl[i].state == VGPU_ST_RESIDENT) { c[nc].touch = tbl[i].last_touch; c[nc].idx = i; c[nc].size = tbl[i].size; total += tbl[i].size; nc++; }
    qsort(c, nc, sizeof *c, cand_cmp);
    uint32_t take = 0;
    uint64_t freed = 0;
    st->need = need; st->insufficient = 0; st->cand_bytes = 0;
    if (need == 0) take = 0;
    else if (total < need) { st->insufficient = 1; st->cand_bytes = total; take = nc; freed = total; }
    else { while (take < nc && freed < need) freed += c[take++].size; }
    uint32_t *idx = malloc(sizeof(uint32_t) * (take ? take : 1));
    for (uint32_t i = 0; i < take; i++) idx[i] = c[i].idx;
    qsort(idx, take, sizeof *idx, u32_cmp);
    for (uint32_t i = 0; i < take && i < cap; i++) out[i] = idx[i];
    st->out_count = take; st->out_freed = freed; st->done_ctas = 0;
    free(idx); free(c);
}

#define ARG(T, i) (*(T *)params[i])
CUresult fx_launch(CUfunction f, void **params) {
    fx_func *x = f;
    if (!x || !params) return CUDA_SUCCESS;
    switch (x->kind) {
    case K_PACK: {
        const VgpuPackParams *p = (const VgpuPackParams *)params[0];
        uint64_t t0 = now_ns();
        for (uint32_t s = 0; s < p->nseg; s++)
            memmove((void *)(uintptr_t)p->seg[s].dst, (const void *)(uintptr_t)p->seg[s].src, p->seg[s].bytes);
        if (p->span) { uint64_t *sp = (uint64_t *)(uintptr_t)p->span; uint64_t t1 = now_ns(); if (t0 < sp[0]) sp[0] = t0; if (t1 > sp[1]) sp[1] = t1; }
        break; }
    case K_VINIT: { VgpuScanState *st = ARG(VgpuScanState *, 0); memset(st, 0, 64); st->need = ARG(uint64_t, 1); st->need_left = st->need; break; }
    case K_VHIST: case K_VCOUNT: break;          /* the selection happens in one piece at emit time */
    case K_VEMIT: { VgpuScanState *st = ARG(VgpuScanState *, 2);
        victim_select(ARG(const VgpuEntry *, 0), ARG(uint32_t, 1), st, st->need, ARG(uint32_t *, 5), ARG(uint32_t, 6)); break; }
    case K_VPERSIST:
    case K_VSMALL: victim_select(ARG(const VgpuEntry *, 0), ARG(uint32_t, 1), ARG(VgpuScanState *, 2), ARG(uint64_t, 3), ARG(uint32_t *, 6), ARG(uint32_t, 7)); break;
    case K_STAMP: *ARG(volatile uint64_t *, 0) = now_ns(); break;
    case K_TOUCH_IND: { uint64_t **t = ARG(uint64_t **, 0); uint32_t np = ARG(uint32_t, 1); uint64_t n = ARG(uint64_t, 2);
        for (uint32_t b = 0; b < np; b++) for (uint64_t j = 0; j < n; j++) t[b][j] += 1;
        break; }
    case K_COPY16: memmove(ARG(void *, 0), ARG(const void *, 1), (size_t)ARG(uint64_t, 2) * 16); break;
    case K_FILL: { uint64_t *b = ARG(uint64_t *, 0), n = ARG(uint64_t, 1), bi = ARG(uint64_t, 2); for (uint64_t j = 0; j < n; j++) b[j] = splitmix64((bi << 32) + j); break; }
    case K_TOUCH: { uint64_t *b = ARG(uint64_t *, 0), n = ARG(uint64_t, 1); for (uint64_t j = 0; j < n; j++) b[j] += 1; break; }
    case K_VERIFY: { const uint64_t *b = ARG(const uint64_t *, 0); uint64_t n = ARG(uint64_t, 1), bi = ARG(uint64_t, 2), add = ARG(uint64_t, 3);
        unsigned long long *bad = ARG(unsigned long long *, 4); uint64_t m = 0;
        for (uint64_t j = 0; j < n; j++) if (b[j] != splitmix64((bi << 32) + j) + add) m++;
        if (m) *bad += m;
        break; }
    default: break;
    }
    return CUDA_SUCCESS;
}
