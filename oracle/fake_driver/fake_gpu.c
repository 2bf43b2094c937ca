/*
 * fake_gpu.c — a deterministic, GPU-less stand-in for libcuda.so.1 AND libnvidia-ml.so.1.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): lets the *reference binary* lib/nvidia/libvgpu.so and the
 * new libvgpu.so run their accounting/limit logic on a CPU-only box, on the same driver-API
 * trace, so their shared-region counters and return codes can be compared word for word
 * (SURVEY.md §8c "CPU-only execution", Appendix D = the entry points the reference calls).
 * Nothing under k8s-device-plugin_b200/ links or loads this file.
 *
 * One ELF object provides both libraries; oracle/Makefile installs it as
 * oracle/_ref/fake/libcuda.so.1 plus a symlink libnvidia-ml.so.1 -> libcuda.so.1, so glibc
 * loads ONE instance (same inode) and the cu* and nvml* halves share state (NVML's process
 * list reflects the fake primary context — needed by the reference's set_task_pid,
 * libvgpu.so@0x16a7f).
 *
 * Behaviour knobs (env): FAKE_GPU_COUNT (1), FAKE_GPU_TOTAL_MIB (183359 = B200),
 * FAKE_GPU_CTX_MIB (512: bytes NVML reports for a process with a primary context),
 * FAKE_GPU_SM_UTIL (0: smUtil returned by nvmlDeviceGetProcessUtilization),
 * FAKE_GPU_LOG (unset: silent), FAKE_GPU_EXEC (unset: bookkeeping only; 1: device memory is real host memory and work
 * executes — fake_exec.c).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "fake_internal.h"
typedef struct { unsigned char bytes[16]; } CUuuid;

#define MAXDEV 16
#define EXPORT __attribute__((visibility("default")))

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_inited, g_ndev = 1, g_log, g_exec;
static uint64_t g_total = 183359ull << 20, g_ctx_bytes = 512ull << 20;
static unsigned g_sm_util;
static uint64_t g_used[MAXDEV];
static int g_ctx_refs[MAXDEV];
/* contexts: &g_ctx_obj[d] is the primary context of device d; &g_ctx_obj[MAXDEV + k] is the k-th context made by
 * cuCtxCreate_v2, whose device is g_ctx_dev[k] (distinct handles, as on the real driver) */
#define MAXCREATED 64
static int g_ctx_obj[MAXDEV + MAXCREATED];
static int g_ctx_dev[MAXCREATED];
static int g_ncreated;
static __thread CUcontext t_cur;
static uint64_t g_next_va = 0x7f4000000000ull;
static uint64_t g_launches;

/* live allocations: open-addressing table keyed by base address */
#define HCAP (1u << 20)
static struct { uint64_t base, size; int dev; int live; } *g_h;

static void fake_init(void) {
    if (g_inited) return;
    const char *e;
    if ((e = getenv("FAKE_GPU_COUNT"))) g_ndev = atoi(e);
    if (g_ndev < 1) g_ndev = 1;
    if (g_ndev > MAXDEV) g_ndev = MAXDEV;
    if ((e = getenv("FAKE_GPU_TOTAL_MIB"))) g_total = strtoull(e, 0, 0) << 20;
    if ((e = getenv("FAKE_GPU_CTX_MIB"))) g_ctx_bytes = strtoull(e, 0, 0) << 20;
    if ((e = getenv("FAKE_GPU_SM_UTIL"))) g_sm_util = (unsigned)atoi(e);
    g_log = getenv("FAKE_GPU_LOG") != NULL;
    g_exec = (e = getenv("FAKE_GPU_EXEC")) && *e && *e != '0';
    g_h = calloc(HCAP, sizeof(*g_h));
    g_inited = 1;
}
#define LOGF(...) do { if (g_log) { fprintf(stderr, "[fake_gpu] " __VA_ARGS__); fputc('\n', stderr); } } while (0)

static int cur_dev(void) {
    if (!t_cur) return -1;
    int i = (int)((int *)t_cur - g_ctx_obj);
    if (i < 0 || i >= MAXDEV + MAXCREATED) return -1;
    return i < MAXDEV ? i : g_ctx_dev[i - MAXDEV];
}

static unsigned hslot(uint64_t base) { return (unsigned)((base >> 9) * 0x9E3779B97F4A7C15ull >> 44) & (HCAP - 1); }

int fake_exec_on(void) { fake_init(); return g_exec; }

int fake_charge(int dev, int64_t delta) {
    fake_init();
    pthread_mutex_lock(&g_mu);
    if (delta > 0 && g_used[dev] + (uint64_t)delta > g_total) { pthread_mutex_unlock(&g_mu); return 1; }
    g_used[dev] = (uint64_t)((int64_t)g_used[dev] + delta);
    pthread_mutex_unlock(&g_mu);
    return 0;
}

CUresult fake_track(uint64_t base, uint64_t size, int d) {
    fake_init();
    pthread_mutex_lock(&g_mu);
    if (g_used[d] + size > g_total) { pthread_mutex_unlock(&g_mu); return CUDA_ERROR_OUT_OF_MEMORY; }
    unsigned s = hslot(base);
    while (g_h[s].live) s = (s + 1) & (HCAP - 1);
    g_h[s].base = base; g_h[s].size = size; g_h[s].dev = d; g_h[s].live = 1;
    g_used[d] += size;
    pthread_mutex_unlock(&g_mu);
    return CUDA_SUCCESS;
}

CUresult fake_untrack(uint64_t p, uint64_t *size) {
    fake_init();
    pthread_mutex_lock(&g_mu);
    unsigned s = hslot(p);
    for (unsigned n = 0; n < HCAP; n++, s = (s + 1) & (HCAP - 1)) {
        if (!g_h[s].live && g_h[s].base == 0) break;
        if (g_h[s].live && g_h[s].base == p) {
            g_h[s].live = 0; /* tombstone keeps base != 0 so probing continues */
            g_used[g_h[s].dev] -= g_h[s].size;
            if (size) *size = g_h[s].size;
            pthread_mutex_unlock(&g_mu);
            return CUDA_SUCCESS;
        }
    }
    pthread_mutex_unlock(&g_mu);
    return CUDA_ERROR_INVALID_VALUE;
}

int fake_is_tracked(uint64_t p) {
    int hit = 0;
    pthread_mutex_lock(&g_mu);
    for (unsigned s = 0; s < HCAP && !hit; s++) hit = g_h[s].live && p >= g_h[s].base && p < g_h[s].base + g_h[s].size;
    pthread_mutex_unlock(&g_mu);
    return hit;
}

static CUresult do_alloc(CUdeviceptr *dptr, size_t size) {
    fake_init();
    int d = cur_dev();
    if (d < 0) return CUDA_ERROR_INVALID_CONTEXT;
    if (!dptr || size == 0) return CUDA_ERROR_INVALID_VALUE;
    if (g_exec) return fx_alloc(dptr, size, d);
    pthread_mutex_lock(&g_mu);
    uint64_t base = g_next_va;
    g_next_va += (size + 511) & ~511ull;
    pthread_mutex_unlock(&g_mu);
    CUresult r = fake_track(base, size, d);
    if (!r) *dptr = base;
    return r;
}

/* managed allocations are remembered so that pointer queries can tell them apart, like the real driver does */
static struct { CUdeviceptr base; size_t size; } g_managed[256];
static int g_nmanaged;
static int is_managed(CUdeviceptr p) {
    int hit = 0;
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < g_nmanaged; i++) if (p >= g_managed[i].base && p < g_managed[i].base + g_managed[i].size) hit = 1;
    pthread_mutex_unlock(&g_mu);
    return hit;
}
static void forget_managed(CUdeviceptr p) {
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < g_nmanaged; i++) if (g_managed[i].base == p) { g_managed[i] = g_managed[--g_nmanaged]; break; }
    pthread_mutex_unlock(&g_mu);
}

static CUresult do_free(CUdeviceptr p) {
    fake_init();
    if (!p) return CUDA_SUCCESS;
    forget_managed(p);
    if (g_exec) return fx_free(p);
    return fake_untrack(p, NULL);
}

/* ------------------------------------------------------------------ CUDA driver half */
EXPORT CUresult cuInit(unsigned flags) { (void)flags; fake_init(); LOGF("cuInit"); return CUDA_SUCCESS; }
EXPORT CUresult cuDriverGetVersion(int *v) { *v = 12090; return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetCount(int *c) { fake_init(); *c = g_ndev; return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGet(CUdevice *d, int ord) { fake_init(); if (ord < 0 || ord >= g_ndev) return CUDA_ERROR_INVALID_DEVICE; *d = ord; return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetName(char *name, int len, CUdevice d) { (void)d; snprintf(name, (size_t)len, "NVIDIA B200 (fake)"); return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetAttribute(int *pi, int attrib, CUdevice d) {
    (void)d;
    switch (attrib) {
    case 16: *pi = 148; break;          /* MULTIPROCESSOR_COUNT */
    case 39: *pi = 2048; break;         /* MAX_THREADS_PER_MULTIPROCESSOR */
    case 75: *pi = 10; break;           /* COMPUTE_CAPABILITY_MAJOR */
    case 76: *pi = 0; break;            /* COMPUTE_CAPABILITY_MINOR */
    case 33: *pi = 0x1b + d; break;     /* PCI_BUS_ID */
    case 34: *pi = 0; break;            /* PCI_DEVICE_ID */
    case 50: *pi = 0; break;            /* PCI_DOMAIN_ID */
    case 1: *pi = 1024; break;          /* MAX_THREADS_PER_BLOCK */
    default: *pi = 0; break;
    }
    return CUDA_SUCCESS;
}
static void fill_uuid(unsigned char *b, int d) { for (int i = 0; i < 16; i++) b[i] = (unsigned char)(0xB2 ^ (i * 17) ^ d); b[15] = (unsigned char)d; }
EXPORT CUresult cuDeviceGetUuid(CUuuid *u, CUdevice d) { fill_uuid(u->bytes, d); return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetUuid_v2(CUuuid *u, CUdevice d) { fill_uuid(u->bytes, d); return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetPCIBusId(char *s, int len, CUdevice d) { snprintf(s, (size_t)len, "00000000:%02X:00.0", 0x1b + d); return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceGetByPCIBusId(CUdevice *d, const char *s) { unsigned dom, bus; if (sscanf(s, "%x:%x", &dom, &bus) == 2) { *d = (int)bus - 0x1b; return CUDA_SUCCESS; } return CUDA_ERROR_INVALID_VALUE; }
EXPORT CUresult cuDeviceTotalMem_v2(size_t *bytes, CUdevice d) { (void)d; fake_init(); *bytes = g_total; return CUDA_SUCCESS; }
EXPORT CUresult cuDeviceTotalMem(size_t *bytes, CUdevice d) { return cuDeviceTotalMem_v2(bytes, d); }
EXPORT CUresult cuDeviceComputeCapability(int *maj, int *min, CUdevice d) { (void)d; *maj = 10; *min = 0; return CUDA_SUCCESS; }

EXPORT CUresult cuDevicePrimaryCtxRetain(CUcontext *ctx, CUdevice d) {
    fake_init();
    if (d < 0 || d >= g_ndev) return CUDA_ERROR_INVALID_DEVICE;
    pthread_mutex_lock(&g_mu); g_ctx_refs[d]++; pthread_mutex_unlock(&g_mu);
    *ctx = &g_ctx_obj[d];
    LOGF("primary ctx retain dev %d", d);
    return CUDA_SUCCESS;
}
EXPORT CUresult cuDevicePrimaryCtxRelease_v2(CUdevice d) { pthread_mutex_lock(&g_mu); if (g_ctx_refs[d] > 0) g_ctx_refs[d]--; pthread_mutex_unlock(&g_mu); return CUDA_SUCCESS; }
EXPORT CUresult cuDevicePrimaryCtxRelease(CUdevice d) { return cuDevicePrimaryCtxRelease_v2(d); }
EXPORT CUresult cuDevicePrimaryCtxGetState(CUdevice d, unsigned *flags, int *active) { if (flags) *flags = 0; if (active) *active = g_ctx_refs[d] > 0; return CUDA_SUCCESS; }
EXPORT CUresult cuDevicePrimaryCtxSetFlags_v2(CUdevice d, unsigned f) { (void)d; (void)f; return CUDA_SUCCESS; }
EXPORT CUresult cuDevicePrimaryCtxReset_v2(CUdevice d) { (void)d; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxCreate_v2(CUcontext *ctx, unsigned flags, CUdevice d) {
    (void)flags;
    fake_init();
    if (d < 0 || d >= g_ndev) return CUDA_ERROR_INVALID_DEVICE;
    pthread_mutex_lock(&g_mu);
    int k = g_ncreated < MAXCREATED ? g_ncreated++ : -1;
    if (k >= 0) g_ctx_dev[k] = d;
    pthread_mutex_unlock(&g_mu);
    if (k < 0) return CUDA_ERROR_OUT_OF_MEMORY;
    *ctx = &g_ctx_obj[MAXDEV + k];
    t_cur = *ctx;
    LOGF("ctx create dev %d -> #%d", d, k);
    return CUDA_SUCCESS;
}
EXPORT CUresult cuCtxDestroy_v2(CUcontext ctx) { (void)ctx; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxSetCurrent(CUcontext ctx) { t_cur = ctx; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxGetCurrent(CUcontext *ctx) { *ctx = t_cur; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxPushCurrent_v2(CUcontext ctx) { t_cur = ctx; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxPopCurrent_v2(CUcontext *ctx) { if (ctx) *ctx = t_cur; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxGetDevice(CUdevice *d) { int c = cur_dev(); if (c < 0) return CUDA_ERROR_INVALID_CONTEXT; *d = c; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxSynchronize(void) { return CUDA_SUCCESS; }
EXPORT CUresult cuCtxGetApiVersion(CUcontext c, unsigned *v) { (void)c; *v = 12090; return CUDA_SUCCESS; }

EXPORT CUresult cuMemAlloc_v2(CUdeviceptr *p, size_t n) { return do_alloc(p, n); }
EXPORT CUresult cuMemAllocManaged(CUdeviceptr *p, size_t n, unsigned flags) {
    (void)flags;
    CUresult r = do_alloc(p, n);
    if (!r) { pthread_mutex_lock(&g_mu); if (g_nmanaged < 256) { g_managed[g_nmanaged].base = *p; g_managed[g_nmanaged].size = n; g_nmanaged++; } pthread_mutex_unlock(&g_mu); }
    return r;
}
EXPORT CUresult cuMemAllocPitch_v2(CUdeviceptr *p, size_t *pitch, size_t w, size_t h, unsigned elem) {
    (void)elem; size_t pt = (w + 511) & ~(size_t)511; if (pitch) *pitch = pt; return do_alloc(p, pt * h);
}
EXPORT CUresult cuMemFree_v2(CUdeviceptr p) { return do_free(p); }
EXPORT CUresult cuMemGetInfo_v2(size_t *fr, size_t *tot) {
    fake_init(); int d = cur_dev(); if (d < 0) return CUDA_ERROR_INVALID_CONTEXT;
    *tot = g_total; *fr = g_total - g_used[d]; return CUDA_SUCCESS;
}
EXPORT CUresult cuMemHostAlloc(void **pp, size_t n, unsigned f) { (void)f; *pp = malloc(n); return *pp ? 0 : CUDA_ERROR_OUT_OF_MEMORY; }
EXPORT CUresult cuMemAllocHost_v2(void **pp, size_t n) { *pp = malloc(n); return *pp ? 0 : CUDA_ERROR_OUT_OF_MEMORY; }
EXPORT CUresult cuMemFreeHost(void *p) { free(p); return CUDA_SUCCESS; }
EXPORT CUresult cuMemGetAddressRange_v2(CUdeviceptr *base, size_t *size, CUdeviceptr p) {
    CUresult r = CUDA_ERROR_INVALID_VALUE;
    pthread_mutex_lock(&g_mu);
    for (unsigned s = 0; s < HCAP; s++) if (g_h && g_h[s].live && p >= g_h[s].base && p < g_h[s].base + g_h[s].size) {
        if (base) *base = g_h[s].base;
        if (size) *size = g_h[s].size;
        r = CUDA_SUCCESS; break; }
    pthread_mutex_unlock(&g_mu);
    return r;
}

EXPORT CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                               unsigned smem, CUstream st, void **params, void **extra) {
    (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)smem; (void)st; (void)extra;
    __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    return fake_exec_on() ? fx_launch(f, params) : CUDA_SUCCESS;
}
EXPORT CUresult cuLaunchKernelEx(const void *cfg, CUfunction f, void **params, void **extra) {
    (void)cfg; (void)extra; __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    return fake_exec_on() ? fx_launch(f, params) : CUDA_SUCCESS;
}
EXPORT CUresult cuLaunchCooperativeKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                          unsigned smem, CUstream st, void **params) {
    return cuLaunchKernel(f, gx, gy, gz, bx, by, bz, smem, st, params, NULL);
}
static int g_mod_obj, g_fn_obj;
EXPORT CUresult cuModuleLoadData(CUmodule *m, const void *img) { (void)img; *m = &g_mod_obj; return CUDA_SUCCESS; }
EXPORT CUresult cuModuleLoad(CUmodule *m, const char *path) { (void)path; *m = &g_mod_obj; return CUDA_SUCCESS; }
EXPORT CUresult cuModuleLoadDataEx(CUmodule *m, const void *img, unsigned n, void *o, void **v) { (void)img; (void)n; (void)o; (void)v; *m = &g_mod_obj; return CUDA_SUCCESS; }
EXPORT CUresult cuModuleGetFunction(CUfunction *f, CUmodule m, const char *name) {
    (void)m; if (fake_exec_on()) return fx_get_function(f, name); *f = &g_fn_obj; return CUDA_SUCCESS;
}
EXPORT CUresult cuFuncGetParamInfo(CUfunction f, size_t idx, size_t *off, size_t *size) {
    return fake_exec_on() ? fx_param_info(f, idx, off, size) : CUDA_ERROR_INVALID_VALUE;
}
EXPORT CUresult cuFuncSetAttribute(CUfunction f, int attr, int v) { (void)f; (void)attr; (void)v; return CUDA_SUCCESS; }
EXPORT CUresult cuOccupancyMaxActiveBlocksPerMultiprocessor(int *n, CUfunction f, int bs, size_t smem) { (void)f; (void)bs; (void)smem; *n = 4; return CUDA_SUCCESS; }
EXPORT CUresult cuModuleUnload(CUmodule m) { (void)m; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamCreate(CUstream *s, unsigned f) { (void)f; *s = &g_mod_obj; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamCreateWithPriority(CUstream *s, unsigned f, int p) { (void)f; (void)p; *s = &g_mod_obj; return CUDA_SUCCESS; }
EXPORT CUresult cuCtxGetStreamPriorityRange(int *lo, int *hi) { if (lo) *lo = 0; if (hi) *hi = -5; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamDestroy_v2(CUstream s) { (void)s; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamSynchronize(CUstream s) { (void)s; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamQuery(CUstream s) { (void)s; return CUDA_SUCCESS; }
EXPORT CUresult cuStreamWaitEvent(CUstream s, CUevent e, unsigned f) { (void)s; (void)e; (void)f; return CUDA_SUCCESS; }
static int g_capturing;                                          /* tests switch "a capture is active" on and off; the work still executes */
EXPORT void fake_set_capturing(int on) { g_capturing = on; }
EXPORT CUresult cuStreamIsCapturing(CUstream s, int *status) { (void)s; if (status) *status = g_capturing ? 1 : 0; return CUDA_SUCCESS; }
/* work completes at call time: an event is the clock at record time */
EXPORT CUresult cuEventCreate(CUevent *e, unsigned f) { (void)f; return fx_event_create(e); }
EXPORT CUresult cuEventRecord(CUevent e, CUstream s) { (void)s; return fx_event_record(e); }
EXPORT CUresult cuEventSynchronize(CUevent e) { (void)e; return CUDA_SUCCESS; }
EXPORT CUresult cuEventQuery(CUevent e) { (void)e; return CUDA_SUCCESS; }
EXPORT CUresult cuEventElapsedTime(float *ms, CUevent a, CUevent b) { return fx_event_elapsed(ms, a, b); }
EXPORT CUresult cuEventDestroy_v2(CUevent e) { return fx_event_destroy(e); }
/* copies and fills act on real memory in exec mode; in bookkeeping mode device addresses are not backed */
#define FX_ONLY(stmt) do { if (fake_exec_on()) { stmt; } return CUDA_SUCCESS; } while (0)
EXPORT CUresult cuMemcpyHtoD_v2(CUdeviceptr d, const void *s, size_t n) { FX_ONLY(memmove((void *)(uintptr_t)d, s, n)); }
EXPORT CUresult cuMemcpyDtoH_v2(void *d, CUdeviceptr s, size_t n) { FX_ONLY(memmove(d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpyDtoD_v2(CUdeviceptr d, CUdeviceptr s, size_t n) { FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr d, const void *s, size_t n, CUstream st) { (void)st; FX_ONLY(memmove((void *)(uintptr_t)d, s, n)); }
EXPORT CUresult cuMemcpyDtoHAsync_v2(void *d, CUdeviceptr s, size_t n, CUstream st) { (void)st; FX_ONLY(memmove(d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpyDtoDAsync_v2(CUdeviceptr d, CUdeviceptr s, size_t n, CUstream st) { (void)st; FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpy(CUdeviceptr d, CUdeviceptr s, size_t n) { FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpyAsync(CUdeviceptr d, CUdeviceptr s, size_t n, CUstream st) { (void)st; FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
/* batched copies (CUDA 12.8): attributes are hints, the copies are plain moves; a counter tells tests which form arrived */
static int g_batch_calls;
EXPORT int fake_batch_calls(void) { return g_batch_calls; }
EXPORT CUresult cuMemcpyBatchAsync(CUdeviceptr *d, CUdeviceptr *s, size_t *n, size_t count, void *attrs, size_t *idx, size_t nattrs, size_t *fail, CUstream st) {
    (void)attrs; (void)idx; (void)nattrs; (void)fail; (void)st;
    __atomic_add_fetch(&g_batch_calls, 1, __ATOMIC_RELAXED);
    if (fake_exec_on()) for (size_t i = 0; i < count; i++) memmove((void *)(uintptr_t)d[i], (const void *)(uintptr_t)s[i], n[i]);
    return CUDA_SUCCESS;
}
EXPORT CUresult cuMemcpyPeer(CUdeviceptr d, CUcontext dc, CUdeviceptr s, CUcontext sc, size_t n) { (void)dc; (void)sc; FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemcpyPeerAsync(CUdeviceptr d, CUcontext dc, CUdeviceptr s, CUcontext sc, size_t n, CUstream st) { (void)dc; (void)sc; (void)st; FX_ONLY(memmove((void *)(uintptr_t)d, (const void *)(uintptr_t)s, n)); }
EXPORT CUresult cuMemsetD8_v2(CUdeviceptr d, unsigned char v, size_t n) { FX_ONLY(memset((void *)(uintptr_t)d, v, n)); }
EXPORT CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream st) { (void)st; FX_ONLY(memset((void *)(uintptr_t)d, v, n)); }
EXPORT CUresult cuMemsetD16_v2(CUdeviceptr d, unsigned short v, size_t n) { FX_ONLY(for (size_t i = 0; i < n; i++) ((unsigned short *)(uintptr_t)d)[i] = v); }
EXPORT CUresult cuMemsetD16Async(CUdeviceptr d, unsigned short v, size_t n, CUstream st) { (void)st; return cuMemsetD16_v2(d, v, n); }
EXPORT CUresult cuMemsetD32_v2(CUdeviceptr d, unsigned v, size_t n) { FX_ONLY(for (size_t i = 0; i < n; i++) ((unsigned *)(uintptr_t)d)[i] = v); }
EXPORT CUresult cuMemsetD32Async(CUdeviceptr d, unsigned v, size_t n, CUstream st) { (void)st; return cuMemsetD32_v2(d, v, n); }
EXPORT CUresult cuMemHostGetDevicePointer_v2(CUdeviceptr *d, void *h, unsigned f) { (void)f; *d = (CUdeviceptr)(uintptr_t)h; return CUDA_SUCCESS; }
static int g_host_unregisters;
EXPORT CUresult cuMemHostRegister_v2(void *p, size_t n, unsigned f) { (void)p; (void)n; (void)f; return CUDA_SUCCESS; }
EXPORT CUresult cuMemHostUnregister(void *p) { (void)p; __atomic_add_fetch(&g_host_unregisters, 1, __ATOMIC_RELAXED); return CUDA_SUCCESS; }
static int g_mipmaps;
EXPORT CUresult cuMipmappedArrayCreate(void **h, const void *desc, unsigned levels) { (void)desc; (void)levels; *h = malloc(16); __atomic_add_fetch(&g_mipmaps, 1, __ATOMIC_RELAXED); return CUDA_SUCCESS; }
EXPORT CUresult cuMipmappedArrayDestroy(void *h) { free(h); __atomic_sub_fetch(&g_mipmaps, 1, __ATOMIC_RELAXED); return CUDA_SUCCESS; }
/* test introspection: what is still registered / alive on the "driver" side */
EXPORT int fake_live_mipmaps(void) { return g_mipmaps; }
EXPORT int fake_host_unregisters(void) { return g_host_unregisters; }
/* pointer queries: device memory when the address lies in a live allocation or a mapped VMM range, INVALID_VALUE otherwise */
static CUresult pointer_attr(int attr, void *data, CUdeviceptr p) {
    fake_init();
    if (!(fake_is_tracked(p) || (fake_exec_on() && fx_is_mapped(p)))) return CUDA_ERROR_INVALID_VALUE;
    if (!data) return CUDA_SUCCESS;
    switch (attr) {
    case 2: *(unsigned *)data = is_managed(p) ? 3 : 2; break; /* MEMORY_TYPE = DEVICE, UNIFIED for managed memory */
    case 3: *(CUdeviceptr *)data = p; break;                 /* DEVICE_POINTER */
    case 8: *(unsigned *)data = (unsigned)is_managed(p); break; /* IS_MANAGED */
    case 9: *(int *)data = cur_dev() < 0 ? 0 : cur_dev(); break;   /* DEVICE_ORDINAL */
    default: memset(data, 0, 4); break;
    }
    return CUDA_SUCCESS;
}
EXPORT CUresult cuPointerGetAttribute(void *data, int attr, CUdeviceptr p) { return pointer_attr(attr, data, p); }
EXPORT CUresult cuPointerGetAttributes(unsigned n, int *attrs, void **data, CUdeviceptr p) {
    for (unsigned i = 0; i < n; i++) { CUresult r = pointer_attr(attrs[i], data ? data[i] : NULL, p); if (r) return r; }
    return CUDA_SUCCESS;
}
EXPORT CUresult cuMemGetAllocationGranularity(size_t *g, const void *prop, int opt) { (void)prop; (void)opt; *g = 2u << 20; return CUDA_SUCCESS; }
EXPORT CUresult cuMemAddressReserve(CUdeviceptr *p, size_t n, size_t align, CUdeviceptr addr, unsigned long long fl) {
    if (fake_exec_on()) return fx_address_reserve(p, n, align, addr, fl);
    (void)align; (void)addr; (void)fl; *p = 0x600000000000ull; return CUDA_SUCCESS;
}
EXPORT CUresult cuMemMap(CUdeviceptr va, size_t n, size_t off, unsigned long long h, unsigned long long fl) {
    return fake_exec_on() ? fx_mem_map(va, n, off, h, fl) : CUDA_SUCCESS;
}
EXPORT CUresult cuGetErrorString(CUresult e, const char **s) { (void)e; *s = "fake"; return CUDA_SUCCESS; }
EXPORT CUresult cuGetErrorName(CUresult e, const char **s) { (void)e; *s = "FAKE"; return CUDA_SUCCESS; }
/* The reference patches slots 2 and 6 of the cudart-interface export table in place (libvgpu.so@0x3f858-0x3f8fd)
 * and copies slots 0..2 of a second one (@0x3f8ff): hand back writable tables of inert entries. */
static CUresult export_stub(void) { return CUDA_ERROR_NOT_SUPPORTED; }
static void *g_export_tbl[4][16] __attribute__((aligned(4096)));
EXPORT CUresult cuGetExportTable(const void **tbl, const CUuuid *id) {
    if (!tbl || !id) return CUDA_ERROR_INVALID_VALUE;
    unsigned k = id->bytes[0] & 3u;
    g_export_tbl[k][0] = (void *)(uintptr_t)(16 * sizeof(void *));
    for (int i = 1; i < 16; i++) if (!g_export_tbl[k][i]) g_export_tbl[k][i] = (void *)export_stub;
    *tbl = g_export_tbl[k];
    return CUDA_SUCCESS;
}
EXPORT CUresult cuArray3DGetDescriptor_v2(void *d, void *a) { (void)d; (void)a; return CUDA_ERROR_NOT_SUPPORTED; }
EXPORT CUresult cuMemAddressFree(CUdeviceptr p, size_t n) { if (fake_exec_on()) return fx_address_free(p, n); return CUDA_SUCCESS; }
/* VMM physical handles and stream-ordered allocations: same bookkeeping as cuMemAlloc (a handle is its fake address) */
EXPORT CUresult cuMemCreate(unsigned long long *h, size_t n, const void *prop, unsigned long long flags) {
    if (fake_exec_on()) return fx_mem_create(h, n, prop, flags);
    (void)prop; (void)flags; CUdeviceptr p = 0; CUresult r = do_alloc(&p, n); if (!r) *h = p; return r;
}
EXPORT CUresult cuMemRelease(unsigned long long h) { if (fake_exec_on()) return fx_mem_release(h); return do_free(h); }
EXPORT CUresult cuMemAllocAsync(CUdeviceptr *p, size_t n, CUstream st) { (void)st; return do_alloc(p, n); }
EXPORT CUresult cuMemAllocFromPoolAsync(CUdeviceptr *p, size_t n, void *pool, CUstream st) { (void)pool; (void)st; return do_alloc(p, n); }
EXPORT CUresult cuMemFreeAsync(CUdeviceptr p, CUstream st) { (void)st; return do_free(p); }
/* explicitly built graphs, just enough to replay kernel and 1-D memset nodes in insertion order (exec mode): a replay touches
 * its operands with no further call into the hook, which is the property the swap-mode tests need */
#define FG_MAGIC 0x6772617068ull
struct fg_node { int kind; CUfunction f; void *pv[16]; unsigned long long val[16][4]; int np; CUdeviceptr dst; unsigned value, esize; size_t width; };
struct fg_graph { unsigned long long magic; int n; struct fg_node node[64]; };
struct fg_kparams { CUfunction func; unsigned gx, gy, gz, bx, by, bz, smem; void **kernelParams; void **extra; void *kern; void *ctx; };
struct fg_msparams { CUdeviceptr dst; size_t pitch; unsigned value, elementSize; size_t width, height; };
static void *g_fg_live[128];                                     /* graphs and instances made here (cuGraphLaunch is also called with opaque handles) */
static void fg_reg(void *p, int on) { pthread_mutex_lock(&g_mu); for (int i = 0; i < 128; i++) if (g_fg_live[i] == (on ? NULL : p)) { g_fg_live[i] = on ? p : NULL; break; } pthread_mutex_unlock(&g_mu); }
static int fake_is_heap_graph(void *p) { int hit = 0; pthread_mutex_lock(&g_mu); for (int i = 0; i < 128; i++) hit |= g_fg_live[i] == p; pthread_mutex_unlock(&g_mu); return hit; }
EXPORT CUresult cuGraphCreate(void **g, unsigned flags) { (void)flags; struct fg_graph *x = calloc(1, sizeof *x); x->magic = FG_MAGIC; *g = x; fg_reg(x, 1); return CUDA_SUCCESS; }
EXPORT CUresult cuGraphDestroy(void *g) { struct fg_graph *x = g; if (x && fake_is_heap_graph(x)) { fg_reg(x, 0); x->magic = 0; free(x); } return CUDA_SUCCESS; }
EXPORT CUresult cuGraphAddKernelNode_v2(void **node, void *g, const void *deps, size_t ndeps, const struct fg_kparams *p) {
    (void)deps; (void)ndeps;
    struct fg_graph *x = g;
    if (!x || x->magic != FG_MAGIC || !p || x->n >= 64) return CUDA_ERROR_INVALID_VALUE;
    struct fg_node *nd = &x->node[x->n];
    memset(nd, 0, sizeof *nd);
    nd->kind = 0; nd->f = p->func;
    for (size_t i = 0; i < 16 && p->kernelParams; i++) {          /* the parameter VALUES are captured when the node is defined */
        size_t off, size;
        if (!fake_exec_on() || fx_param_info(p->func, i, &off, &size) != CUDA_SUCCESS) break;
        if (size > sizeof nd->val[i]) return CUDA_ERROR_INVALID_VALUE;
        memcpy(nd->val[i], p->kernelParams[i], size);
        nd->pv[i] = nd->val[i];
        nd->np = (int)i + 1;
    }
    if (node) *node = nd;
    x->n++;
    return CUDA_SUCCESS;
}
EXPORT CUresult cuGraphAddMemsetNode(void **node, void *g, const void *deps, size_t ndeps, const struct fg_msparams *m, void *ctx) {
    (void)deps; (void)ndeps; (void)ctx;
    struct fg_graph *x = g;
    if (!x || x->magic != FG_MAGIC || !m || x->n >= 64 || m->height > 1) return CUDA_ERROR_INVALID_VALUE;
    struct fg_node *nd = &x->node[x->n];
    memset(nd, 0, sizeof *nd);
    nd->kind = 1; nd->dst = m->dst; nd->value = m->value; nd->esize = m->elementSize; nd->width = m->width;
    if (node) *node = nd;
    x->n++;
    return CUDA_SUCCESS;
}
/* CUDA_MEMCPY3D as cuda.h lays it out (200 bytes); only linear device/host operands with Height = Depth = 1 are replayed */
struct fg_copy3d {
    size_t srcXInBytes, srcY, srcZ, srcLOD; unsigned srcMemoryType, pad0; const void *srcHost; CUdeviceptr srcDevice; void *srcArray; void *reserved0; size_t srcPitch, srcHeight;
    size_t dstXInBytes, dstY, dstZ, dstLOD; unsigned dstMemoryType, pad1; void *dstHost; CUdeviceptr dstDevice; void *dstArray; void *reserved1; size_t dstPitch, dstHeight;
    size_t WidthInBytes, Height, Depth;
};
static CUresult fg_copy_now(const struct fg_copy3d *c) {
    if (!c || c->Height > 1 || c->Depth > 1 || c->srcArray || c->dstArray) return CUDA_ERROR_INVALID_VALUE;
    const void *src = c->srcMemoryType == 1 ? c->srcHost : (const void *)(uintptr_t)c->srcDevice;      /* CU_MEMORYTYPE_HOST = 1, DEVICE = 2, UNIFIED = 4 */
    void *dst = c->dstMemoryType == 1 ? c->dstHost : (void *)(uintptr_t)c->dstDevice;
    if (fake_exec_on()) memmove((char *)dst + c->dstXInBytes, (const char *)src + c->srcXInBytes, c->WidthInBytes);
    return CUDA_SUCCESS;
}
EXPORT CUresult cuMemcpy3D_v2(const struct fg_copy3d *c) { return fg_copy_now(c); }
EXPORT CUresult cuMemcpy3DAsync_v2(const struct fg_copy3d *c, CUstream st) { (void)st; return fg_copy_now(c); }
static struct fg_copy3d g_fg_copies[64];
static int g_fg_ncopies;
EXPORT CUresult cuGraphAddMemcpyNode(void **node, void *g, const void *deps, size_t ndeps, const struct fg_copy3d *c, void *ctx) {
    (void)deps; (void)ndeps; (void)ctx;
    struct fg_graph *x = g;
    if (!x || x->magic != FG_MAGIC || !c || x->n >= 64 || g_fg_ncopies >= 64) return CUDA_ERROR_INVALID_VALUE;
    struct fg_node *nd = &x->node[x->n];
    memset(nd, 0, sizeof *nd);
    nd->kind = 2; nd->np = g_fg_ncopies;                     /* index into the copy-parameter store */
    g_fg_copies[g_fg_ncopies++] = *c;
    if (node) *node = nd;
    x->n++;
    return CUDA_SUCCESS;
}
/* CUgraphNodeParams: { int type; int reserved0[3]; union { ... } at offset 16; long long reserved2 }. Types: KERNEL 0, MEMCPY 1, MEMSET 2 */
struct fg_generic { int type, r0[3]; union { long long pad[29]; struct fg_kparams kernel; struct { int flags, reserved; void *copyCtx; struct fg_copy3d copyParams; } memcpy;
                                             struct { CUdeviceptr dst; size_t pitch; unsigned value, elementSize; size_t width, height; void *ctx; } memset; } u; long long r2; };
EXPORT CUresult cuGraphAddKernelNode_v2(void **node, void *g, const void *deps, size_t ndeps, const struct fg_kparams *p);
EXPORT CUresult cuGraphAddMemsetNode(void **node, void *g, const void *deps, size_t ndeps, const struct fg_msparams *m, void *ctx);
EXPORT CUresult cuGraphAddNode(void **node, void *g, const void *deps, size_t ndeps, struct fg_generic *np) {
    if (!np) return CUDA_ERROR_INVALID_VALUE;
    if (np->type == 0) return cuGraphAddKernelNode_v2(node, g, deps, ndeps, &np->u.kernel);
    if (np->type == 1) return cuGraphAddMemcpyNode(node, g, deps, ndeps, &np->u.memcpy.copyParams, np->u.memcpy.copyCtx);
    if (np->type == 2) { struct fg_msparams m = {np->u.memset.dst, np->u.memset.pitch, np->u.memset.value, np->u.memset.elementSize, np->u.memset.width, np->u.memset.height};
                         return cuGraphAddMemsetNode(node, g, deps, ndeps, &m, np->u.memset.ctx); }
    return CUDA_ERROR_INVALID_VALUE;
}
EXPORT CUresult cuGraphInstantiateWithFlags(void **ge, void *g, unsigned long long flags) {
    (void)flags;
    struct fg_graph *x = g;
    if (!x || x->magic != FG_MAGIC) return CUDA_ERROR_INVALID_VALUE;
    struct fg_graph *c = malloc(sizeof *c);
    memcpy(c, x, sizeof *c);
    for (int i = 0; i < c->n; i++) if (c->node[i].kind == 0) for (int k = 0; k < c->node[i].np; k++) c->node[i].pv[k] = c->node[i].val[k];
    *ge = c;
    fg_reg(c, 1);
    return CUDA_SUCCESS;
}
EXPORT CUresult cuGraphExecDestroy(void *ge) { return cuGraphDestroy(ge); }
EXPORT CUresult cuGraphLaunch(void *g, CUstream st) {
    (void)st; __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED);
    struct fg_graph *x = g;
    if (!fake_exec_on() || !x || !fake_is_heap_graph(x)) return CUDA_SUCCESS;    /* tests of the limiter pass opaque handles */
    for (int i = 0; i < x->n; i++) {
        struct fg_node *nd = &x->node[i];
        if (nd->kind == 0) { CUresult r = fx_launch(nd->f, nd->pv); if (r != CUDA_SUCCESS) return r; }
        else if (nd->kind == 2) { CUresult r = fg_copy_now(&g_fg_copies[nd->np]); if (r != CUDA_SUCCESS) return r; }
        else if (nd->esize == 1) memset((void *)(uintptr_t)nd->dst, (int)nd->value, nd->width);
        else if (nd->esize == 2) for (size_t k = 0; k < nd->width; k++) ((unsigned short *)(uintptr_t)nd->dst)[k] = (unsigned short)nd->value;
        else for (size_t k = 0; k < nd->width; k++) ((unsigned *)(uintptr_t)nd->dst)[k] = nd->value;
    }
    return CUDA_SUCCESS;
}
EXPORT CUresult cuMemSetAccess(CUdeviceptr p, size_t n, const void *d, size_t c) { if (fake_exec_on()) return fx_mem_set_access(p, n, d, c); return CUDA_SUCCESS; }
EXPORT CUresult cuMemUnmap(CUdeviceptr p, size_t n) { if (fake_exec_on()) return fx_mem_unmap(p, n); return CUDA_SUCCESS; }

/* test hook: number of kernel launches that reached the "hardware" */
EXPORT uint64_t fake_gpu_launch_count(void) { return g_launches; }
EXPORT uint64_t fake_gpu_used_bytes(int d) { return g_used[d]; }

static void *self_sym(const char *name) {
    static void *self;
    if (!self) { Dl_info di; if (dladdr((void *)&cuInit, &di)) self = dlopen(di.dli_fname, RTLD_NOW | RTLD_NOLOAD); }
    return self ? dlsym(self, name) : NULL;
}
static CUresult proc_addr(const char *sym, void **pfn) {
    char buf[160];
    void *p = NULL;
    snprintf(buf, sizeof buf, "%s_v3", sym); p = self_sym(buf);
    if (!p) { snprintf(buf, sizeof buf, "%s_v2", sym); p = self_sym(buf); }
    if (!p) p = self_sym(sym);
    *pfn = p;
    return p ? CUDA_SUCCESS : CUDA_ERROR_NOT_FOUND;
}
EXPORT CUresult cuGetProcAddress(const char *sym, void **pfn, int ver, uint64_t flags) { (void)ver; (void)flags; return proc_addr(sym, pfn); }
EXPORT CUresult cuGetProcAddress_v2(const char *sym, void **pfn, int ver, uint64_t flags, int *status) {
    (void)ver; (void)flags; CUresult r = proc_addr(sym, pfn); if (status) *status = r ? 1 : 0; return CUDA_SUCCESS;
}

/* ------------------------------------------------------------------ NVML half */
typedef int nvmlReturn_t;
typedef void *nvmlDevice_t;
#define NVML_SUCCESS 0
#define NVML_ERROR_INVALID_ARGUMENT 2
#define NVML_ERROR_INSUFFICIENT_SIZE 7
#define NVML_ERROR_NOT_FOUND 6
static int g_nvdev_obj[MAXDEV];
static int nv_index(nvmlDevice_t h) { return (int)((int *)h - g_nvdev_obj); }

typedef struct { unsigned pid; unsigned long long usedGpuMemory; unsigned gi, ci; } fake_procinfo_v2;
typedef struct { unsigned pid; unsigned long long usedGpuMemory; } fake_procinfo_v1;
typedef struct { unsigned pid; unsigned long long timeStamp; unsigned smUtil, memUtil, encUtil, decUtil; } fake_procutil;
typedef struct { unsigned long long total, free, used; } fake_meminfo;
typedef struct { unsigned version; unsigned long long total, reserved, free, used; } fake_meminfo_v2;
typedef struct { char busIdLegacy[16]; unsigned domain, bus, device, pciDeviceId, pciSubSystemId; char busId[32]; } fake_pciinfo;

/* driver 580's libnvidia-ml resolves symbols with dlsym() while it initialises; FAKE_NVML_DLSYM=1 mimics that, which is
 * what deadlocks the reference hook there (its dlsym override re-enters pthread_once(preInit) from inside preInit) */
static void nvml_init_side_effects(void) { if (getenv("FAKE_NVML_DLSYM")) (void)dlsym(RTLD_DEFAULT, "cuGetProcAddress"); }
EXPORT nvmlReturn_t nvmlInit_v2(void) { fake_init(); nvml_init_side_effects(); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlInit(void) { fake_init(); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlInitWithFlags(unsigned f) { (void)f; fake_init(); nvml_init_side_effects(); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlShutdown(void) { return NVML_SUCCESS; }
EXPORT const char *nvmlErrorString(nvmlReturn_t r) { (void)r; return "fake nvml"; }
EXPORT nvmlReturn_t nvmlDeviceGetCount_v2(unsigned *c) { fake_init(); *c = (unsigned)g_ndev; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetCount(unsigned *c) { return nvmlDeviceGetCount_v2(c); }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned i, nvmlDevice_t *h) { fake_init(); if ((int)i >= g_ndev) return NVML_ERROR_INVALID_ARGUMENT; *h = &g_nvdev_obj[i]; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByIndex(unsigned i, nvmlDevice_t *h) { return nvmlDeviceGetHandleByIndex_v2(i, h); }
EXPORT nvmlReturn_t nvmlDeviceGetIndex(nvmlDevice_t h, unsigned *i) { *i = (unsigned)nv_index(h); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetUUID(nvmlDevice_t h, char *uuid, unsigned len) {
    unsigned char b[16]; fill_uuid(b, nv_index(h));
    snprintf(uuid, len, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x",
             b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
    return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetHandleByUUID(const char *uuid, nvmlDevice_t *h) {
    char buf[96];
    for (int i = 0; i < g_ndev; i++) { nvmlDeviceGetUUID(&g_nvdev_obj[i], buf, sizeof buf); if (!strcmp(buf, uuid)) { *h = &g_nvdev_obj[i]; return NVML_SUCCESS; } }
    return NVML_ERROR_NOT_FOUND;
}
EXPORT nvmlReturn_t nvmlDeviceGetName(nvmlDevice_t h, char *name, unsigned len) { (void)h; snprintf(name, len, "NVIDIA B200 (fake)"); return NVML_SUCCESS; }
static void fill_pci(fake_pciinfo *p, int d) {
    memset(p, 0, sizeof *p); p->bus = 0x1b + (unsigned)d;
    snprintf(p->busIdLegacy, sizeof p->busIdLegacy, "0000:%02X:00.0", p->bus);
    snprintf(p->busId, sizeof p->busId, "00000000:%02X:00.0", p->bus);
}
EXPORT nvmlReturn_t nvmlDeviceGetPciInfo_v3(nvmlDevice_t h, fake_pciinfo *p) { fill_pci(p, nv_index(h)); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetPciInfo_v2(nvmlDevice_t h, fake_pciinfo *p) { fill_pci(p, nv_index(h)); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetPciInfo(nvmlDevice_t h, fake_pciinfo *p) { fill_pci(p, nv_index(h)); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByPciBusId_v2(const char *s, nvmlDevice_t *h) { unsigned dom, bus; if (sscanf(s, "%x:%x", &dom, &bus) == 2 && (int)bus - 0x1b < g_ndev) { *h = &g_nvdev_obj[bus - 0x1b]; return NVML_SUCCESS; } return NVML_ERROR_NOT_FOUND; }

static unsigned long long proc_bytes(int d) { return g_ctx_refs[d] > 0 ? g_ctx_bytes + g_used[d] : 0; }
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses_v3(nvmlDevice_t h, unsigned *cnt, fake_procinfo_v2 *infos) {
    int d = nv_index(h); unsigned have = g_ctx_refs[d] > 0 ? 1u : 0u;
    if (!cnt) return NVML_ERROR_INVALID_ARGUMENT;
    if (*cnt < have) { *cnt = have; return NVML_ERROR_INSUFFICIENT_SIZE; }
    if (have && infos) { infos[0].pid = (unsigned)getpid(); infos[0].usedGpuMemory = proc_bytes(d); infos[0].gi = infos[0].ci = 0xFFFFFFFFu; }
    *cnt = have; return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses_v2(nvmlDevice_t h, unsigned *cnt, fake_procinfo_v2 *infos) { return nvmlDeviceGetComputeRunningProcesses_v3(h, cnt, infos); }
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses(nvmlDevice_t h, unsigned *cnt, fake_procinfo_v1 *infos) {
    int d = nv_index(h); unsigned have = g_ctx_refs[d] > 0 ? 1u : 0u;
    if (!cnt) return NVML_ERROR_INVALID_ARGUMENT;
    if (*cnt < have) { *cnt = have; return NVML_ERROR_INSUFFICIENT_SIZE; }
    if (have && infos) { infos[0].pid = (unsigned)getpid(); infos[0].usedGpuMemory = proc_bytes(d); }
    *cnt = have; return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetGraphicsRunningProcesses_v3(nvmlDevice_t h, unsigned *cnt, void *infos) { (void)h; (void)infos; *cnt = 0; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetProcessUtilization(nvmlDevice_t h, fake_procutil *u, unsigned *cnt, unsigned long long since) {
    (void)since; int d = nv_index(h); unsigned have = g_ctx_refs[d] > 0 ? 1u : 0u;
    if (!cnt) return NVML_ERROR_INVALID_ARGUMENT;
    if (!u || *cnt < have) { *cnt = have; return NVML_ERROR_INSUFFICIENT_SIZE; }
    if (have) { struct timeval tv; gettimeofday(&tv, 0); memset(u, 0, sizeof *u); u[0].pid = (unsigned)getpid();
        u[0].timeStamp = (unsigned long long)tv.tv_sec * 1000000ull + (unsigned long long)tv.tv_usec; u[0].smUtil = g_sm_util; }
    *cnt = have; return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t h, fake_meminfo *m) { int d = nv_index(h); m->total = g_total; m->used = proc_bytes(d); m->free = g_total - m->used; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo_v2(nvmlDevice_t h, fake_meminfo_v2 *m) { int d = nv_index(h); m->total = g_total; m->reserved = 0; m->used = proc_bytes(d); m->free = g_total - m->used; return NVML_SUCCESS; }
/* health events (rm/health.go): FAKE_NVML_XID="<device index>:<xid>@<ms after the set was created>" injects one critical
 * Xid event; otherwise every wait times out */
typedef struct { nvmlDevice_t device; unsigned long long eventType, eventData; unsigned gpuInstanceId, computeInstanceId; } fake_eventdata;
static struct { int dev; unsigned long long xid; long at_ms; int fired; struct timeval t0; int armed; } g_xid;
EXPORT nvmlReturn_t nvmlEventSetCreate(void **set) {
    static int obj; *set = &obj;
    const char *e = getenv("FAKE_NVML_XID");
    memset(&g_xid, 0, sizeof g_xid);
    if (e && sscanf(e, "%d:%llu@%ld", &g_xid.dev, &g_xid.xid, &g_xid.at_ms) == 3) { g_xid.armed = 1; gettimeofday(&g_xid.t0, 0); }
    return NVML_SUCCESS;
}
EXPORT nvmlReturn_t nvmlEventSetFree(void *set) { (void)set; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceGetSupportedEventTypes(nvmlDevice_t h, unsigned long long *types) { (void)h; *types = 0x1f; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlDeviceRegisterEvents(nvmlDevice_t h, unsigned long long types, void *set) { (void)h; (void)types; (void)set; return NVML_SUCCESS; }
static nvmlReturn_t event_wait(fake_eventdata *d, unsigned timeout_ms) {
    struct timeval now; gettimeofday(&now, 0);
    long el = (now.tv_sec - g_xid.t0.tv_sec) * 1000 + (now.tv_usec - g_xid.t0.tv_usec) / 1000;
    if (g_xid.armed && !g_xid.fired) {
        long wait = g_xid.at_ms - el;
        if (wait <= (long)timeout_ms) {
            if (wait > 0) usleep((useconds_t)wait * 1000);
            g_xid.fired = 1;
            memset(d, 0, sizeof *d);
            d->device = &g_nvdev_obj[g_xid.dev]; d->eventType = 0x8; d->eventData = g_xid.xid; d->gpuInstanceId = d->computeInstanceId = 0xFFFFFFFFu;
            return NVML_SUCCESS;
        }
    }
    usleep((useconds_t)(timeout_ms > 200 ? 200 : timeout_ms) * 1000);
    return 10;   /* NVML_ERROR_TIMEOUT */
}
EXPORT nvmlReturn_t nvmlEventSetWait_v2(void *set, fake_eventdata *d, unsigned timeout_ms) { (void)set; return event_wait(d, timeout_ms); }
EXPORT nvmlReturn_t nvmlEventSetWait(void *set, fake_eventdata *d, unsigned timeout_ms) { (void)set; return event_wait(d, timeout_ms); }
EXPORT nvmlReturn_t nvmlDeviceGetUtilizationRates(nvmlDevice_t h, unsigned *u) { (void)h; u[0] = g_sm_util; u[1] = 0; return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlSystemGetDriverVersion(char *v, unsigned len) { snprintf(v, len, "580.159"); return NVML_SUCCESS; }
EXPORT nvmlReturn_t nvmlSystemGetCudaDriverVersion(int *v) { *v = 12090; return NVML_SUCCESS; }
