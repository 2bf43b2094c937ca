/*
 * trace_replay.c — TEST INFRASTRUCTURE (oracle/). A pure CUDA-driver-API program (no cudart, so the
 * reference's cuGetExportTable hooks stay out of the picture — SURVEY.md §7 hard part (d)) that
 * replays an allocation trace and, after every op, prints the return code plus the accounting words
 * of the container's shared region (Go mirror cmd/vGPUmonitor/cudevshr.go:18-58).
 *
 * Run it three ways and diff the output streams:
 *   LD_PRELOAD=dlsym_shim.so:_ref/libvgpu.so   (the reference binary — the real oracle)
 *   LD_PRELOAD=<new>/libvgpu.so                (the product)
 *   no preload, mode "model"                   (oracle/vgpu_oracle.c, the CPU restatement)
 * with LD_LIBRARY_PATH pointing at oracle/_ref/fake (CPU box) or the real driver (GPU box).
 *
 * Trace grammar, one op per line (ids index a pointer table):
 *   A id size | M id size (managed) | P id width height (pitch, elem 4) | F id | X hexaddr (free raw)
 *   I (cuMemGetInfo_v2) | T (cuDeviceTotalMem_v2) | L gx gy gz (cuLaunchKernel of an empty kernel) | S ms (sleep)
 *   Y id size (cuMemAllocAsync) | Z id (cuMemFreeAsync) | C id size (cuMemCreate on device 0) | R id (cuMemRelease)
 *   G (cuGraphLaunch of a null graph — only meaningful on the fake driver)
 *   D n (make device n's primary context current; the counters printed from then on are device n's lane)
 *   N (nvmlDeviceGetMemoryInfo of NVML device 0 — the hook's exported wrapper when one is preloaded; prints
 *      " nv_total=.. nv_free=.. nv_used=..")
 *   U ms (sleep until the absolute time <ms since the epoch>: lets several replayers act on one shared schedule)
 *   K (kill(getpid(), SIGKILL): a process that dies without running its exit handler)
 *   W id byte (cuMemsetD8_v2 of the whole buffer) | O dst src (cuMemcpyDtoD_v2 of min(size)) | V id byte (cuMemcpyDtoH_v2 of
 *      the whole buffer and a check that every byte equals <byte>; prints " ok=<0|1>") — data ops for the swap path
 *   Q id (cuPointerGetAttributes {MEMORY_TYPE, IS_MANAGED} of pointer id; prints " type=<n> managed=<n>")
 * Output line: "<op#> <opcode> rc=<int> ctx=<u64> mod=<u64> buf=<u64> off=<u64> tot=<u64> [free=.. total=..]"
 * where the five counters are SUMMED over every process slot of device 0 (== own slot for one process).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <signal.h>
#include <time.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
typedef void *CUfunction;
typedef void *CUmodule;
typedef void *CUstream;

extern CUresult cuInit(unsigned);
extern CUresult cuDeviceGet(CUdevice *, int);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, CUdevice);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemHostAlloc(void **, size_t, unsigned);
extern CUresult cuMemAllocHost_v2(void **, size_t);
extern CUresult cuMemFreeHost(void *);
extern CUresult cuMemHostRegister_v2(void *, size_t, unsigned) __attribute__((weak));
extern CUresult cuMemHostUnregister(void *) __attribute__((weak));
extern CUresult cuMipmappedArrayCreate(void **, const void *, unsigned) __attribute__((weak));
extern CUresult cuMipmappedArrayDestroy(void *) __attribute__((weak));
extern int fake_live_mipmaps(void) __attribute__((weak));
extern int fake_host_unregisters(void) __attribute__((weak));
extern CUresult cuCtxCreate_v2(CUcontext *, unsigned, CUdevice);
extern CUresult cuCtxDestroy_v2(CUcontext);
extern CUresult cuCtxGetDevice(CUdevice *);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemAllocManaged(CUdeviceptr *, size_t, unsigned);
extern CUresult cuMemAllocPitch_v2(CUdeviceptr *, size_t *, size_t, size_t, unsigned);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);
extern CUresult cuDeviceTotalMem_v2(size_t *, CUdevice);
extern CUresult cuLaunchKernel(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **);
extern CUresult cuModuleLoadData(CUmodule *, const void *) __attribute__((weak));
extern CUresult cuModuleGetFunction(CUfunction *, CUmodule, const char *) __attribute__((weak));
extern CUresult cuCtxSynchronize(void);
extern CUresult cuMemsetD8_v2(CUdeviceptr, unsigned char, size_t);
extern CUresult cuMemcpyDtoD_v2(CUdeviceptr, CUdeviceptr, size_t);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);
/* entry points the reference forwards untouched */
extern CUresult cuMemAllocAsync(CUdeviceptr *, size_t, CUstream) __attribute__((weak));
extern CUresult cuMemFreeAsync(CUdeviceptr, CUstream) __attribute__((weak));
extern CUresult cuMemCreate(unsigned long long *, size_t, const void *, unsigned long long) __attribute__((weak));
extern CUresult cuMemRelease(unsigned long long) __attribute__((weak));
extern CUresult cuMemAddressReserve(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) __attribute__((weak));
extern CUresult cuMemMap(CUdeviceptr, size_t, size_t, unsigned long long, unsigned long long) __attribute__((weak));
extern CUresult cuMemUnmap(CUdeviceptr, size_t) __attribute__((weak));
extern CUresult cuGraphLaunch(void *, CUstream) __attribute__((weak));
extern CUresult cuPointerGetAttributes(unsigned, int *, void **, CUdeviceptr) __attribute__((weak));
/* NVML: bound to a preloaded hook's exported wrappers when there is one (the reference run preloads a dlsym shim that
 * would hide its dlsym override, so symbol interposition is the route compared), else found with dlopen + dlsym */
extern int nvmlInit_v2(void) __attribute__((weak));
extern int nvmlDeviceGetHandleByIndex_v2(unsigned, void **) __attribute__((weak));
extern int nvmlDeviceGetMemoryInfo(void *, void *) __attribute__((weak));
struct mem_prop { int type; int requested_handle_types; struct { int type; int id; } location; void *win32; struct { unsigned char c, g; unsigned short u; unsigned char r[4]; } flags; };

/* Appendix A offsets */
#define REGION_SIZE 0xC4748
#define OFF_PROCS 0x738
#define SLOT_STRIDE 0x310
#define OFF_USED 0x8
#define USED_STRIDE 40
#define OFF_PROCNUM 0xC4738
#define MAXPROC 1024

static const unsigned char *g_region;

static void map_region(void) {
    const char *path = getenv("CUDA_DEVICE_MEMORY_SHARED_CACHE");
    if (!path) path = "/tmp/cudevshr.cache";
    int fd = open(path, O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    if (fstat(fd, &st) == 0 && st.st_size >= REGION_SIZE)
        g_region = mmap(NULL, REGION_SIZE, PROT_READ, MAP_SHARED, fd, 0);
    if (g_region == MAP_FAILED) g_region = NULL;
    close(fd);
}

static void counters(int dev, uint64_t out[5]) {
    memset(out, 0, 5 * sizeof(uint64_t));
    if (!g_region) return;
    int32_t procnum; memcpy(&procnum, g_region + OFF_PROCNUM, 4);   /* live slots only: the reference compacts without clearing */
    for (int s = 0; s < procnum && s < MAXPROC; s++) {
        const unsigned char *slot = g_region + OFF_PROCS + (size_t)s * SLOT_STRIDE;
        int32_t pid; memcpy(&pid, slot, 4);
        if (pid == 0) continue;
        const unsigned char *u = slot + OFF_USED + (size_t)dev * USED_STRIDE;
        for (int k = 0; k < 5; k++) { uint64_t v; memcpy(&v, u + 8 * k, 8); out[k] += v; }
    }
}

static const char *k_ptx =
    ".version 7.0\n.target sm_52\n.address_size 64\n.visible .entry vgpu_empty()\n{\n ret;\n}\n";

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s trace.txt [nptr]\n", argv[0]); return 2; }
    FILE *tf = fopen(argv[1], "r");
    if (!tf) { perror("trace"); return 2; }
    size_t nptr = argc > 2 ? strtoull(argv[2], 0, 0) : (1u << 20);
    CUdeviceptr *ptrs = calloc(nptr, sizeof *ptrs);
    size_t *sizes = calloc(nptr, sizeof *sizes);

    CUdevice dev; CUcontext ctx;
    CUresult r;
    if ((r = cuInit(0))) { printf("cuInit rc=%d\n", r); return 1; }
    if ((r = cuDeviceGet(&dev, 0))) { printf("cuDeviceGet rc=%d\n", r); return 1; }
    if ((r = cuDevicePrimaryCtxRetain(&ctx, dev))) { printf("retain rc=%d\n", r); return 1; }
    if ((r = cuCtxSetCurrent(ctx))) { printf("setcurrent rc=%d\n", r); return 1; }
    map_region();

    CUfunction fn = NULL;
    if (!getenv("TRACE_NO_MODULE") && cuModuleLoadData && cuModuleGetFunction) {
        CUmodule mod;
        if (cuModuleLoadData(&mod, k_ptx) == 0) cuModuleGetFunction(&fn, mod, "vgpu_empty");
    }

    uint64_t c[5];
    counters(0, c);
    printf("init rc=0 ctx=%lu mod=%lu buf=%lu off=%lu tot=%lu region=%d\n", c[0], c[1], c[2], c[3], c[4], g_region != NULL);
    const int flush_each = getenv("TRACE_FLUSH") != NULL;   /* op-by-op driving over pipes (tests/tools/multiproc_fuzz.py) */
    if (flush_each) fflush(stdout);

    char line[256];
    unsigned long opn = 0;
    int cur = 0;                       /* device whose lane is reported */
    CUcontext ctxs[16] = {ctx};
    CUcontext uctx[16] = {0}; int udev[16] = {0};   /* contexts made with cuCtxCreate_v2 (ops E H e) */
    while (fgets(line, sizeof line, tf)) {
        char op; unsigned long long a = 0, b = 0, d = 0;
        if (line[0] == '#' || line[0] == '\n') continue;
        int n = sscanf(line, " %c %lli %lli %lli", &op, (long long *)&a, (long long *)&b, (long long *)&d);
        if (n < 1) continue;
        size_t fr = 0, tot = 0; int has_x = 0, xval = 0, has_info = 0, has_q = 0, qtype = 0, qman = 0, has_nv = 0, has_v = 0, vok = 0;
        unsigned long long nvmem[3] = {0, 0, 0};     /* nvmlMemory_t {total, free, used} */
        switch (op) {
        case 'A': ptrs[a] = 0; r = cuMemAlloc_v2(&ptrs[a], (size_t)b); if (r) ptrs[a] = 0; else sizes[a] = (size_t)b; break;
        case 'W': r = ptrs[a] ? cuMemsetD8_v2(ptrs[a], (unsigned char)b, sizes[a]) : 1; break;
        case 'O': r = (ptrs[a] && ptrs[b]) ? cuMemcpyDtoD_v2(ptrs[a], ptrs[b], sizes[a] < sizes[b] ? sizes[a] : sizes[b]) : 1; break;
        case 'V': { r = 1; vok = 0; has_v = 1;
                    if (ptrs[a]) { unsigned char *h = malloc(sizes[a]); r = cuMemcpyDtoH_v2(h, ptrs[a], sizes[a]);
                        if (!r) { vok = 1; for (size_t i = 0; i < sizes[a]; i++) if (h[i] != (unsigned char)b) { vok = 0; break; } }
                        free(h); }
                    break; }
        case 'M': ptrs[a] = 0; r = cuMemAllocManaged(&ptrs[a], (size_t)b, 1); if (r) ptrs[a] = 0; break;
        case 'P': { size_t pitch = 0; ptrs[a] = 0; r = cuMemAllocPitch_v2(&ptrs[a], &pitch, (size_t)b, (size_t)d, 4); if (r) ptrs[a] = 0; break; }
        case 'F': r = cuMemFree_v2(ptrs[a]); if (!r) ptrs[a] = 0; break;
        case 'X': r = cuMemFree_v2((CUdeviceptr)a); break;
        case 'I': r = cuMemGetInfo_v2(&fr, &tot); has_info = 1; break;
        case 'T': r = cuDeviceTotalMem_v2(&tot, dev); fr = 0; has_info = 1; break;
        case 'L': r = cuLaunchKernel(fn, (unsigned)a, (unsigned)b, (unsigned)d, 1, 1, 1, 0, NULL, NULL, NULL); break;
        case 'S': usleep((useconds_t)a * 1000); r = 0; break;   /* sleep a ms (multi-process tests) */
        case 'U': { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
                    long long nowms = (long long)ts.tv_sec * 1000 + ts.tv_nsec / 1000000, wait = (long long)a - nowms;
                    if (wait > 0) usleep((useconds_t)wait * 1000);
                    r = 0; break; }
        case 'K': fflush(stdout); kill(getpid(), SIGKILL); r = 0; break;
        case 'D': { int n2 = (int)a & 15; CUdevice d2; r = cuDeviceGet(&d2, n2);
                    if (!r && !ctxs[n2]) r = cuDevicePrimaryCtxRetain(&ctxs[n2], d2);
                    if (!r) r = cuCtxSetCurrent(ctxs[n2]);
                    if (!r) { cur = n2; dev = d2; }
                    break; }
        /* context family: B n = retain device n's primary context once more; E s n = cuCtxCreate_v2 on device n into
         * slot s (becomes current); H s = make slot s current; e s = destroy slot s */
        case 'B': { CUcontext t = NULL; CUdevice d2;
                    r = cuDeviceGet(&d2, (int)a & 15);
                    if (!r) r = cuDevicePrimaryCtxRetain(&t, d2);
                    if (!r && !ctxs[a & 15]) ctxs[a & 15] = t;
                    break; }
        case 'E': { CUdevice d2; r = cuDeviceGet(&d2, (int)b & 15); if (!r) r = cuCtxCreate_v2(&uctx[a & 15], 0, d2);
                    if (!r) { udev[a & 15] = (int)b & 15; cur = (int)b & 15; dev = d2; } break; }
        case 'H': r = uctx[a & 15] ? cuCtxSetCurrent(uctx[a & 15]) : 1; if (!r) { cur = udev[a & 15]; dev = cur; } break;
        case 'e': r = uctx[a & 15] ? cuCtxDestroy_v2(uctx[a & 15]) : 1; if (!r) uctx[a & 15] = NULL; break;
        /* host-side family (the reference only runs check_oom after the real call): h n = cuMemHostAlloc(n) + free,
         * a n = cuMemAllocHost_v2(n) + free, r n = cuMemHostRegister_v2 of n malloc'd bytes (+ unregister),
         * m = cuMipmappedArrayCreate (+ destroy); the extra field says what was left behind on the driver side */
        case 'h': { void *hp = NULL; r = cuMemHostAlloc(&hp, (size_t)a, 0); if (!r) cuMemFreeHost(hp); break; }
        case 'a': { void *hp = NULL; r = cuMemAllocHost_v2(&hp, (size_t)a); if (!r) cuMemFreeHost(hp); break; }
        case 'r': { void *hp = malloc((size_t)a); int before = fake_host_unregisters ? fake_host_unregisters() : 0;
                    r = cuMemHostRegister_v2 ? cuMemHostRegister_v2(hp, (size_t)a, 0) : 801;
                    has_x = 1; xval = (fake_host_unregisters ? fake_host_unregisters() : 0) - before;   /* undone by the hook? */
                    if (!r && cuMemHostUnregister) cuMemHostUnregister(hp);
                    free(hp); break; }
        case 'm': { unsigned long long desc[5] = {64, 64, 0, 0x20 /* FLOAT */ | (1ull << 32), 0}; void *arr = NULL;
                    r = cuMipmappedArrayCreate ? cuMipmappedArrayCreate(&arr, desc, 2) : 801;
                    if (!r && cuMipmappedArrayDestroy) cuMipmappedArrayDestroy(arr);
                    has_x = 1; xval = fake_live_mipmaps ? fake_live_mipmaps() : 0; break; }
        case 'Y': ptrs[a] = 0; r = cuMemAllocAsync ? cuMemAllocAsync(&ptrs[a], (size_t)b, NULL) : 801; if (r) ptrs[a] = 0; break;
        case 'Z': r = cuMemFreeAsync ? cuMemFreeAsync(ptrs[a], NULL) : 801; if (!r) ptrs[a] = 0; break;
        case 'C': { struct mem_prop pr; memset(&pr, 0, sizeof pr); pr.type = 1 /* PINNED */; pr.location.type = 1 /* DEVICE */; pr.location.id = 0;
                    ptrs[a] = 0; r = cuMemCreate ? cuMemCreate(&ptrs[a], (size_t)b, &pr, 0) : 801; if (r) ptrs[a] = 0; else sizes[a] = (size_t)b; break; }
        /* application-side VMM mappings: p s = map the handle of slot s into its window of a reserved range, u s = unmap it
         * (the handle may have been released in between: the CUDA samples release right after mapping) */
        case 'p': case 'u': {
            static CUdeviceptr vbase; static size_t vsz[64]; const size_t stride = (size_t)256 << 20; unsigned k = (unsigned)a & 63;
            r = 801;
            if (!cuMemAddressReserve || !cuMemMap || !cuMemUnmap) break;
            if (!vbase && (r = cuMemAddressReserve(&vbase, 64 * stride, 0, 0, 0))) break;
            if (op == 'p') { r = ptrs[a] ? cuMemMap(vbase + k * stride, sizes[a], 0, ptrs[a], 0) : 1; if (!r) vsz[k] = sizes[a]; }
            else { r = vsz[k] ? cuMemUnmap(vbase + k * stride, vsz[k]) : 1; if (!r) vsz[k] = 0; }
            break; }
        case 'R': r = cuMemRelease ? cuMemRelease(ptrs[a]) : 801; if (!r) ptrs[a] = 0; break;
        case 'G': r = cuGraphLaunch ? cuGraphLaunch(NULL, NULL) : 801; break;
        case 'N': { static void *nv; static int (*init)(void), (*byidx)(unsigned, void **), (*meminfo)(void *, void *);
                    if (!init && nvmlInit_v2 && nvmlDeviceGetHandleByIndex_v2 && nvmlDeviceGetMemoryInfo) {
                        init = nvmlInit_v2; byidx = nvmlDeviceGetHandleByIndex_v2; meminfo = nvmlDeviceGetMemoryInfo; init();
                    } else if (!init && !nv && (nv = dlopen("libnvidia-ml.so.1", RTLD_NOW))) {
                        init = (int (*)(void))dlsym(nv, "nvmlInit_v2"); byidx = (int (*)(unsigned, void **))dlsym(nv, "nvmlDeviceGetHandleByIndex_v2");
                        meminfo = (int (*)(void *, void *))dlsym(nv, "nvmlDeviceGetMemoryInfo");
                        if (init) init();
                    }
                    void *h = NULL; r = 801;
                    if (byidx && meminfo && byidx(0, &h) == 0) { r = meminfo(h, nvmem); has_nv = 1; }
                    break; }
        case 'Q': { int attrs[2] = {2, 8}; unsigned ty = 0, mg = 0; void *data[2] = {&ty, &mg};
                    r = cuPointerGetAttributes ? cuPointerGetAttributes(2, attrs, data, ptrs[a]) : 801; qtype = (int)ty; qman = (int)mg; has_q = 1; break; }
        default: continue;
        }
        counters(cur, c);
        printf("%lu %c rc=%d ctx=%lu mod=%lu buf=%lu off=%lu tot=%lu", opn, op, r, c[0], c[1], c[2], c[3], c[4]);
        if (has_info) {
            /* free/total only compared when the hook owns them (limit set): the reference reports the
             * real driver's view in the unlimited case */
            printf(" free=%zu total=%zu", fr, tot);
        }
        if (has_q) printf(" type=%d managed=%d", qtype, qman);
        if (has_x) printf(" left=%d", xval);
        if (has_v) printf(" ok=%d", vok);
        if (op == 'L' && g_region && getenv("TRACE_SHOW_WORDS")) {   /* the monitor handshake words next to a launch */
            int32_t w[3]; memcpy(w, g_region + 0xC473C, 12);
            uint64_t sm; memcpy(&sm, g_region + 0x6B8, 8);
            printf(" us=%d rk=%d pr=%d sm0=%lu", w[0], w[1], w[2], (unsigned long)sm);
        }
        if (has_nv) printf(" nv_total=%llu nv_free=%llu nv_used=%llu", nvmem[0], nvmem[1], nvmem[2]);
        putchar('\n');
        if (flush_each) fflush(stdout);
        opn++;
    }
    cuCtxSynchronize();
    fflush(stdout);
    return 0;
}
