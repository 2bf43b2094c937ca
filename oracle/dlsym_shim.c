/*
 * dlsym_shim.c — TEST INFRASTRUCTURE (oracle/). Two things the 2021-era reference hook binary
 * (lib/nvidia/libvgpu.so) needs to run on this image / this driver. Preloaded BEFORE the reference binary,
 * never with the new library.
 *
 * 1. `_dl_sym`: the binary was built against glibc 2.31 and falls back to the private
 *    `_dl_sym@GLIBC_PRIVATE` (removed in glibc 2.34) inside its dlsym override (libvgpu.so@0x11b36,
 *    libvgpu.c:L108-123).
 * 2. `dlsym` bypass: driver 580's libnvidia-ml resolves cu* symbols with dlsym() while nvmlInit runs. The reference's dlsym override answers cu* / nvml* names
 *    through pthread_once(preInit) — and the hook calls nvmlInit from INSIDE preInit, so that lookup
 *    re-enters the same once-control and the process deadlocks (seen on the B200 boxes; reproduced here
 *    with FAKE_NVML_DLSYM=1 and no shim, tests/test_oracle_pin.py). With the shim every dlsym() in the process goes
 *    straight to the C library; the reference's hooks still interpose programs that link the driver
 *    API directly (trace_replay, swap_bench), which is all the oracle needs.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>

__attribute__((visibility("default")))
void *_dl_sym(void *handle, const char *name, void *who) {
    (void)who;
    return dlvsym(handle, name, "GLIBC_2.2.5");
}

typedef void *(*dlsym_fn)(void *, const char *);
static dlsym_fn libc_dlsym(void) {
    static dlsym_fn fn;
    if (!fn) {
        void *libc = dlopen("libc.so.6", RTLD_LAZY | RTLD_NOLOAD);
        if (libc) fn = (dlsym_fn)dlvsym(libc, "dlsym", "GLIBC_2.2.5");
        if (!fn && libc) fn = (dlsym_fn)dlvsym(libc, "dlsym", "GLIBC_2.34");
    }
    return fn;
}
__attribute__((visibility("default")))
void *dlsym(void *handle, const char *name) {
    dlsym_fn f = libc_dlsym();
    return f ? f(handle, name) : NULL;
}
