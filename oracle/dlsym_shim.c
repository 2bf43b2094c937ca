/*
 * dlsym_shim.c — TEST INFRASTRUCTURE (oracle/). The reference hook lib/nvidia/libvgpu.so was
 * built against glibc 2.31 and falls back to the private `_dl_sym@GLIBC_PRIVATE` (removed in
 * glibc 2.34) inside its dlsym override (libvgpu.so@0x11b36, libvgpu.c:L108-123). Preloading this
 * shim BEFORE the reference binary supplies that one symbol so the unmodified binary runs on this
 * image's glibc 2.39. It is never loaded with the new library.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
__attribute__((visibility("default")))
void *_dl_sym(void *handle, const char *name, void *who) {
    (void)who;
    return dlvsym(handle, name, "GLIBC_2.2.5");
}
