// driver.cc — see driver.h.
#include "driver.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "log.h"

namespace vgpu {

int log_level() {
    static int lvl = [] {
        const char *e = std::getenv("LIBCUDA_LOG_LEVEL");
        return e ? std::atoi(e) : 2;  // reference default prints Warn/Msg (level test ">1" with unset == print)
    }();
    return lvl;
}

using dlsym_fn = void *(*)(void *, const char *);

void *real_dlsym(void *handle, const char *name) {
    // The hook library exports an unversioned `dlsym` that interposes the C library's; the versioned libc symbol is
    // still reachable with dlvsym (which is not interposed). GLIBC_2.2.5 is the x86-64 baseline version and is kept
    // as a compat symbol after glibc 2.34 moved libdl into libc; the reference does the same lookup
    // (libvgpu.so@0x11b36 init_dlsym) but then falls back to the private _dl_sym, which no longer exists.
    static dlsym_fn fn = [] {
        void *p = dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.2.5");
        if (!p) p = dlvsym(RTLD_NEXT, "dlsym", "GLIBC_2.34");
        if (!p) p = dlvsym(RTLD_DEFAULT, "dlsym", "GLIBC_2.2.5");
        return reinterpret_cast<dlsym_fn>(p);
    }();
    if (!fn) return nullptr;
    return fn(handle, name);
}

static DriverTable g_drv;
static NvmlTable g_nvml;
static std::once_flag g_drv_once, g_nvml_once, g_nvml_init_once;

const DriverTable &drv() {
    std::call_once(g_drv_once, [] {
        const char *lib = std::getenv("VGPU_LIBCUDA_PATH");  // test override
        void *h = dlopen(lib ? lib : "libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            LOG_WARN("libcuda.so.1 not loadable: %s", dlerror());
            return;
        }
        g_drv.handle = h;
#define X(name)                                                            \
    g_drv.name = reinterpret_cast<decltype(g_drv.name)>(real_dlsym(h, #name)); \
    if (!g_drv.name) LOG_DEBUG("driver lacks %s", #name);
        VGPU_DRV_FUNCS(X)
#undef X
#define X(name, sfx) g_drv.name##sfx = reinterpret_cast<decltype(g_drv.name##sfx)>(real_dlsym(h, #name #sfx));
        VGPU_DRV_PT_FUNCS(X)
#undef X
        g_drv.cuGetProcAddress_v1 =
            reinterpret_cast<decltype(g_drv.cuGetProcAddress_v1)>(real_dlsym(h, "cuGetProcAddress"));
        if (!g_drv.cuDeviceGetUuid_v2)  // pre-11.4 drivers
            g_drv.cuDeviceGetUuid_v2 = reinterpret_cast<decltype(g_drv.cuDeviceGetUuid_v2)>(real_dlsym(h, "cuDeviceGetUuid"));
        g_drv.loaded = g_drv.cuInit != nullptr;
    });
    return g_drv;
}

const NvmlTable &nvml() {
    std::call_once(g_nvml_once, [] {
        const char *lib = std::getenv("VGPU_LIBNVML_PATH");
        void *h = dlopen(lib ? lib : "libnvidia-ml.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            LOG_INFO("libnvidia-ml.so.1 not loadable: %s", dlerror());
            return;
        }
        g_nvml.handle = h;
#define X(name) g_nvml.name = reinterpret_cast<decltype(g_nvml.name)>(real_dlsym(h, #name));
        VGPU_NVML_FUNCS(X)
#undef X
        if (!g_nvml.nvmlDeviceGetComputeRunningProcesses_v3)
            g_nvml.nvmlDeviceGetComputeRunningProcesses_v3 =
                reinterpret_cast<decltype(g_nvml.nvmlDeviceGetComputeRunningProcesses_v3)>(
                    real_dlsym(h, "nvmlDeviceGetComputeRunningProcesses_v2"));
        g_nvml.loaded = g_nvml.nvmlInit_v2 != nullptr;
    });
    return g_nvml;
}

bool nvml_ready() {
    const NvmlTable &t = nvml();
    if (!t.loaded) return false;
    std::call_once(g_nvml_init_once, [] {
        nvmlReturn_t r = g_nvml.nvmlInit_v2();
        g_nvml.inited = (r == NVML_SUCCESS);
        if (!g_nvml.inited) LOG_WARN("nvmlInit_v2 failed: %d", (int)r);
    });
    return g_nvml.inited;
}

void *real_cuda_symbol(const char *name) {
    const DriverTable &t = drv();
    return t.handle ? real_dlsym(t.handle, name) : nullptr;
}
void *real_nvml_symbol(const char *name) {
    const NvmlTable &t = nvml();
    return t.handle ? real_dlsym(t.handle, name) : nullptr;
}

const char *cu_err(CUresult r) {
    const char *s = nullptr;
    if (drv().cuGetErrorName && drv().cuGetErrorName(r, &s) == CUDA_SUCCESS && s) return s;
    return "CUDA_ERROR_?";
}

}  // namespace vgpu
