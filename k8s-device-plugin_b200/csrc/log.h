// log.h — leveled stderr logging. The level is read ONCE from LIBCUDA_LOG_LEVEL (same variable and level
// numbering as the reference: ERROR always, Warn/Msg >1, Info >2, Debug >3 — SURVEY.md §5), instead of the
// reference's getenv+atoi inside every log macro on every hooked call (e.g. libvgpu.so@0x3f9ea), which is a
// large part of its per-call intercept overhead.
#pragma once
#include <cstdio>
#include <unistd.h>

namespace vgpu {
int log_level();  // cached
}

#define VGPU_LOG_AT(minlvl, tag, fmt, ...)                                                              \
    do {                                                                                                \
        if (vgpu::log_level() >= (minlvl))                                                              \
            std::fprintf(stderr, "[vgpu-b200 " tag "(%d:%s:%d)]: " fmt "\n", (int)getpid(), __FILE_NAME__, \
                         __LINE__, ##__VA_ARGS__);                                                      \
    } while (0)
#define LOG_ERROR(fmt, ...) VGPU_LOG_AT(0, "ERROR", fmt, ##__VA_ARGS__)
#define LOG_WARN(fmt, ...) VGPU_LOG_AT(2, "Warn", fmt, ##__VA_ARGS__)
#define LOG_MSG(fmt, ...) VGPU_LOG_AT(2, "Msg", fmt, ##__VA_ARGS__)
#define LOG_INFO(fmt, ...) VGPU_LOG_AT(3, "Info", fmt, ##__VA_ARGS__)
#define LOG_DEBUG(fmt, ...) VGPU_LOG_AT(4, "Debug", fmt, ##__VA_ARGS__)
