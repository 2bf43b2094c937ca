/*
 * ref_kat.c — TEST INFRASTRUCTURE. Calls two pure functions INSIDE the reference binary (both are exported:
 * `nm -D lib/nvidia/libvgpu.so` lists delta@0x45c7b and get_limit_from_env@0x40d00) to produce known-answer vectors
 * for oracle/vgpu_oracle.c. delta() reads the process-local globals g_sm_num@0x6113c, g_max_thread_per_sm@0x61140,
 * g_total_cuda_cores@0x61148; they are not exported, so they are poked through their
 * offsets from the load base (base = &delta - 0x45c7b). Prints JSON.
 *   usage: ref_kat <path to reference libvgpu.so>     (LD_LIBRARY_PATH must offer a libcuda.so.1, e.g. _ref/fake)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    void *h = dlopen(argv[1], RTLD_LAZY | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    int (*delta)(int, int, int) = (int (*)(int, int, int))dlsym(h, "delta");
    uint64_t (*getlim)(const char *) = (uint64_t(*)(const char *))dlsym(h, "get_limit_from_env");
    if (!delta || !getlim) { fprintf(stderr, "symbols missing\n"); return 1; }
    char *base = (char *)delta - 0x45c7b;
    int *g_sm_num = (int *)(base + 0x6113c), *g_thr = (int *)(base + 0x61140), *g_total = (int *)(base + 0x61148);
    *g_sm_num = 148; *g_thr = 2048; *g_total = 148 * 2048 * 32;
    if (argc > 2 && !strcmp(argv[2], "--stdin")) {
        /* bulk mode for differential fuzzing: "L <text>" -> get_limit_from_env, "D up cur share" -> delta (B200 geometry) */
        char line[512];
        while (fgets(line, sizeof line, stdin)) {
            size_t n = strlen(line);
            if (n && line[n - 1] == '\n') line[n - 1] = 0;
            if (line[0] == 'L' && line[1] == ' ') { setenv("CUDA_DEVICE_MEMORY_LIMIT_9", line + 2, 1); printf("%lu\n", (unsigned long)getlim("CUDA_DEVICE_MEMORY_LIMIT_9")); }
            else if (line[0] == 'D') { int a, b, c; if (sscanf(line + 1, "%d %d %d", &a, &b, &c) == 3) printf("%d\n", delta(a, b, c)); }
        }
        return 0;
    }
    static const int dv[][3] = {{30, 0, 0}, {30, 28, 1000000}, {30, 40, 5000000}, {50, 0, 0}, {30, 100, 5000000}, {30, 30, 9699328},
                                {100, 0, 0}, {10, 9, 123456}, {10, 90, 9000000}, {75, 20, 3000000}, {1, 0, 0}, {99, 100, 42}};
    printf("{\"sm_num\": 148, \"max_threads_per_sm\": 2048, \"total_cores\": %d, \"delta\": [", *g_total);
    for (unsigned i = 0; i < sizeof dv / sizeof dv[0]; i++)
        printf("%s[%d, %d, %d, %d]", i ? ", " : "", dv[i][0], dv[i][1], dv[i][2], delta(dv[i][0], dv[i][1], dv[i][2]));
    /* a second geometry (V100: 80 SMs x 2048) so the non-overflowing regime is pinned too */
    *g_sm_num = 80; *g_thr = 2048; *g_total = 80 * 2048 * 32;
    printf("], \"delta_v100\": [");
    for (unsigned i = 0; i < sizeof dv / sizeof dv[0]; i++)
        printf("%s[%d, %d, %d, %d]", i ? ", " : "", dv[i][0], dv[i][1], dv[i][2], delta(dv[i][0], dv[i][1], dv[i][2]));
    static const char *lv[] = {"8192m", "8g", "8G", "1000k", "4096", "0m", "", "m", "17179869184g", "0x10m", "12abc", "7K", "3M", "1.5g", "-1", " 64m"};
    printf("], \"limit\": [");
    for (unsigned i = 0; i < sizeof lv / sizeof lv[0]; i++) {
        setenv("CUDA_DEVICE_MEMORY_LIMIT_9", lv[i], 1);
        printf("%s[\"%s\", %lu]", i ? ", " : "", lv[i], (unsigned long)getlim("CUDA_DEVICE_MEMORY_LIMIT_9"));
    }
    printf("]}\n");
    return 0;
}
