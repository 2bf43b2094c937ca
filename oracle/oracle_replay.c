/* oracle_replay.c — TEST INFRASTRUCTURE. Replays a trace (grammar: trace_replay.c) through the CPU
 * restatement vgpu_oracle.c and prints the SAME line format as trace_replay, so the three streams
 * (reference binary / new library / restatement) can be diffed textually. */
#include "vgpu_oracle.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s trace.txt\n", argv[0]); return 2; }
    FILE *tf = fopen(argv[1], "r"); if (!tf) { perror("trace"); return 2; }
    const char *e;
    uint64_t limit = vo_parse_limit(getenv("CUDA_DEVICE_MEMORY_LIMIT_0") ? getenv("CUDA_DEVICE_MEMORY_LIMIT_0") : getenv("CUDA_DEVICE_MEMORY_LIMIT"));
    uint64_t ctx = ((e = getenv("FAKE_GPU_CTX_MIB")) ? strtoull(e, 0, 0) : 512ull) << 20;
    if ((e = getenv("ORACLE_CTX_BYTES"))) ctx = strtoull(e, 0, 0);   /* exact context size measured on a real box */
    uint64_t tot = ((e = getenv("FAKE_GPU_TOTAL_MIB")) ? strtoull(e, 0, 0) : 183359ull) << 20;
    vo_state_t *s = vo_create(limit, ctx, tot);
    uint64_t c[5]; vo_counters(s, c);
    printf("init rc=0 ctx=%lu mod=%lu buf=%lu off=%lu tot=%lu region=1\n", c[0], c[1], c[2], c[3], c[4]);
    char line[256]; unsigned long opn = 0;
    while (fgets(line, sizeof line, tf)) {
        char op; unsigned long long a = 0, b = 0, d = 0;
        if (line[0] == '#' || line[0] == '\n') continue;
        if (sscanf(line, " %c %lli %lli %lli", &op, (long long *)&a, (long long *)&b, (long long *)&d) < 1) continue;
        int r = 0, info = 0; uint64_t fr = 0, total = 0;
        switch (op) {
        case 'A': r = vo_alloc(s, a, b); break;
        case 'M': r = vo_alloc_managed(s, a, b); break;
        case 'P': r = vo_alloc_pitch(s, a, b, d, 4); break;
        case 'F': r = vo_free(s, a); break;
        case 'X': r = vo_free_untracked(s, a); break;
        case 'I': r = vo_mem_get_info(s, &fr, &total); info = 1; break;
        case 'T': total = vo_total_mem(s); info = 1; break;
        case 'L': r = 0; break;
        default: continue;
        }
        vo_counters(s, c);
        printf("%lu %c rc=%d ctx=%lu mod=%lu buf=%lu off=%lu tot=%lu", opn, op, r, c[0], c[1], c[2], c[3], c[4]);
        if (info) printf(" free=%lu total=%lu", fr, total);
        putchar('\n'); opn++;
    }
    vo_destroy(s);
    return 0;
}
