/* vgpu_oracle.c — see vgpu_oracle.h. TEST INFRASTRUCTURE, not product code. */
#include "vgpu_oracle.h"
#include <stdlib.h>
#include <string.h>

/* get_limit_from_env@0x40d00: scalar from the LAST character (G/g, M/m, K/k; @0x40d8d-0x40e3a), number by
 * strtoul(base 0) (@0x40e55), product 0 -> 0 (unlimited), overflow check product/scalar != n -> 0 (@0x41064). */
uint64_t vo_parse_limit(const char *v) {
    if (!v) return 0;
    size_t len = strlen(v);
    if (len == 0) return 0;
    uint64_t scalar = 1;
    char c = v[len - 1];
    if (c == 'G' || c == 'g') scalar = 1ull << 30;
    else if (c == 'M' || c == 'm') scalar = 1ull << 20;
    else if (c == 'K' || c == 'k') scalar = 1ull << 10;
    uint64_t n = strtoul(v, NULL, 0);
    uint64_t prod = n * scalar;
    if (prod == 0) return 0;
    if (prod / scalar != n) return 0;
    return prod;
}

struct chunk { uint64_t size, real; int live; }; /* size = accounted bytes, real = bytes taken from the device */
struct vo_state {
    uint64_t limit, dev_total, dev_used;
    uint64_t ctx, mod, buf, off, tot;
    struct chunk *chunks; size_t cap; /* direct-indexed by trace id */
};

vo_state_t *vo_create(uint64_t limit, uint64_t ctx_bytes, uint64_t dev_total) {
    vo_state_t *s = calloc(1, sizeof *s);
    s->limit = limit; s->dev_total = dev_total;
    /* cuDevicePrimaryCtxRetain (context.c:L72-86): add_gpu_device_memory_usage(pid, dev, context_size, 0) */
    s->ctx = ctx_bytes; s->tot = ctx_bytes;
    return s;
}
void vo_destroy(vo_state_t *s) { if (s) { free(s->chunks); free(s); } }

static struct chunk *slot(vo_state_t *s, uint64_t id) {
    if (id >= s->cap) {
        size_t nc = s->cap ? s->cap : 1024;
        while (nc <= id) nc *= 2;
        s->chunks = realloc(s->chunks, nc * sizeof *s->chunks);
        memset(s->chunks + s->cap, 0, (nc - s->cap) * sizeof *s->chunks);
        s->cap = nc;
    }
    return &s->chunks[id];
}

/* oom_check@0x3f9bc (allocator.c:L35-53): limit==0 -> ok; usage+addon > limit (STRICT, ja @0x3fb88) -> OOM.
 * usage = get_gpu_memory_usage@0x420bd = sum of used[dev].total over slots (one slot here). The
 * rm_quitted_process() retry only changes the answer when another slot's pid died; single process: no. */
static int oom(const vo_state_t *s, uint64_t addon) {
    if (s->limit == 0) return 0;
    return s->tot + addon > s->limit;
}
/* the real driver under the hook (fake_driver/fake_gpu.c do_alloc): device OOM at dev_used+size > dev_total */
static int real_alloc(vo_state_t *s, uint64_t size) {
    if (size == 0) return 1; /* CUDA_ERROR_INVALID_VALUE */
    if (s->dev_used + size > s->dev_total) return 2;
    s->dev_used += size;
    return 0;
}
/* add_gpu_device_memory_usage@0x42a0e type 2: total += bytes, bufferSize += bytes */
static void account(vo_state_t *s, uint64_t id, uint64_t size) {
    struct chunk *c = slot(s, id);
    c->size = size; c->real = size; c->live = 1;
    s->buf += size; s->tot += size;
}

int vo_alloc(vo_state_t *s, uint64_t id, uint64_t size) {
    slot(s, id)->live = 0;
    if (oom(s, size)) return -1;                 /* add_chunk@0x4005d */
    int r = real_alloc(s, size);                 /* size<=IPCSIZE: cuMemAlloc_v2, else cuMemoryAllocate@0x315da */
    if (r) return r;                             /* add_chunk@0x40365: returns the driver code */
    account(s, id, size);
    return 0;
}
int vo_alloc_managed(vo_state_t *s, uint64_t id, uint64_t size) {
    slot(s, id)->live = 0;
    if (oom(s, size)) return 2;                  /* @0x31eab */
    int r = real_alloc(s, size);
    if (r) return r;
    account(s, id, size);                        /* add_chunk_only@0x404a5 */
    return 0;
}
int vo_alloc_pitch(vo_state_t *s, uint64_t id, uint64_t width, uint64_t height, unsigned elem) {
    slot(s, id)->live = 0;
    /* @0x3206b-0x32096: guess_pitch = ((W-1)/elem + 1) * elem ; bytesize = guess_pitch * H */
    uint64_t guess = ((width - 1) / elem + 1) * (uint64_t)elem;
    uint64_t bytes = guess * height;
    if (oom(s, bytes)) return 2;                 /* @0x321ef */
    /* the real driver allocates its own pitch (fake: width rounded to 512) x height */
    uint64_t real_bytes = ((width + 511) & ~511ull) * height;
    int r = real_alloc(s, real_bytes);
    if (r) return r;
    struct chunk *c = slot(s, id);
    c->size = bytes; c->real = real_bytes; c->live = 1;   /* add_chunk_only@0x404a5 with the GUESSED size */
    s->buf += bytes; s->tot += bytes;
    return 0;
}
int vo_free(vo_state_t *s, uint64_t id) {
    struct chunk *c = slot(s, id);
    if (!c->live) return 0;                      /* the replayer passes ptr 0 for never-allocated ids: cuMemFree_v2(0) == 0 (@0x32383) */
    s->dev_used -= c->real;
    s->buf -= c->size; s->tot -= c->size;        /* rm_gpu_device_memory_usage@0x42de1 type 2 */
    c->live = 0;
    return 0;
}
int vo_free_untracked(vo_state_t *s, uint64_t addr) { (void)s; return addr ? -1 : 0; } /* remove_chunk@0x409f0 */

int vo_mem_get_info(vo_state_t *s, uint64_t *fr, uint64_t *total) {
    /* cuMemGetInfo_v2@0x367dc (memory.c:L549-566) */
    if (s->limit == 0) { *total = s->dev_total; *fr = s->dev_total - s->tot; return 0; }
    if (s->limit < s->tot) return 1;
    *fr = s->limit - s->tot; *total = s->limit;
    return 0;
}
/* cuDeviceTotalMem_v2@0x2d3f8 (device.c:L197): *bytes = get_current_device_memory_limit(dev), never the real
 * driver — so an UNLIMITED container is told 0 bytes (quirk; the product returns the real total there) */
uint64_t vo_total_mem(vo_state_t *s) { return s->limit; }
void vo_counters(const vo_state_t *s, uint64_t o[5]) { o[0] = s->ctx; o[1] = s->mod; o[2] = s->buf; o[3] = s->off; o[4] = s->tot; }

/* setspec@0x45d5b: g_total_cuda_cores = max_threads_per_sm * sm_num * 32 (int32) */
int32_t vo_total_cuda_cores(int32_t sm, int32_t thr) { return (int32_t)((uint32_t)thr * (uint32_t)sm * 32u); }

/* delta@0x45c7b — every product/sum is 32-bit two's complement like the -O0 imul/add sequence */
static int32_t mul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t sub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
int32_t vo_delta(int32_t sm, int32_t thr, int32_t total, int32_t up, int32_t cur, int32_t share) {
    int32_t d = sub32(up, cur);
    if (d < 0) d = -d;
    if (d < 5) d = 5;
    int32_t inc = mul32(mul32(mul32(sm, sm), thr), d) / 2560;
    if (d > up / 2) inc = mul32(mul32(inc, d), 2) / add32(up, 1);
    if (cur <= up) { int32_t t = add32(share, inc); share = t > total ? total : t; }
    else { int32_t t = sub32(share, inc); share = t < 0 ? 0 : t; }
    return share;
}

/* ---- spec oracles for the new swap kernels */
struct key { uint64_t touch; uint32_t idx; uint64_t size; };
static int keycmp(const void *a, const void *b) {
    const struct key *x = a, *y = b;
    if (x->touch != y->touch) return x->touch < y->touch ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
static int u32cmp(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }
int64_t vo_select_victims(const vo_entry_t *t, uint64_t n, uint64_t need, uint32_t *out, uint64_t *freed) {
    struct key *k = malloc((n ? n : 1) * sizeof *k);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++)
        if (t[i].state == VO_ST_RESIDENT) { k[m].touch = t[i].last_touch; k[m].idx = (uint32_t)i; k[m].size = t[i].size; m++; }
    qsort(k, m, sizeof *k, keycmp);
    uint64_t sum = 0, cnt = 0;
    while (cnt < m && sum < need) { sum += k[cnt].size; out[cnt] = k[cnt].idx; cnt++; }
    if (freed) *freed = sum;
    qsort(out, cnt, sizeof *out, u32cmp);
    free(k);
    if (sum < need) return -1;
    return (int64_t)cnt;
}
void vo_pack(uint8_t *dst, const uint8_t *src, const vo_seg_t *s, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) memcpy(dst + s[i].dst_off, src + s[i].src_off, s[i].bytes);
}
uint64_t vo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
