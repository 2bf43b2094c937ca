"""Engine integrity fuzz with a full re-check after EVERY operation (the in-suite version checks sampled buffers): the C-ABI
swap engine on the functional fake driver — ragged allocations, frees, touches, two-operand admissions — every word of every
live buffer verified after each step. One JSON line; exit 1 with the step on the first corrupted word.
  env FAKE_GPU_EXEC=1 FAKE_GPU_CTX_MIB=16 LD_LIBRARY_PATH=oracle/_ref/fake VGPU_ROOT=$PWD FUZZ_SEED=100 FUZZ_STEPS=300 \\
      VGPU_SWAP_CHUNK_MB=4 VGPU_SWAP_ARENA_GB=8 VGPU_SWAP_SLAB_MB=64 VGPU_SWAP_SPARE_MB=16 [VGPU_SWAP_RING=2 ...] python tests/tools/engine_fuzz.py
(profiles/r01_engine_fuzz_sweep_cpu.jsonl: 5 configurations x 13 seeds.)"""

import ctypes as C, json, os, random, sys
sys.path.insert(0, os.environ["VGPU_ROOT"])
import k8s_device_plugin_b200 as v
L = v.lib()
drv = C.CDLL("libcuda.so.1")
ctx = C.c_void_p(); dev = C.c_int(0)
assert drv.cuInit(0) == 0 and drv.cuDeviceGet(C.byref(dev), 0) == 0 and drv.cuDevicePrimaryCtxRetain(C.byref(ctx), dev) == 0 and drv.cuCtxSetCurrent(ctx) == 0
M = 1 << 20
rng = random.Random(int(os.environ["FUZZ_SEED"]))
CAP = 48 * M
sw = v.Swap(dev=0, resident_cap=CAP, chunk_bytes=4 * M)
live = {}            # id -> [ptr, nbytes, fill index, touches]
nid = 0
bad = (C.c_uint64 * 1)(0)
peak_resident = 0
ops = {"alloc": 0, "free": 0, "touch": 0, "pair": 0, "verify": 0, "refused": 0}
def resident():
    return sum((r.size + 2 * M - 1) // (2 * M) * (2 * M) for r in sw.table() if r.state & 1)
for step in range(int(os.environ.get("FUZZ_STEPS", "260"))):
    r = rng.random()
    if r < 0.22 or len(live) < 3:
        n = rng.choice([2 * M, 3 * M + 4096, 5 * M + 8, 7 * M, 8 * M + 256 * 3, 12 * M, 16 * M + 64, 20 * M, 2 * M + 8])
        n -= n % 8
        try:
            p = sw.alloc(n)
        except Exception:
            ops["refused"] += 1
            continue
        sw.acquire([p], 0); L.vgpu_wl_fill(C.c_uint64(p), C.c_uint64(n // 8), C.c_uint64(nid), None); sw.release([p], 0)
        live[nid] = [p, n, nid, 0]; nid += 1; ops["alloc"] += 1
    elif r < 0.34:
        k = rng.choice(list(live)); sw.free(live.pop(k)[0]); ops["free"] += 1
    elif r < 0.70:
        k = rng.choice(list(live)); e = live[k]
        sw.acquire([e[0]], 0); L.vgpu_wl_touch(C.c_uint64(e[0]), C.c_uint64(e[1] // 8), None); sw.release([e[0]], 0); e[3] += 1; ops["touch"] += 1
    elif r < 0.85 and len(live) >= 2:
        a, b = rng.sample(list(live), 2)
        if (live[a][1] + 2 * M - 1) // (2 * M) * 2 * M + (live[b][1] + 2 * M - 1) // (2 * M) * 2 * M <= int(os.environ.get('FUZZ_PAIR_CAP', CAP)):
            ptrs = [live[a][0], live[b][0]]
            sw.acquire(ptrs, 0)                       # one admission, two operands: neither may evict the other
            for k in (a, b):
                L.vgpu_wl_touch(C.c_uint64(live[k][0]), C.c_uint64(live[k][1] // 8), None); live[k][3] += 1
            sw.release(ptrs, 0); ops["pair"] += 1
    else:
        k = rng.choice(list(live)); e = live[k]
        sw.acquire([e[0]], 0); L.vgpu_wl_verify(C.c_uint64(e[0]), C.c_uint64(e[1] // 8), C.c_uint64(e[2]), C.c_uint64(e[3]), C.c_uint64(C.addressof(bad)), None); sw.release([e[0]], 0)
        ops["verify"] += 1
    if step % 10 == 0:
        peak_resident = max(peak_resident, resident())
    before = int(bad[0])
    for kk, e in live.items():
        sw.acquire([e[0]], 0); L.vgpu_wl_verify(C.c_uint64(e[0]), C.c_uint64(e[1] // 8), C.c_uint64(e[2]), C.c_uint64(e[3]), C.c_uint64(C.addressof(bad)), None); sw.release([e[0]], 0)
        if int(bad[0]) != before:
            print("CORRUPT at step", step, "r=%.3f" % r, "buffer", kk, e[1], "touches", e[3], "bad", int(bad[0]) - before, file=sys.stderr); sys.exit(1)
for k, e in live.items():
    sw.acquire([e[0]], 0); L.vgpu_wl_verify(C.c_uint64(e[0]), C.c_uint64(e[1] // 8), C.c_uint64(e[2]), C.c_uint64(e[3]), C.c_uint64(C.addressof(bad)), None); sw.release([e[0]], 0)
st = sw.stats()
print(json.dumps({"bad": int(bad[0]), "peak_resident": peak_resident, "ops": ops, "live": st["live_bytes"], "expect_live": sum(e[1] for e in live.values()),
                  "entries": st["entries"], "expect_entries": len(live), "faults": st["faults"], "evictions": st["evictions"]}))
