"""Why is the in-process engine arm slower than the hooked app? Same loop, a few variants, host-time breakdown."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import k8s_device_plugin_b200 as v

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
variant = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "bind" in variant:
    print("numa", bench.bind_to_gpu_numa(torch, 0))
sampler = None
if "smi" in variant:
    sampler = bench.ClockSampler(0); sampler.start()
L = v.lib()
st = torch.cuda.current_stream().cuda_stream
if "stream" in variant:
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); st = s.cuda_stream
stp = C.c_void_p(st)
sw = v.Swap(dev=0, resident_cap=8192 << 20, profile="prof" in variant)
nbuf, nbytes = 384, 64 << 20
bufs = []
for i in range(nbuf):
    p = sw.alloc(nbytes); bufs.append(p)
    sw.acquire([p], st); L.vgpu_wl_fill(p, nbytes // 8, i, stp); sw.release([p], st)
torch.cuda.synchronize()
pos = 0
def touch(n):
    global pos
    for _ in range(n):
        p = bufs[pos % nbuf]; pos += 1
        sw.acquire([p], st); L.vgpu_wl_touch(p, nbytes // 8, stp); sw.release([p], st)
touch(64); torch.cuda.synchronize(); sw.drain()
s0 = sw.stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record(); touch(512); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); sw.drain()
ms = e0.elapsed_time(e1); s1 = sw.stats()
d = {k: s1[k] - s0[k] for k in s1}
print(json.dumps({"variant": variant, "GBps": round((d["page_in_bytes"] + d["page_out_bytes"]) / ms / 1e6, 1), "ms": round(ms, 1), "enqueue_ms": round((t1 - t0) * 1e3, 1),
                  "host_ms": {k[5:-3]: round(d[k] / 1e6, 1) for k in d if k.startswith("host_")}, "scans": d["scans"], "reuses": d["phys_reuses"]}))
if sampler: sampler.finish()
