"""Geometry sweep of vgpu_pack_tma on a B200 (tile size x ring stages x CTAs per SM), isolated kernel, 3 GiB per
launch (96 segments x 32 MiB, inputs larger than L2), CUDA events on the launching stream. Writes gpurun_out/pack_sweep.json."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import k8s_device_plugin_b200 as v

torch.zeros(1, device="cuda")
L = v.lib()
seg = 32 << 20
nseg = 96
total = seg * nseg
src = torch.empty(total // 8, dtype=torch.int64, device="cuda").random_()
dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
segs = [(src.data_ptr() + i * seg, dst.data_ptr() + i * seg, seg) for i in range(nseg)]
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6567.4
res = []
for tile_kb, stages, ctas in itertools.product((8, 16, 32, 64), (2, 3, 4, 6), (1, 2, 3, 4)):
    if 128 + stages * tile_kb * 1024 > 227 * 1024:
        continue
    if ctas * (128 + stages * tile_kb * 1024) > 227 * 1024:
        continue
    L.vgpu_pack_config(tile_kb * 1024, stages, ctas)
    for _ in range(2):
        v.pack(segs, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        v.pack(segs, st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gbs = 2 * total / ms / 1e6
    res.append({"tile_kb": tile_kb, "stages": stages, "ctas_per_sm": ctas, "smem_kb": (128 + stages * tile_kb * 1024) / 1024, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3)})
    print(res[-1], flush=True)
assert torch.equal(src, dst)
# torch's own device-to-device copy as a yardstick (the MEASURED_PEAKS method)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dst.copy_(src); torch.cuda.synchronize()
a.record()
for _ in range(5):
    dst.copy_(src)
b.record(); torch.cuda.synchronize()
ref = 2 * total / (a.elapsed_time(b) / 5) / 1e6
best = max(res, key=lambda r: r["gbs"])
json.dump({"sweep": res, "best": best, "torch_copy_gbs": round(ref, 1), "peak": peak}, open("gpurun_out/pack_sweep.json", "w"), indent=1)
print("BEST", best, "torch copy", round(ref, 1))
