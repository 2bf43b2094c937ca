"""Victim-scan kernel time vs table size (isolated): 32 B/row algorithmic traffic against the HBM roofline."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import k8s_device_plugin_b200 as v

torch.zeros(1, device="cuda")
v.lib()
out = []
for n in (1 << 10, 1 << 13, 1 << 16, 1 << 18, 1 << 20, 1 << 22):
    rng = np.random.default_rng(n)
    arr = np.zeros((n, 4), dtype=np.uint64)
    arr[:, 1] = rng.integers(1, 1 << 22, size=n)
    arr[:, 2] = rng.integers(0, 1 << 24, size=n)
    arr[:, 3] = 1
    d = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).cuda()
    need = int(arr[:, 1].sum() // 4)
    for _ in range(3):
        v.victim_scan(d.data_ptr(), n, need, (1 << 24) - 1)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        got, freed, ins = v.victim_scan(d.data_ptr(), n, need, (1 << 24) - 1)
    us = (time.perf_counter() - t0) / reps * 1e6
    out.append({"rows": n, "wall_us_per_scan_incl_readback": round(us, 1), "victims": len(got), "algorithmic_GBps": round(32 * n / us / 1e3, 1)})
    print(out[-1], flush=True)
json.dump(out, open("gpurun_out/scan_scaling.json", "w"), indent=1)
