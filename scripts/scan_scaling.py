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
sizes = [int(x) for x in sys.argv[1:]] or [1 << 10, 1 << 13, 1 << 16, 1 << 18, 1 << 20, 1 << 22]
for n in sizes:
    rng = np.random.default_rng(n)
    arr = np.zeros((n, 4), dtype=np.uint64)
    arr[:, 1] = rng.integers(1, 1 << 22, size=n)
    arr[:, 2] = rng.integers(0, 1 << 24, size=n)
    arr[:, 3] = 1
    d = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).cuda()
    need = int(arr[:, 1].sum() // 4)
    for _ in range(3):
        v.victim_scan(d.data_ptr(), n, need, (1 << 24) - 1)
    reps = 20
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    us = 0.0
    for _ in range(reps):
        flush.fill_(1)                      # table out of L2 between scans (126 MB L2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got, freed, ins = v.victim_scan(d.data_ptr(), n, need, (1 << 24) - 1)
        us += (time.perf_counter() - t0) * 1e6
    us /= reps
    out.append({"rows": n, "wall_us_per_scan_incl_readback": round(us, 1), "victims": len(got), "algorithmic_GBps": round(32 * n / us / 1e3, 1)})
    print(out[-1], flush=True)
json.dump(out, open("gpurun_out/scan_scaling.json", "w"), indent=1)
