"""Victim-scan time vs table size through the C ABI (one cached scanner per context): 32 B/row algorithmic traffic
against the HBM roofline. Times only the native call (launches + the header/index readback it synchronises on), with the
table flushed out of L2 between scans; kernel-only durations come from the ncu launch list of the same script
(profiles/r01_scan_launches_1m.csv)."""
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
L = v.lib()
out = []
sizes = [int(x) for x in sys.argv[1:]] or [1 << 10, 1 << 13, 1 << 16, 1 << 18, 1 << 20, 1 << 22]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in sizes:
    rng = np.random.default_rng(n)
    arr = np.zeros((n, 4), dtype=np.uint64)
    arr[:, 1] = rng.integers(1, 1 << 22, size=n)
    arr[:, 2] = rng.integers(0, 1 << 24, size=n)
    arr[:, 3] = 1
    d = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).cuda()
    need = int(arr[:, 1].sum() // 64)                 # ~1.6 % of the table is evicted: a realistic admission, not a quarter of it
    idx = (C.c_uint32 * n)()
    cnt, freed, ins = C.c_uint32(0), C.c_uint64(0), C.c_int(0)

    def scan():
        rc = L.vgpu_victim_scan(d.data_ptr(), n, need, (1 << 24) - 1, C.c_void_p(0), idx, n, C.byref(cnt), C.byref(freed), C.byref(ins))
        assert rc == 0, rc

    for _ in range(3):
        scan()
    reps, us = 20, 0.0
    for _ in range(reps):
        flush.fill_(1)                                # table out of L2 between scans (126 MB L2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scan()
        us += (time.perf_counter() - t0) * 1e6
    us /= reps
    warm = 0.0
    for _ in range(reps):                             # warm: the table stays in L2 (32 MB << 126 MB), as between two real admissions
        t0 = time.perf_counter()
        scan()
        warm += (time.perf_counter() - t0) * 1e6
    warm /= reps
    out.append({"rows": n, "call_us": round(us, 1), "call_us_warm": round(warm, 1), "victims": cnt.value, "algorithmic_GBps_of_call": round(32 * n / us / 1e3, 1),
                "path": "small" if n <= 8192 else ("persist" if n <= 148 * 8192 and not os.environ.get("VGPU_SCAN_MULTILAUNCH") else "multi-launch")})
    print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/scan_scaling.json", "w"), indent=1)
