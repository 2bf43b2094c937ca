"""Small engine-only swap loop for ncu (never a bench number): 32 x 64 MiB under a 1 GiB quota, 48 cyclic touches."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import k8s_device_plugin_b200 as v

torch.zeros(1, device="cuda")
L = v.lib()
st = torch.cuda.current_stream().cuda_stream
sw = v.Swap(resident_cap=1 << 30)
n, nbytes = 32, 64 << 20
bufs = [sw.alloc(nbytes) for _ in range(n)]
for i, p in enumerate(bufs):
    sw.acquire([p], st); L.vgpu_wl_fill(p, nbytes // 8, i, C.c_void_p(st)); sw.release([p], st)
for t in range(int(os.environ.get("TOUCHES", "48"))):
    p = bufs[t % n]
    sw.acquire([p], st); L.vgpu_wl_touch(p, nbytes // 8, C.c_void_p(st)); sw.release([p], st)
torch.cuda.synchronize()
sw.drain()
print(sw.stats())
