"""vgpu_pack_tma launches for `ncu --set full` (never a bench number): two launches of one 32 MiB chunk (the engine's
launch shape) and two of 1 GiB (large enough that the writes cannot hide in the 126 MB L2), through the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import k8s_device_plugin_b200 as v

n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, device="cuda").fill_(5)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
v.lib()
for size in (32 << 20, 32 << 20, n, n):
    v.pack([(src.data_ptr(), dst.data_ptr(), size)], st)
    torch.cuda.synchronize()
assert torch.equal(src, dst)
print("ok")
