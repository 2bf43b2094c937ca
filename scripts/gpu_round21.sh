#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; mkdir -p /tmp/vgpulock
cat > /tmp/sw.py <<'PY'
import json, torch
bufs = [torch.full((1536 << 20,), i, dtype=torch.uint8, device="cuda") for i in range(4)]
for rnd in range(2):
    for b in bufs:
        b.add_(1)
torch.cuda.synchronize()
print("phase1 ok", flush=True)
b = bufs[0]
print("item0", int(b[0].item()), flush=True)
print("itemN", int(b[-1].item()), flush=True)
print("sum", int(b.sum(dtype=torch.int64).item()), flush=True)
PY
( export LD_PRELOAD=$PWD/k8s-device-plugin_b200/lib/libvgpu.so CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=3072m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sw.cache LIBCUDA_LOG_LEVEL=3 VGPU_STRICT_CUDA_ERRORS=1 PYTORCH_NO_CUDA_MEMORY_CACHING=1 CUDA_LAUNCH_BLOCKING=1; timeout 60 python /tmp/sw.py ) > $O/torch_swap_dbg.log 2>&1
grep -v "gpa\]" $O/torch_swap_dbg.log | tail -40 | cut -c1-400
