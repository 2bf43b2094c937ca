#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
cat > /tmp/g.py <<'PY'
import torch
x = torch.zeros(1 << 20, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x.add_(1.0)
    torch.cuda.current_stream().synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        x.add_(1.0)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("ok", float(x[0]))
PY
( export LD_PRELOAD=$PWD/k8s-device-plugin_b200/lib/libvgpu.so VGPU_TRACE_GPA=1 CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/gp.cache LIBCUDA_LOG_LEVEL=0 VGPU_PRINT_STATS=1; timeout 120 python /tmp/g.py ) > $O/gpa_trace.log 2>&1
grep -c "gpa\]" $O/gpa_trace.log; grep -i "graph\|Launch" $O/gpa_trace.log | head -40; grep -c HOOKED $O/gpa_trace.log; tail -4 $O/gpa_trace.log
