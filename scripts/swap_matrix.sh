#!/bin/bash
# scripts/swap_matrix.sh — the unmodified driver-API app (tools/swap_bench.c) under LD_PRELOAD=libvgpu.so over a list of
# engine settings: one JSON line per variant in gpurun_out/swap_matrix.jsonl. Usage (on the GPU box, from the repo root):
#   scripts/swap_matrix.sh [buffers] [steps] [order] -- "NAME ENV=VAL ENV=VAL" "NAME2 ..." ...
set -u
nbuf=${1:-1152}; steps=${2:-1152}; order=${3:-cyclic}; shift 3 || true
[ "${1:-}" = "--" ] && shift
P=k8s-device-plugin_b200
out=gpurun_out/swap_matrix.jsonl
mkdir -p gpurun_out
for spec in "$@"; do
    name=${spec%% *}; envs=${spec#"$name"}
    rm -f /tmp/sm.cache
    line=$(env LD_PRELOAD=$PWD/$P/lib/libvgpu.so CUDA_DEVICE_MEMORY_LIMIT_0=${QUOTA_MIB:-8192}m CUDA_OVERSUBSCRIBE=true \
        CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sm.cache LIBCUDA_LOG_LEVEL=1 $envs \
        timeout 600 $P/lib/swap_bench --cubin $P/build/vgpu_kernels.cubin --buffers $nbuf --mib 64 --steps $steps --warmup 64 --order $order ${EXTRA_ARGS:-} 2> gpurun_out/swap_matrix_$name.err | tail -1)
    echo "{\"variant\": \"$name\", \"env\": \"$envs\", \"result\": ${line:-null}}" | tee -a $out | python3 -c '
import json,sys
d=json.loads(sys.stdin.read()); r=d["result"] or {}
if r and "event_ms" in r:
    gb=(r["page_in_bytes"]+r["page_out_bytes"])/1e9
    print(d["variant"], "GB/s=%.1f"%(gb/(r["event_ms"]/1e3)), "mism=%s"%r["mismatches"], "host", r["host_ms"], "pager", r["pager_ms"], "vmm_calls", r["vmm_calls"], "slow", r.get("vmm_slow"), "prefetch", r["prefetch"], "direct_in_GB=%.1f out=%.1f"%(r["direct_in_bytes"]/1e9, r["direct_out_bytes"]/1e9), "waits", r["demand_waits"], "pack", r["pack_launches"], "fill_ms", r["alloc_fill_ms"])
else: print(d["variant"], "FAILED", r)
'
    tail -3 gpurun_out/swap_matrix_$name.err
done
