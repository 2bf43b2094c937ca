#!/bin/bash
# scripts/ab_r1.sh — A/B on ONE box: the round-1 engine (baseline_r1/, built from the round-1 commit) against the current
# one, same unmodified app, same workload, alternating. Usage: scripts/ab_r1.sh <buffers> <steps> <reps> ["NAME ENV=VAL ..." ...]
nbuf=${1:-512}; steps=${2:-1024}; reps=${3:-2}; shift 3 || true
P=k8s-device-plugin_b200
run() {  # name lib bench cubin extra-env...
    name=$1; lib=$2; bench=$3; cubin=$4; shift 4
    rm -f /tmp/ab.cache
    env LD_PRELOAD=$PWD/$lib CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ab.cache LIBCUDA_LOG_LEVEL=1 "$@" \
        timeout 600 $bench --cubin $cubin --buffers $nbuf --mib 64 --steps $steps --warmup 64 --order cyclic 2>/dev/null | tail -1 | python3 -c '
import json,sys
r=json.loads(sys.stdin.read()); gb=(r["page_in_bytes"]+r["page_out_bytes"])/1e9
p=r.get("pager_ms", {})
print("%-14s GB/s=%5.1f mism=%s"%(sys.argv[1], gb/(r["event_ms"]/1e3), r["mismatches"]), "app_vmm_ms=%.0f"%r["host_ms"]["vmm"], "pager busy=%.0f vmm=%.0f steps=%s"%(p.get("busy",0), p.get("vmm",0), p.get("steps")), "calls", r.get("vmm_calls"), "slow", r.get("vmm_slow"), "pf", r.get("prefetch"))' $name
}
$P/lib/linkbench 1024 | python3 -c 'import json,sys; r=json.loads(sys.stdin.read()); print("link", {k: r[k] for k in ("h2d_gbs","d2h_gbs","bidir_gbs")})'
for i in $(seq 1 $reps); do
    run r1 baseline_r1/libvgpu.so baseline_r1/swap_bench baseline_r1/vgpu_kernels.cubin
    run r2 $P/lib/libvgpu.so $P/lib/swap_bench $P/build/vgpu_kernels.cubin
    for spec in "$@"; do name=${spec%% *}; envs=${spec#"$name"}; run $name $P/lib/libvgpu.so $P/lib/swap_bench $P/build/vgpu_kernels.cubin $envs; done
done
