#!/bin/bash
# scripts/gpu.sh — the ONE lease script: runs a command on a B200 box through gpurun, retrying while the pod is busy
# (exit code 3 = nothing charged), and leaves the tail of the output in gpurun_out/<name>.log.
#   scripts/gpu.sh <name> <timeout-seconds> [--gpus N] -- '<command run from the repo root on the GPU box>'
set -u
name=$1; tmo=$2; shift 2
gpus=()
if [ "${1:-}" = "--gpus" ]; then gpus=(--gpus "$2"); shift 2; fi
[ "${1:-}" = "--" ] && shift
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$tmo" "${gpus[@]}" -- "$@" > "gpurun_out/$name.log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then echo "[gpu.sh] $name finished rc=$rc after $attempt attempt(s)"; exit $rc; fi
    sleep 90
done
echo "[gpu.sh] $name: pod stayed busy"; exit 3
