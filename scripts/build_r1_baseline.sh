#!/bin/bash
# scripts/build_r1_baseline.sh — rebuilds the ROUND-1 engine (commit c3991e4) into the git-ignored baseline_r1/ so that
# scripts/ab_r1.sh can run it against the current one on the same box (the directory travels with gpurun).
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/r1wt
git worktree add -q /tmp/r1wt c3991e4
make -C /tmp/r1wt/k8s-device-plugin_b200/csrc -j8 > /dev/null
make -C /tmp/r1wt/k8s-device-plugin_b200/tools > /dev/null
mkdir -p baseline_r1
cp /tmp/r1wt/k8s-device-plugin_b200/lib/libvgpu.so /tmp/r1wt/k8s-device-plugin_b200/lib/swap_bench /tmp/r1wt/k8s-device-plugin_b200/build/vgpu_kernels.cubin baseline_r1/
git worktree remove --force /tmp/r1wt
ls -la baseline_r1
