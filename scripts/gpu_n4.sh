#!/bin/bash
# where does the 8-replica run lose per-GPU throughput? (1) VMM remap latency vs number of processes remapping on other
# GPUs, (2) concurrent pinned-DMA bandwidth per GPU, (3) the bench at N with the host-time breakdown.
cd "$(dirname "$0")/.."
N=${1:-4}
O=gpurun_out; mkdir -p $O
L=k8s-device-plugin_b200/lib
: > $O/vmm_contend.txt
for k in 1 2 $N; do
  echo "--- $k concurrent remappers" >> $O/vmm_contend.txt
  for g in $(seq 0 $((k-1))); do $L/vmm_contend $g 3 >> $O/vmm_contend.txt 2>&1 & done; wait
done
for k in 1 $N; do
  echo "--- $k concurrent linkbench" >> $O/vmm_contend.txt
  for g in $(seq 0 $((k-1))); do CUDA_VISIBLE_DEVICES=$g $L/linkbench 2>&1 | tail -1 >> $O/vmm_contend.txt & done; wait
done
cat $O/vmm_contend.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 24 --warmup 3 > $O/bench_n${N}.json 2> $O/bench_n${N}.err; echo "n$N rc=$?"
grep "^{" $O/bench_n${N}.json; tail -3 $O/bench_n${N}.err
