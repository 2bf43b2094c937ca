#!/bin/bash
# BASELINE.json configs[4]: 8 replicas, one per GPU (no data-path collective), plus the victim-scan scaling table.
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $O/n${N}_gpus.txt 2>&1
for g in $(seq 0 $((N-1))); do cat /sys/bus/pci/devices/$(nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader | sed 's/^0000//' | tr 'A-F' 'a-f')/numa_node >> $O/n${N}_gpus.txt 2>&1; done
grep -E "MemTotal|MemAvailable" /proc/meminfo >> $O/n${N}_gpus.txt; nproc >> $O/n${N}_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 24 --warmup 3 > $O/bench_n${N}.json 2> $O/bench_n${N}.err; echo "n$N rc=$?"
cat $O/n${N}_gpus.txt; grep "^{" $O/bench_n${N}.json; tail -5 $O/bench_n${N}.err
timeout 200 python scripts/scan_scaling.py > $O/scan_scaling.log 2>&1; tail -8 $O/scan_scaling.log
