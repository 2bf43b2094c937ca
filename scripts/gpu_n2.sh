#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $O/n2_gpus.txt 2>&1
for g in 0 1; do cat /sys/bus/pci/devices/$(nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader | sed 's/^0000//' | tr 'A-F' 'a-f')/numa_node >> $O/n2_gpus.txt 2>&1; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"
cat $O/n2_gpus.txt; grep "^{" $O/bench_n2.json; tail -5 $O/bench_n2.err
