#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.log
export PATH=/usr/local/cuda/bin:$PATH
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
printf 'A 0 1048576\nI\nA 1 4194304\nF 0\nX 0x1234\nF 1\nT\nL 1 1 1\n' > /tmp/t8.txt
# --- why does the reference binary hang on this driver? debug log + backtrace
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ref_real.cache LIBCUDA_LOG_LEVEL=4 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; oracle/_ref/trace_replay /tmp/t8.txt > $O/ref_real.out 2> $O/ref_real.err.full & echo $! > /tmp/ref.pid )
sleep 12
RP=$(cat /tmp/ref.pid)
if kill -0 $RP 2>/dev/null; then
  timeout 60 cuda-gdb -p $RP -batch -ex "thread apply all bt 25" > $O/ref_bt.txt 2>&1
  kill -9 $RP
  echo "reference still running after 12 s (hung)" > $O/ref_status.txt
else echo "reference finished" > $O/ref_status.txt; fi
grep -v "LOADING\|loading\|can't find\|find_symbols\|into dlsym" $O/ref_real.err.full | tail -40 > $O/ref_real.err.tail; rm -f $O/ref_real.err.full
# --- tests
timeout 1500 python -m pytest tests -m gpu -q --timeout 420 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/status.txt
# --- swap through the hook: 8 GiB quota, 24 GiB set
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 1 ) > $O/swap_small.json 2> $O/swap_small.err; echo "swap rc=$?" >> $O/status.txt
# --- bench (short)
timeout 900 python bench.py --steps 24 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
# --- ncu: launch list + one full capture of the pack kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python scripts/ncu_target.py > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/status.txt
TOUCHES=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vgpu_pack_tma -s 40 -c 3 -o $O/prof_pack python scripts/ncu_target.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/status.txt
cat $O/status.txt $O/ref_status.txt; tail -30 $O/pytest_gpu.log; cat $O/swap_small.json; tail -3 $O/swap_small.err; cat $O/bench.json; tail -5 $O/bench.err; tail -3 $O/ncu_list.log; tail -3 $O/ncu_full.log
