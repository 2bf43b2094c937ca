#!/bin/bash
# First contact with a B200 box: environment probe, link/VMM microbench, reference binary on the real driver,
# smoke, GPU tests, a small swap run. Everything is logged under gpurun_out/; failures do not stop later steps.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
export PATH=/usr/local/cuda/bin:$PATH
{
  echo "== nvidia-smi"; nvidia-smi | head -25
  echo "== cpu"; nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2; ulimit -l
  echo "== libs"; ldconfig -p | grep -E "libcuda.so|libnvidia-ml.so" ; ls /dev/nvidia* 2>/dev/null | head
  echo "== numa"; lscpu | grep -i numa | head -5
} > $O/probe.txt 2>&1
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
timeout 300 $L/linkbench 1024 > $O/linkbench.json 2> $O/linkbench.err; echo "linkbench rc=$?" >> $O/probe.txt
timeout 120 $L/intercept_bench $CUBIN 20000 200000 > $O/intercept_bare.json 2>&1
mkdir -p /tmp/vgpulock
cat > /tmp/t8.txt <<T
A 0 1048576
I
A 1 4194304
F 0
I
X 0x1234
F 1
T
L 1 1 1
T
echo "== reference binary on real driver" >> $O/probe.txt
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ref_real.cache LIBCUDA_LOG_LEVEL=3 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 120 oracle/_ref/trace_replay /tmp/t8.txt ) > $O/ref_real.out 2> $O/ref_real.err; echo "ref rc=$?" >> $O/probe.txt
tail -c 3000 $O/ref_real.err > $O/ref_real.err.tail; rm -f $O/ref_real.err
echo "== new hook on real driver" >> $O/probe.txt
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/new_real.cache LIBCUDA_LOG_LEVEL=3 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 120 oracle/_ref/trace_replay /tmp/t8.txt ) > $O/new_real.out 2> $O/new_real.err; echo "new rc=$?" >> $O/probe.txt
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/new_real2.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 120 $L/intercept_bench $CUBIN 20000 200000 ) > $O/intercept_new.json 2>&1
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ref_real2.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 300 $L/intercept_bench $CUBIN 2000 20000 ) > $O/intercept_ref.json 2>/dev/null
echo "== smoke" >> $O/probe.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/probe.txt
echo "== swap bench small (new hook, 8 GiB quota, 24 GiB set)" >> $O/probe.txt
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 600 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 1 ) > $O/swap_small.json 2> $O/swap_small.err; echo "swap rc=$?" >> $O/probe.txt
cat $O/probe.txt; cat $O/linkbench.json; cat $O/intercept_*.json; tail -5 $O/smoke.log; tail -15 $O/pytest_gpu.log; cat $O/swap_small.json; tail -5 $O/swap_small.err
