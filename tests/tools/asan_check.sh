#!/bin/bash
# Builds the hook with AddressSanitizer + UBSan, then with ThreadSanitizer (into k8s-device-plugin_b200/build/{asan,tsan}/, never shipped) and drives it on
# the fake driver through the scenarios the CPU suite uses: swap (cyclic / zipf / physical pressure / virtual limit mode),
# threads + fork, randomised three-GPU traces, the limiter launch loop. Any sanitizer report fails the script.
set -u
cd "$(dirname "$0")/../.."
R=$PWD; P=$R/k8s-device-plugin_b200; A=$P/build/asan; L=/usr/lib/x86_64-linux-gnu
mkdir -p $A; ln -sf $L/libasan.so.8 $A/libasan.so; ln -sf $L/libubsan.so.1 $A/libubsan.so
FL="-std=c++17 -O1 -g -fno-omit-frame-pointer -DVGPU_NO_DLSYM_OVERRIDE -fsanitize=address,undefined -fPIC -fvisibility=hidden -I$R/include -I/usr/local/cuda/include"
for f in driver region kmod swap limiter runtime cabi plugin_core sched_core hook passthrough; do g++ $FL -c -o $A/$f.o $P/csrc/$f.cc || exit 1; done
g++ -shared -fsanitize=address,undefined -L$A -Wl,-soname,libvgpu.so -o $A/libvgpu.so $A/{driver,region,kmod,swap,limiter,runtime,cabi,plugin_core,sched_core,hook,passthrough}.o $P/build/kernels_cubin.o -ldl -lpthread || exit 1
T=$(mktemp -d); LOG=$T/san.log; mkdir -p /tmp/vgpulock
export LD_LIBRARY_PATH=$R/oracle/_ref/fake ASAN_OPTIONS=detect_leaks=0:log_path=$T/asan UBSAN_OPTIONS=print_stacktrace=1:log_path=$T/ubsan LIBCUDA_LOG_LEVEL=0 FAKE_GPU_CTX_MIB=16
PRE=$L/libasan.so.8:$L/libubsan.so.1:$A/libvgpu.so
SW="FAKE_GPU_EXEC=1 CUDA_OVERSUBSCRIBE=true VGPU_SWAP_CHUNK_MB=4 VGPU_SWAP_ARENA_GB=8 VGPU_SWAP_SLAB_MB=64 VGPU_SWAP_SPARE_MB=16"
SB="$P/lib/swap_bench --cubin $P/build/vgpu_kernels.cubin --mib 16 --warmup 8"
# usage: run VAR=value ... -- program args   (ONE env invocation: the preload must reach the program only, not a nested env)
n=0; run() { n=$((n+1)); local vars=(); while [ "$1" != "--" ]; do vars+=("$1"); shift; done; shift
  env "${vars[@]}" CUDA_DEVICE_MEMORY_SHARED_CACHE=$T/c$n.cache LD_PRELOAD=$PRE "$@" > $T/out$n.txt 2>&1; echo "  [$n] rc=$? $(tail -c 120 $T/out$n.txt | tr '\n' ' ' | cut -c1-100)"; }
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m -- $SB --buffers 32 --steps 96 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m -- $SB --buffers 32 --steps 200 --order zipf
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m VGPU_SWAP_PREFETCH_MB=0 -- $SB --buffers 24 --steps 72 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m VGPU_SWAP_HOST_BACKED=1 -- $SB --buffers 24 --steps 72 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m -- $SB --buffers 24 --steps 96 --order zipf --ro-every 2
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=384m FAKE_GPU_TOTAL_MIB=200 -- $SB --buffers 32 --steps 96 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=384m FAKE_GPU_TOTAL_MIB=200 VGPU_SWAP_LIMIT_MODE=virtual -- $SB --buffers 20 --steps 60 --order cyclic
run CUDA_DEVICE_MEMORY_LIMIT_0=64m -- $R/oracle/_ref/hook_stress threads 8 2000
run CUDA_DEVICE_MEMORY_LIMIT_0=64m -- $R/oracle/_ref/hook_stress fork
run FAKE_GPU_EXEC=1 CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force -- $P/lib/launch_loop $P/build/vgpu_kernels.cubin 8 2
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=128m -- $R/oracle/_ref/hook_stress swap 4 120
python - "$T" <<'PY'
import os, sys
sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
src = open("tests/tools/fuzz_vs_reference.py").read().split("import tempfile")[0]
ns = {"__file__": os.path.abspath("tests/tools/fuzz_vs_reference.py")}
exec(compile(src, "fuzz", "exec"), ns)
for seed in (1, 2):
    open(os.path.join(sys.argv[1], f"fz{seed}.txt"), "w").write("\n".join(ns["gen"](seed)) + "\n")
PY
for s in 1 2; do run FAKE_GPU_COUNT=3 CUDA_DEVICE_MEMORY_LIMIT_0=96m CUDA_DEVICE_MEMORY_LIMIT_1=64m CUDA_DEVICE_MEMORY_LIMIT_2=200m -- $R/oracle/_ref/trace_replay $T/fz$s.txt; done
for s in 1 2; do run $SW FAKE_GPU_COUNT=3 VGPU_SWAP_LIMIT_MODE=virtual CUDA_DEVICE_MEMORY_LIMIT_0=196m CUDA_DEVICE_MEMORY_LIMIT_1=164m CUDA_DEVICE_MEMORY_LIMIT_2=300m -- $R/oracle/_ref/trace_replay $T/fz$s.txt; done
if ls $T/asan.* $T/ubsan.* > /dev/null 2>&1; then echo "SANITIZER REPORTS:"; head -60 $T/asan.* $T/ubsan.* 2>/dev/null; exit 1; fi
echo "no ASan/UBSan reports in $n scenarios"
# ---- ThreadSanitizer: the multi-threaded scenarios (application threads racing on the allocation table; the pager and the pool-grower threads)
TS=$P/build/tsan; mkdir -p $TS; ln -sf $L/libtsan.so.2 $TS/libtsan.so
FT="-std=c++17 -O1 -g -fno-omit-frame-pointer -DVGPU_NO_DLSYM_OVERRIDE -fsanitize=thread -fPIC -fvisibility=hidden -I$R/include -I/usr/local/cuda/include"
for f in driver region kmod swap limiter runtime cabi plugin_core sched_core hook passthrough; do g++ $FT -c -o $TS/$f.o $P/csrc/$f.cc || exit 1; done
g++ -shared -fsanitize=thread -L$TS -Wl,-soname,libvgpu.so -o $TS/libvgpu.so $TS/{driver,region,kmod,swap,limiter,runtime,cabi,plugin_core,sched_core,hook,passthrough}.o $P/build/kernels_cubin.o -ldl -lpthread || exit 1
PRE=$L/libtsan.so.2:$TS/libvgpu.so
export TSAN_OPTIONS="halt_on_error=0 log_path=$T/tsan"
run CUDA_DEVICE_MEMORY_LIMIT_0=64m -- $R/oracle/_ref/hook_stress threads 8 1500
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m -- $SB --buffers 24 --steps 96 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m VGPU_SWAP_PREFETCH_MB=0 -- $SB --buffers 24 --steps 72 --order cyclic
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=256m VGPU_SWAP_HOST_BACKED=1 -- $SB --buffers 24 --steps 72 --order zipf
run FAKE_GPU_EXEC=1 CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_LIMIT_0=1g -- $P/lib/launch_loop $P/build/vgpu_kernels.cubin 8 2
run $SW CUDA_DEVICE_MEMORY_LIMIT_0=128m -- $R/oracle/_ref/hook_stress swap 4 80
# the C-ABI engine driven from three Python threads with shared operands (host-backed: operands used in place; default: demands
# that wait for other threads' pins) — same objects linked as libvgpu_core.so, loaded by the package through VGPU_CORE_SO
g++ -shared -fsanitize=thread -L$TS -Wl,-soname,libvgpu_core.so -o $TS/libvgpu_core.so $TS/{driver,region,kmod,swap,limiter,runtime,cabi,plugin_core,sched_core,hook,passthrough}.o $P/build/kernels_cubin.o -ldl -lpthread || exit 1
python - > $T/threaded.py <<'PY'
import sys
sys.path.insert(0, "tests")
import test_engine_on_functional_fake as t
print(t.THREADED_OPERANDS_SCRIPT + "\nsw.close()\n")
PY
for hb in 1 0; do n=$((n+1)); env FAKE_GPU_EXEC=1 VGPU_SWAP_CHUNK_MB=4 VGPU_SWAP_ARENA_GB=8 VGPU_SWAP_SLAB_MB=64 VGPU_SWAP_SPARE_MB=16 VGPU_SWAP_HOST_BACKED=$hb VGPU_ROOT=$R \
  VGPU_CORE_SO=$TS/libvgpu_core.so CUDA_DEVICE_MEMORY_SHARED_CACHE=$T/c$n.cache LD_PRELOAD=$L/libtsan.so.2 python $T/threaded.py > $T/out$n.txt 2>&1
  echo "  [$n] rc=$? $(tail -c 120 $T/out$n.txt | tr '\n' ' ' | cut -c1-100)"; done
if ls $T/tsan.* > /dev/null 2>&1; then echo "TSAN REPORTS:"; head -80 $T/tsan.*; exit 1; fi
echo "no TSan reports"; rm -rf $T
