"""The C-ABI library loads and exports every symbol include/vgpu.h declares; region ABI offsets; host-side region
logic (no compute calls, no GPU)."""
import ctypes as C
import os
import re

import k8s_device_plugin_b200 as v
from conftest import CORE_SO, HOOK_SO, ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "vgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vgpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_by_both_libraries():
    names = _declared()
    assert len(names) >= 35
    assert sorted(v.ABI) == names, "Python binding and header disagree"
    for so in (CORE_SO, HOOK_SO):
        lib = C.CDLL(so, mode=C.RTLD_LOCAL)
        for n in names:
            assert hasattr(lib, n), f"{n} missing from {so}"


def test_every_header_under_include_is_fully_exported():
    import glob
    for hdr in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
        names = sorted(set(re.findall(r"\b(vgpu_[a-z0-9_]+)\s*\(", text)))
        for so in (CORE_SO, HOOK_SO):
            lib = C.CDLL(so, mode=C.RTLD_LOCAL)
            missing = [n for n in names if not hasattr(lib, n)]
            assert not missing, f"{os.path.basename(hdr)}: {missing} missing from {so}"


def test_hook_library_exports_the_interposed_driver_surface():
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", HOOK_SO], stdout=subprocess.PIPE, text=True).stdout
    for n in ("dlsym", "cuInit", "cuGetProcAddress", "cuGetProcAddress_v2", "cuMemAlloc_v2", "cuMemFree_v2", "cuMemAllocManaged",
              "cuMemAllocPitch_v2", "cuMemGetInfo_v2", "cuDeviceTotalMem_v2", "cuLaunchKernel", "cuLaunchKernelEx",
              "cuLaunchCooperativeKernel", "cuDevicePrimaryCtxRetain", "nvmlDeviceGetMemoryInfo", "cuMemoryAllocate",
              "cuVGPUViewAllocator"):
        assert re.search(r" T %s$" % n, syms, flags=re.M), n
    core = subprocess.run(["nm", "-D", "--defined-only", CORE_SO], stdout=subprocess.PIPE, text=True).stdout
    assert not re.search(r" T (dlsym|cuInit|cuMemAlloc_v2)$", core, flags=re.M), "core library must not interpose"
    soname = subprocess.run(["readelf", "-d", HOOK_SO], stdout=subprocess.PIPE, text=True).stdout
    assert "soname: [libvgpu.so]" in soname        # DT_SONAME the chart/ld.so.preload contract expects


def test_hook_exports_every_symbol_the_reference_hook_exports():
    """203 cu* + 243 nvml* exported names (nm -D lib/nvidia/libvgpu.so, committed as a golden list)."""
    import subprocess
    from conftest import GOLDEN
    want = open(os.path.join(GOLDEN, "ref_exported_symbols.txt")).read().split()
    assert len(want) == 446
    have = set(l.split()[2] for l in subprocess.run(["nm", "-D", "--defined-only", HOOK_SO], stdout=subprocess.PIPE, text=True).stdout.splitlines()
               if len(l.split()) == 3)
    missing = [n for n in want if n not in have]
    assert not missing, missing[:10]


def test_passthrough_trampolines_forward_with_arguments_intact(tmp_path):
    """A directly linked program calling names that only forward (cuDriverGetVersion, nvmlDeviceGetCount_v2 ...)."""
    import subprocess
    from conftest import FAKE
    src = tmp_path / "pt.c"
    src.write_text("""#include <stdio.h>
extern int cuInit(unsigned); extern int cuDriverGetVersion(int*); extern int cuDeviceGetCount(int*);
extern int cuDeviceGetAttribute(int*, int, int); extern int nvmlInit_v2(void); extern int nvmlDeviceGetCount_v2(unsigned*);
extern int cuDeviceGetName(char*, int, int);
int main(){ int v=0,c=0,a=0; unsigned n=0; char nm[64]; cuInit(0); cuDriverGetVersion(&v); cuDeviceGetCount(&c);
 cuDeviceGetAttribute(&a, 16, 0); nvmlInit_v2(); nvmlDeviceGetCount_v2(&n); cuDeviceGetName(nm, 64, 0);
 printf("%d %d %d %u %s\\n", v, c, a, n, nm); return 0; }
""")
    exe = tmp_path / "pt"
    subprocess.run(["gcc", "-o", str(exe), str(src), "-L", FAKE, "-l:libcuda.so.1", "-Wl,--allow-shlib-undefined"], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=FAKE, LD_PRELOAD=HOOK_SO, LIBCUDA_LOG_LEVEL="0", CUDA_DEVICE_MEMORY_SHARED_CACHE=str(tmp_path / "pt.cache"))
    out = subprocess.run([str(exe)], env=env, stdout=subprocess.PIPE, text=True, check=True).stdout.strip()
    assert out == "12090 1 148 1 NVIDIA B200 (fake)"


def test_region_layout_matches_appendix_a(tmp_path):
    path = str(tmp_path / "r.cache")
    lim = [0] * 16
    lim[0] = 8 << 30
    sm = [100] * 16
    sm[0] = 30
    with v.Region(path, create=True, mem_limits=lim, sm_limits=sm, priority=0) as r:
        # the reference's 0xC4748 bytes, then (page-aligned, invisible to consumers that map the reference size) the
        # swap-counter extension block of include/vgpu_region.h
        assert v.REGION_SIZE == 0xC4748 and os.path.getsize(path) == 0xC5000 + 64 + 1024 * 64
        pid = 4242
        assert r.claim(pid) == 0
        assert r.try_add(pid, 0, 1000, v.MEM_BUFFER)
        assert r.try_add(pid, 0, 77, v.MEM_CONTEXT, enforce=False)
        raw = open(path, "rb").read()
        i32 = lambda o: int.from_bytes(raw[o:o + 4], "little", signed=True)
        u64 = lambda o: int.from_bytes(raw[o:o + 8], "little")
        assert i32(0x0) == v.REGION_MAGIC
        assert u64(0x638) == 8 << 30 and u64(0x6B8) == 30
        assert i32(0x738) == pid                       # procs[0].pid
        assert u64(0x738 + 8 + 0x00) == 77             # used[0].context_size
        assert u64(0x738 + 8 + 0x10) == 1000           # used[0].buffer_size
        assert u64(0x738 + 8 + 0x20) == 1077           # used[0].total
        assert i32(0x738 + 0x308) == 1                 # status RUNNING
        assert i32(0xC4738) == 1 and i32(0xC473C) == 1 and i32(0xC4740) == 2 and i32(0xC4744) == 0
        snap = r.snapshot()
        assert snap.initialized == 1 and snap.proc_num == 1 and snap.usage_total[0] == 1077
        # monitor feedback path (feedback.go:207-251)
        r.set_feedback(recent_kernel=-1, utilization_switch=0)
        s2 = r.snapshot()
        assert s2.recent_kernel == -1 and s2.utilization_switch == 0
        r.set_feedback(recent_kernel=2)
        assert r.snapshot().utilization_switch == 0
        # slot compaction on release (exit_handler@0x43501)
        assert r.claim(5000) == 1 and r.claim(5001) == 2
        r.try_add(5001, 0, 5, v.MEM_BUFFER, enforce=False)
        r.release(pid)
        assert r.snapshot().proc_num == 2
        assert r.proc(0).pid == 5001 and r.proc(0).used[0].total == 5
        assert r.proc(1).pid == 5000


def test_quota_is_strict_and_reclaims_dead_processes(tmp_path):
    path = str(tmp_path / "q.cache")
    lim = [0] * 16
    lim[0] = 1000
    with v.Region(path, create=True, mem_limits=lim) as r:
        me = os.getpid()
        r.claim(me)
        r.claim(2 ** 22 + 12345)                 # a pid that does not exist
        assert r.try_add(2 ** 22 + 12345, 0, 600, enforce=False)
        assert r.try_add(me, 0, 400)              # 600 + 400 == limit: admitted
        r.sub(me, 0, 400)
        # 600 (dead) + 900 > 1000 -> dead slot is reaped, then admitted
        assert r.try_add(me, 0, 900)
        assert r.snapshot().proc_num == 1 and r.usage(0) == 900
        assert not r.try_add(me, 0, 101)
        assert r.try_add(me, 0, 100)


def test_parse_limit_equals_reference_binary_vectors():
    import json
    from conftest import GOLDEN
    for text, want in json.load(open(os.path.join(GOLDEN, "ref_kat.json")))["limit"]:
        assert v.parse_limit(text) == want, text


def test_parse_limit_vectors():
    assert v.parse_limit("8192m") == 8 << 30
    assert v.parse_limit("17179869184g") == 0      # overflow -> unlimited, like the reference
    assert v.parse_limit("0x10m") == 16 << 20
    assert v.parse_limit("") == 0 and v.parse_limit(None) == 0
