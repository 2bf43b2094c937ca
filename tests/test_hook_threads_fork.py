"""Threading and fork behaviour of the preloaded hook on the fake driver (SURVEY.md §8b "Threading"): concurrent
alloc/free from many threads leaves nothing charged, a forked child gets its own slot, and the bytes of a child that
died without freeing are reclaimed before a request is refused (rm_quitted_process). CPU only."""
import json
import os
import subprocess
import sys

from conftest import FAKE, HOOK_SO, OREF, REF_SO, SHIM_SO, have_reference


def _run(tmp_path, args, preload=HOOK_SO, timeout=300):
    env = dict(os.environ)
    env.update({"CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "hs.cache"), "CUDA_DEVICE_MEMORY_LIMIT_0": "64m", "FAKE_GPU_CTX_MIB": "16",
                "LIBCUDA_LOG_LEVEL": "0", "LD_LIBRARY_PATH": FAKE + ":" + env.get("LD_LIBRARY_PATH", ""), "LD_PRELOAD": preload})
    r = subprocess.run([os.path.join(OREF, "hook_stress")] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_concurrent_threads_leave_nothing_charged(tmp_path):
    (out,) = _run(tmp_path, ["threads", "16", "4000"])
    M = 1 << 20
    assert out == {"mode": "threads", "free_errors": 0, "ctx": 16 * M, "buf": 0, "tot": 16 * M, "procs": 1, "driver_bytes": 0}


def test_forked_child_gets_a_slot_and_its_orphaned_bytes_are_reclaimed(tmp_path):
    child, parent = _run(tmp_path, ["fork"])
    M = 1 << 20
    assert child == {"mode": "child", "rc": 0, "buf": 24 * M, "procs": 2}           # parent's 8 MiB + its own 16 MiB, two slots
    assert parent["r0"] == 0 and parent["buf_after_child"] == 24 * M and parent["procs_after_child"] == 2
    # 16 ctx + 8 + 16 (orphan) + 32 > 64: the orphan is reclaimed, the request fits, the dead slot is gone
    assert parent["r1"] == 0 and parent["buf_after_reclaim"] == 40 * M and parent["procs_after_reclaim"] == 1
    assert parent["r2"] == -1                                                        # 16 + 8 + 32 + 16 > 64: refused
    if have_reference():
        os.makedirs("/tmp/vgpulock", exist_ok=True)
        ref_dir = tmp_path / "ref"; ref_dir.mkdir()
        rchild, rparent = _run(ref_dir, ["fork"], preload=SHIM_SO + ":" + REF_SO)
        # same story in the reference: "rm pid=<child>" (rm_quitted_process@0x41a8e), then the request fits
        assert rchild == child and rparent == parent


def test_preload_is_inert_in_processes_that_never_touch_cuda(tmp_path):
    """/etc/ld.so.preload puts the hook into EVERY process of the container (server.go:386-391): shells, coreutils,
    interpreters. With no driver library on the box at all, nothing may happen: no output, no region file, exit codes
    intact, children and forks fine."""
    import subprocess
    from conftest import HOOK_SO
    cache = tmp_path / "never.cache"
    env = {"PATH": os.environ.get("PATH", "/usr/bin:/bin"), "LD_PRELOAD": HOOK_SO, "CUDA_DEVICE_MEMORY_LIMIT_0": "1g", "CUDA_DEVICE_SM_LIMIT": "30",
           "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(cache)}
    r = subprocess.run(["/bin/sh", "-c", "ls / > /dev/null && echo one | cat && (exit 7)"], env=env, capture_output=True, text=True)
    assert (r.returncode, r.stdout, r.stderr) == (7, "one\n", "")
    code = "import os,subprocess; print(subprocess.run(['echo','kid'],capture_output=True,text=True).stdout.strip()); pid=os.fork(); os._exit(0) if pid==0 else print(os.waitpid(pid,0)[1])"
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert (r.returncode, r.stdout.split(), r.stderr) == (0, ["kid", "0"], "")
    assert not cache.exists()
