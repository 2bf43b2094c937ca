"""Host-side consumer of the shared region: the reference's monitor goldens
(cmd/vGPUmonitor/pathmonitor_test.go:10-45), its feedback rules (feedback.go:165-255) and its metric surface
(metrics.go:66-97), driven against regions WRITTEN BY THE HOOK running on the fake driver."""
import os
import subprocess
import time
import urllib.request

import k8s_device_plugin_b200 as v
from conftest import FAKE, HOOK_SO, OREF
from k8s_device_plugin_b200.plugin import monitor as M


def test_is_valid_pod_golden_from_reference_test():
    assert M.is_valid_pod("123", ["123", "456"]) is True       # pathmonitor_test.go:28-31
    assert M.is_valid_pod("789", ["123", "456"]) is False      # :32-35
    assert M.is_valid_pod("123_main", ["123"]) is True         # directories are "<podUID>_<ctr>"


def _region(tmp_path, name, uuid, priority, recent_kernel=2):
    d = tmp_path / "containers" / name
    d.mkdir(parents=True)
    r = v.Region(str(d / "x.cache"), create=True, mem_limits=[8 << 30] + [0] * 15, sm_limits=[30] + [100] * 15, priority=priority)
    r.set_uuid(0, uuid)
    r.set_feedback(recent_kernel=recent_kernel)
    return r


def test_observe_blocks_low_priority_and_switches_limiter_only_under_contention(tmp_path):
    gpu = "GPU-aaaa-bbbb"
    hi = _region(tmp_path, "u1_a", gpu, priority=0)
    lo = _region(tmp_path, "u2_b", gpu, priority=1)
    alone = _region(tmp_path, "u3_c", "GPU-other", priority=1)
    v.monitor_observe([hi, lo, alone])
    s_hi, s_lo, s_al = hi.snapshot(), lo.snapshot(), alone.snapshot()
    assert s_hi.recent_kernel == 1 and s_lo.recent_kernel == -1          # high-priority task active -> low one blocked
    assert s_hi.utilization_switch == 0                                   # nobody above it, alone in its class
    assert s_lo.utilization_switch == 1                                   # contended by a higher class
    assert s_al.recent_kernel == 1 and s_al.utilization_switch == 0       # a lone task may exceed its SM limit (work conserving)
    # high-priority task goes idle: its counter decays to 0, the low one is released
    v.monitor_observe([hi, lo, alone])
    assert hi.snapshot().recent_kernel == 0
    assert lo.snapshot().recent_kernel == 0 and lo.snapshot().utilization_switch == 0
    # two tasks of the same class on one GPU -> both limited, none blocked
    a = _region(tmp_path, "u4_d", "GPU-shared", priority=1)
    b = _region(tmp_path, "u5_e", "GPU-shared", priority=1)
    v.monitor_observe([a, b])
    assert a.snapshot().utilization_switch == 1 and b.snapshot().utilization_switch == 1
    assert a.snapshot().recent_kernel == 1 and b.snapshot().recent_kernel == 1
    for r in (hi, lo, alone, a, b):
        r.close()


def test_monitor_reads_a_region_written_by_the_hook_and_exports_reference_metric_names(tmp_path):
    uid, ctr = "pod-uid-42", "main"
    cdir = tmp_path / "containers" / f"{uid}_{ctr}"
    cdir.mkdir(parents=True)
    cache = str(cdir / "abcd.cache")
    trace = tmp_path / "t.txt"
    trace.write_text("A 0 %d\nA 1 %d\nL 1 1 1\nS 3000\n" % (100 << 20, 28 << 20))
    env = dict(os.environ, LD_LIBRARY_PATH=FAKE, LD_PRELOAD=HOOK_SO, LIBCUDA_LOG_LEVEL="0", CUDA_DEVICE_MEMORY_LIMIT_0="1024m",
               CUDA_DEVICE_SM_LIMIT="30", CUDA_DEVICE_MEMORY_SHARED_CACHE=cache, FAKE_GPU_CTX_MIB="64")
    p = subprocess.Popen([os.path.join(OREF, "trace_replay"), str(trace)], env=env, stdout=subprocess.DEVNULL)
    try:
        mon = M.Monitor(str(tmp_path / "containers"), lambda: [M.PodInfo(uid, "default", "trainer-0", [ctr])])
        for _ in range(100):
            time.sleep(0.05)
            text = mon.collect()
            if "vGPU_device_memory_usage_in_bytes{" in text and f" {float((64 + 128) << 20)}" in text:
                break
        want_labels = f'podnamespace="default",podname="trainer-0",ctrname="{ctr}",vdeviceid="0",deviceuuid="GPU-'
        assert f"vGPU_device_memory_usage_in_bytes{{{want_labels}" in text
        assert f" {float((64 + 128) << 20)}" in text                       # context 64 MiB + 128 MiB of buffers
        assert f"vGPU_device_memory_limit_in_bytes{{{want_labels}" in text and f" {float(1 << 30)}" in text
        assert f'context="{64 << 20}",module="0",data="{128 << 20}",offset="0"' in text
        # feedback: the hook set recentKernel=2 on its launch; one Observe pass decrements it, limiter switch goes off (alone)
        mon.observe()
        with v.Region(cache) as r:
            s = r.snapshot()
            assert s.recent_kernel == 1 and s.utilization_switch == 0 and s.sm_limit[0] == 30
        srv = mon.serve(port=0)
        body = urllib.request.urlopen(f"http://127.0.0.1:{srv.server_address[1]}/metrics", timeout=5).read().decode()
        assert "Device_memory_desc_of_container{" in body
        srv.shutdown()
    finally:
        p.kill()
        p.wait()


def test_stale_container_dirs_are_collected_after_300_seconds(tmp_path):
    d = tmp_path / "containers" / "gone-uid_main"
    d.mkdir(parents=True)
    v.Region(str(d / "z.cache"), create=True).close()
    mon = M.Monitor(str(tmp_path / "containers"), lambda: [])
    mon.monitor_path()
    assert d.exists()                                   # younger than 300 s: kept
    mon.monitor_path(now=time.time() + 301)
    assert not d.exists()
    three = tmp_path / "containers" / "u_x"
    three.mkdir()
    for n in ("a.cache", "b.cache", "c.txt"):
        (three / n).write_text("")
    import pytest
    with pytest.raises(ValueError):
        M.check_files(str(three))                       # "cache num not matched"


def test_swap_counter_extension_block_layout_and_export(tmp_path):
    """The extension block sits page-aligned BEHIND the reference's 0xC4748-byte region in the same file: reference
    offsets, the file name and the directory contract are untouched; the monitor sums the records of a device."""
    import mmap
    import struct
    uid, ctr = "pod-uid-7", "main"
    cdir = tmp_path / "containers" / f"{uid}_{ctr}"
    cdir.mkdir(parents=True)
    path = str(cdir / "x.cache")
    r = v.Region(path, create=True, mem_limits=[8 << 30] + [0] * 15, sm_limits=[100] * 16, priority=1)
    assert os.path.getsize(path) == 0xC5000 + 64 + 1024 * 64
    assert r.swap_counters(0) == {"page_out_bytes": 0, "page_in_bytes": 0, "evictions": 0, "faults": 0, "resident_bytes": 0,
                                  "live_bytes": 0, "host_bytes": 0, "processes": 0}
    slot = r.claim(os.getpid())
    assert slot == 0
    with open(path, "r+b") as f:                                    # what two hooked processes of the container would publish
        m = mmap.mmap(f.fileno(), 0)
        assert struct.unpack_from("<II", m, 0xC5000) == (0x30303242, 1)
        m[0xC5000 + 64:0xC5000 + 128] = struct.pack("<iiQQQQQQQ", os.getpid(), 0, 100, 200, 3, 4, 50, 60, 10)
        m[0xC5000 + 128:0xC5000 + 192] = struct.pack("<iiQQQQQQQ", 999999, 0, 1, 2, 1, 1, 5, 6, 1)
        m[0xC5000 + 192:0xC5000 + 256] = struct.pack("<iiQQQQQQQ", os.getpid(), 1, 7, 7, 7, 7, 7, 7, 7)
        m.flush(); m.close()
    assert r.swap_counters(0) == {"page_out_bytes": 101, "page_in_bytes": 202, "evictions": 4, "faults": 5, "resident_bytes": 55,
                                  "live_bytes": 66, "host_bytes": 11, "processes": 2}
    assert r.swap_counters(1)["page_out_bytes"] == 7
    r.set_uuid(0, "GPU-x")
    mon = M.Monitor(str(tmp_path / "containers"), lambda: [M.PodInfo(uid, "default", "train", [ctr])],
                    host_gpus=lambda: [(0, "GPU-x", 123456, 42)])
    text = mon.collect()
    base = 'podnamespace="default",podname="train",ctrname="main",vdeviceid="0",deviceuuid="GPU-x"'
    assert f"vGPU_swap_page_out_bytes_total{{{base}}} 101.0" in text and f"vGPU_swap_resident_bytes{{{base}}} 55.0" in text
    assert 'HostGPUMemoryUsage{deviceidx="0",deviceuuid="GPU-x"} 123456.0' in text
    assert 'HostCoreUtilization{deviceidx="0",deviceuuid="GPU-x"} 42.0' in text
    # a process that exits takes its records with it (exit_handler path)
    r.release(os.getpid())
    assert r.swap_counters(0)["processes"] == 1 and r.swap_counters(1)["processes"] == 0
    r.close()
    # a region file written by the reference hook has no extension block: the monitor reports none, nothing breaks
    small = tmp_path / "small.cache"
    with open(path, "rb") as f:
        small.write_bytes(f.read(0xC4748))
    r2 = v.Region(str(small))
    assert r2.swap_counters(0) is None and r2.snapshot().initialized == 1
    r2.close()


def _observe_restated(regs):
    """feedback.go:197-255 (Observe, CheckBlocking :164-179, CheckPriority :181-195) on dicts {uuids, priority, rk, us}."""
    ut = {}
    for r in regs:
        if r["rk"] > 0:
            r["rk"] -= 1
            if r["rk"] > 0:
                for u in r["uuids"]:
                    if not u:
                        continue
                    ut.setdefault(u, [0, 0])[r["priority"]] += 1

    def blocking(r):
        for u in r["uuids"]:
            if u in ut:
                return any(ut[u][i] > 0 for i in range(r["priority"]))      # decided by the FIRST uuid found
        return False

    def contended(r):
        for u in r["uuids"]:
            if u in ut:
                if any(ut[u][i] > 0 for i in range(r["priority"])) or ut[u][r["priority"]] > 1:
                    return True
        return False

    for r in regs:
        if blocking(r):
            if r["rk"] >= 0:
                r["rk"] = -1
        elif r["rk"] < 0:
            r["rk"] = 0
        r["us"] = 1 if contended(r) else 0


def test_observe_equals_the_restated_feedback_loop_on_random_nodes(tmp_path):
    import random
    rng = random.Random(4)
    gpus = [f"GPU-{i:04d}-aaaa-bbbb" for i in range(4)]
    for case in range(60):
        n = rng.randint(1, 7)
        model, regions = [], []
        for i in range(n):
            k = rng.randint(1, 3)
            uu = rng.sample(gpus, k) if rng.random() < 0.9 else []
            slots = uu + [""] * (3 - len(uu))
            rng.shuffle(slots)
            m = {"uuids": slots, "priority": rng.randint(0, 1), "rk": rng.choice([-1, 0, 1, 2, 2, 3]), "us": rng.randint(0, 1)}
            d = tmp_path / f"c{case}_{i}"
            d.mkdir()
            r = v.Region(str(d / "x.cache"), create=True, mem_limits=[0] * 16, sm_limits=[100] * 16, priority=m["priority"])
            for lane, u in enumerate(slots):
                if u:
                    r.set_uuid(lane, u)
            r.set_feedback(recent_kernel=m["rk"], utilization_switch=m["us"])
            model.append(m); regions.append(r)
        for _round in range(4):
            v.monitor_observe(regions)
            _observe_restated(model)
            got = [(r.snapshot().recent_kernel, r.snapshot().utilization_switch) for r in regions]
            assert got == [(m["rk"], m["us"]) for m in model], (case, _round, model)
            if rng.random() < 0.5:                                   # a container launches again: the hook resets its counter
                j = rng.randrange(n)
                if model[j]["rk"] >= 0:
                    model[j]["rk"] = 2; regions[j].set_feedback(recent_kernel=2)
        for r in regions:
            r.close()
