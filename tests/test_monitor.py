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
