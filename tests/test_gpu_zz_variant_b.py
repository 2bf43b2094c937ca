"""SURVEY.md §8d cfg 3 variant B on a real B200: the oversubscribed alloc+touch loop with buffers of log-uniform size
(2-256 MiB, seed 0x5EED) instead of uniform 64 MiB — the unmodified driver-API program under LD_PRELOAD=libvgpu.so.
Runs after the other GPU suites (file name): it is the newest scenario, first exercised on a GPU by the round-end pass."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402
from conftest import CUBIN, LIBDIR  # noqa: E402


@pytest.mark.parametrize("order", ["cyclic", "zipf"])
def test_ragged_sizes_swap_and_verify(tmp_path, order):
    env = dict(os.environ)
    env.update(v.hook_env(limit_mib=4096, oversubscribe=True, cache_path=str(tmp_path / "vb.cache")))
    env.setdefault("LIBCUDA_LOG_LEVEL", "1")
    cmd = [os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", "128", "--mib", "64", "--ragged-lo", "2", "--ragged-hi", "256",
           "--steps", "240", "--warmup", "16", "--order", order]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:] + r.stdout[-500:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["mismatches"] == 0 and out["verified"] == 1 and out["hooked_stats"] is True and out["ragged_mib"] == [2, 256]
    assert out["page_out_bytes"] > 0
    if order == "cyclic":                       # 8 GiB live under a 4 GiB quota, LRU worst case: every touch misses
        assert abs(out["page_in_bytes"] - out["touched_bytes"]) <= 8 * (256 << 20)     # the prefetch window straddles the timed region
    else:
        assert out["page_in_bytes"] < out["touched_bytes"]
    gbps = (out["page_in_bytes"] + out["page_out_bytes"]) / (out["event_ms"] * 1e-3) / 1e9
    print(f"variant B {order}: {gbps:.1f} GB/s combined, {out['buffers']} buffers")
