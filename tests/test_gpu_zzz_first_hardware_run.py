"""First hardware run of what was developed after the round's GPU budget was spent.

These scenarios are the ones `tests/test_engine_on_functional_fake.py` runs on the functional fake driver (same scripts, sizes
scaled to a real context: the 0.6 GB a primary context takes on a B200 counts against the container's quota). They had
never met a GPU when they were committed, so each runs in a process of its own (a sticky CUDA error stays there), under a
timeout, and is marked xfail(strict=False): the round-end record then says XPASS or XFAIL per scenario without the verdict
on everything that WAS verified on hardware depending on them. The file sorts last on purpose."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402
from conftest import CUBIN, ROOT  # noqa: E402
from test_engine_on_functional_fake import (BATCH_COPY_SCRIPT, EXPLICIT_GRAPH_SCRIPT, OVERSIZED_LAUNCH_SCRIPT,  # noqa: E402
                                            PREFETCH_HINT_SCRIPT, THREADED_OPERANDS_SCRIPT)

FIRST_RUN = pytest.mark.xfail(strict=False, reason="developed on the functional fake after the round's GPU budget was spent: first run on hardware")


def _run(code, env_extra, timeout=120):
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env.update({"LIBCUDA_LOG_LEVEL": "2", "VGPU_ROOT": ROOT, "CUBIN": CUBIN})
    env.update({k: str(val) for k, val in env_extra.items()})
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@FIRST_RUN
def test_oversized_launches_run_in_host_backed_mode():
    """C ABI, 48 MiB cap, operand sets of up to 5 x 20 MiB: the surplus operands are used where they are (host-mapped)."""
    out = _run(OVERSIZED_LAUNCH_SCRIPT, {"VGPU_SWAP_HOST_BACKED": 1, "VGPU_SWAP_CHUNK_MB": 4})
    assert out["bad"] == 0 and out["refused"] == 0 and out["oversized"] >= 15 and out["inplace_uses"] >= out["oversized"] and out["live_after"] == 0, out


@FIRST_RUN
def test_default_mode_refuses_exactly_the_oversized_launches():
    out = _run(OVERSIZED_LAUNCH_SCRIPT, {"VGPU_SWAP_CHUNK_MB": 4})
    assert out["bad"] == 0 and out["refused"] >= 15 and out["inplace_uses"] == 0 and out["live_after"] == 0, out


@FIRST_RUN
@pytest.mark.parametrize("mode", ["host_backed", "default"])
def test_multi_operand_launches_from_three_threads(mode):
    out = _run(THREADED_OPERANDS_SCRIPT, {"VGPU_SWAP_HOST_BACKED": int(mode == "host_backed"), "VGPU_SWAP_CHUNK_MB": 4}, timeout=180)
    assert out["errors"] == [] and out["bad"] == 0, out
    assert out["inplace_uses"] > 10 if mode == "host_backed" else out["inplace_uses"] == 0, out


@FIRST_RUN
def test_batched_copies_3d_copy_and_address_range_through_the_hook(tmp_path):
    """32 x 64 MiB buffers = 2 GiB live; quota = context + staging + room for ~12 of them: the 8-operand batch fits, the
    24-operand batch is issued copy by copy."""
    env = v.hook_env(limit_mib=1700, oversubscribe=True, cache_path=str(tmp_path / "batch.cache"), extra={"BUF_MIB": "64"})
    out = _run(BATCH_COPY_SCRIPT, env)
    assert out["bad"] == 0 and out["ranges_ok"] == 33 and out["faults"] > 30, out


@FIRST_RUN
def test_explicitly_built_graph_replays_with_its_operands_pinned(tmp_path):
    """14 x 64 MiB cycling buffers, four of them + a copy target pinned by the graph's nodes; quota = context + staging +
    room for ~9 buffers."""
    env = v.hook_env(limit_mib=1500, oversubscribe=True, cache_path=str(tmp_path / "graph.cache"), extra={"BUF_MIB": "64"})
    out = _run(EXPLICIT_GRAPH_SCRIPT, env)
    assert out == {"bad": 0, "scratch": "0xabababababababab", "bad_copy": 0}, out


@FIRST_RUN
def test_prefetch_hints_steer_the_pager():
    """C ABI twin of cuMemPrefetchAsync: "to the host" makes the most recently used buffer the next victim, "to the device"
    pages it back in without a touch."""
    out = _run(PREFETCH_HINT_SCRIPT, {"VGPU_SWAP_CHUNK_MB": 4})
    assert out["bad"] == 0 and out["before"] == [2, 3, 4, 5] and out["after_evict_hint"] == [0, 2, 3, 4], out
    assert 5 in out["after_prefetch"] and 0 in out["after_prefetch"], out
