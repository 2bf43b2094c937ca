"""The LD_PRELOAD path on a real B200: an unmodified driver-API program under libvgpu.so. Accounting parity against
the REFERENCE BINARY running on the same box (when it runs there), hard cap, swap mode integrity."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402
from conftest import CUBIN, LIBDIR, OREF, have_reference, run_replay  # noqa: E402
from trace_gen import gen_trace  # noqa: E402


def _strip_ctx(text):
    # contextSize is a property of the box/driver and of NVML visibility: compared separately
    out = []
    for line in text.splitlines():
        parts = [p for p in line.split() if not p.startswith(("ctx=", "tot=", "free="))]
        out.append(" ".join(parts))
    return out


def test_hard_cap_and_accounting_on_real_driver(tmp_path):
    t = tmp_path / "t.txt"
    t.write_text(gen_trace(3000, seed=0xB200, max_size=256 << 20, kinds="AAAM"))
    env = {"CUDA_DEVICE_MEMORY_LIMIT_0": "8192m", "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "new.cache")}
    new = run_replay(str(t), "new", env, fake=False)
    assert "rc=-1" in new
    # the oracle with the context size the hook measured on this box must reproduce the stream exactly
    ctx = int(new.splitlines()[0].split("ctx=")[1].split()[0])
    ora = run_replay(str(t), "oracle", dict(env, ORACLE_CTX_BYTES=str(ctx)), fake=False)
    assert new.splitlines()[1:] == ora.splitlines()[1:]
    bufs = [int(l.split("buf=")[1].split()[0]) for l in new.splitlines()[1:]]
    assert max(bufs) <= 8192 << 20


@pytest.mark.skipif(not have_reference(), reason="reference binary not shipped to this box")
def test_accounting_bit_exact_vs_reference_binary_on_real_driver(tmp_path):
    t = tmp_path / "t.txt"
    t.write_text(gen_trace(2000, seed=77, max_size=128 << 20, kinds="AAAM"))
    env_n = {"CUDA_DEVICE_MEMORY_LIMIT_0": "4096m", "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "new.cache")}
    env_r = dict(env_n, CUDA_DEVICE_MEMORY_SHARED_CACHE=str(tmp_path / "ref.cache"))
    try:
        ref = run_replay(str(t), "reference", env_r, fake=False, timeout=45)
    except Exception as e:  # the 2021-era binary may not survive this driver (cuGetExportTable patches, NVML pids)
        pytest.skip(f"reference binary does not run on this box: {e}")
    new = run_replay(str(t), "new", env_n, fake=False)
    r0, n0 = ref.splitlines()[0], new.splitlines()[0]
    if r0 == n0:
        assert new == ref                      # same context size measured: full bit parity
    else:
        assert _strip_ctx(new) == _strip_ctx(ref), (r0, n0)


def _swap_bench(tmp_path, extra_env, args):
    env = dict(os.environ)
    env.update(v.hook_env(cache_path=str(tmp_path / "sb.cache")))
    env.update(extra_env)
    env.setdefault("LIBCUDA_LOG_LEVEL", "1")
    cmd = [os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN] + args
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:] + r.stdout[-500:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not have_reference(), reason="reference binary not shipped to this box")
def test_cfg2_trace_20k_ops_at_8192m_bit_exact_vs_reference_binary_on_real_driver(tmp_path):
    """BASELINE.json configs[1] as specified (SURVEY.md §8d cfg 2): the seed-0xB200 trace, sizes up to 512 MiB, 8 GiB hard
    cap, 20 000 ops on the REAL driver — return codes, the five counter words and cuMemGetInfo after every op — reference
    binary against the new hook. (The full 100 000-op stream is pinned on the fake driver: tests/golden/ref_hashes.json;
    the reference forks `ps ax` on every quota breach, which makes 100 k ops a ten-minute run on real hardware.)"""
    t = tmp_path / "t.txt"
    t.write_text(gen_trace(20000, seed=0xB200, max_size=512 << 20, kinds="A"))
    env_n = {"CUDA_DEVICE_MEMORY_LIMIT_0": "8192m", "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "new.cache")}
    env_r = dict(env_n, CUDA_DEVICE_MEMORY_SHARED_CACHE=str(tmp_path / "ref.cache"))
    try:
        ref = run_replay(str(t), "reference", env_r, fake=False, timeout=600)
    except Exception as e:
        pytest.skip(f"reference binary does not run on this box: {e}")
    new = run_replay(str(t), "new", env_n, fake=False, timeout=600)
    assert len(new.splitlines()) == len(ref.splitlines()) > 20000
    assert sum("rc=-1" in l for l in ref.splitlines()) > 1000          # the cap is crossed thousands of times
    r0, n0 = ref.splitlines()[0], new.splitlines()[0]
    if r0 == n0:
        assert new == ref
    else:
        assert _strip_ctx(new) == _strip_ctx(ref), (r0, n0)


def test_unmodified_app_under_hook_swaps_and_verifies(tmp_path):
    """Pure demand paging (prefetch off): exact counts, every byte through the staged path's TMA kernels."""
    out = _swap_bench(tmp_path, {"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "2048m", "VGPU_SWAP_PREFETCH_MB": "0"},
                      ["--buffers", "48", "--mib", "64", "--steps", "96", "--warmup", "8", "--order", "cyclic"])
    assert out["mismatches"] == 0 and out["hooked_stats"] is True
    assert out["page_in_bytes"] == 96 * (64 << 20)          # every cyclic touch misses: 64 MiB in ...
    assert out["page_out_bytes"] >= 95 * (64 << 20)         # ... and 64 MiB out
    assert out["phys_reuses"] > 0 and out["pack_launches"] > 0 and out["unpack_launches"] > 0
    assert out["host_ms"]["vmm"] == 0                       # VMM calls belong to the pager thread


def test_prefetch_pipeline_under_the_hook_keeps_every_word(tmp_path):
    """Default engine: after the populate phase the predictor knows the cycle, the pager pages ahead with plain DMA."""
    out = _swap_bench(tmp_path, {"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "2048m"},
                      ["--buffers", "48", "--mib", "64", "--steps", "144", "--warmup", "24", "--order", "cyclic"])
    assert out["mismatches"] == 0 and out["faults"] == 144
    window = 8                                             # quota / 4 = 512 MiB = 8 buffers
    assert (144 - window) * (64 << 20) <= out["page_in_bytes"] <= (144 + window) * (64 << 20)
    assert out["direct_in_bytes"] >= 0.8 * out["page_in_bytes"] and out["direct_out_bytes"] >= 0.8 * out["page_out_bytes"]
    assert out["host_ms"]["vmm"] == 0


def test_two_processes_of_one_container_share_one_resident_quota_on_the_gpu(tmp_path):
    """VERDICT r1 'next' #4, the done-criterion: 2 processes x 6 GiB live under ONE 4 GiB quota (one region file = one
    container). Both make progress and verify every word; the node monitor's view of the region never shows more than the
    quota resident for the two engines together."""
    import threading
    import time
    cache = str(tmp_path / "shared.cache")
    env = dict(os.environ)
    env.update(v.hook_env(limit_mib=4096, oversubscribe=True, cache_path=cache))
    env["LIBCUDA_LOG_LEVEL"] = "1"
    args = [os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", "96", "--mib", "64", "--steps", "288", "--warmup", "16", "--order", "cyclic"]
    procs = [subprocess.Popen(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for _ in range(2)]
    seen, stop = [], threading.Event()

    def monitor():
        reg = None
        while not stop.is_set():
            try:
                reg = reg or v.Region(cache)
                c = reg.swap_counters(0)
                seen.append((c["processes"], c["resident_bytes"], c["live_bytes"]))
            except Exception:
                pass
            time.sleep(0.005)

    th = threading.Thread(target=monitor)
    th.start()
    outs = []
    try:
        for p in procs:
            out, err = p.communicate(timeout=300)
            assert p.returncode == 0, err[-2000:] + out[-300:]
            outs.append(json.loads(out.strip().splitlines()[-1]))
    finally:
        stop.set(); th.join()
        for p in procs:                      # never leave a sibling running behind a failed assertion
            if p.poll() is None:
                p.kill(); p.communicate()
    assert all(o["mismatches"] == 0 and o["verified"] == 1 for o in outs)
    assert all(o["page_in_bytes"] >= 250 * (64 << 20) for o in outs)          # both made progress through their whole loop
    both = [s for s in seen if s[0] == 2]
    assert len(both) > 20
    assert max(s[1] for s in both) <= 4096 << 20, max(s[1] for s in both)     # sum of both resident sets within the ONE quota
    assert max(s[2] for s in both) >= 8 << 30                                  # while at least twice the quota was live in the two together
                                                                               # (12 GiB at the peak; the two populate phases only partly overlap)


def test_read_mostly_advice_through_cumemadvise_saves_the_write_back(tmp_path):
    """Every second buffer is advised read-mostly with cuMemAdvise (what a UVM application does; the reference's swappable
    memory is managed memory) and only read: its evictions are clean, so page-out traffic is about half the page-in's."""
    out = _swap_bench(tmp_path, {"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "2048m"},
                      ["--buffers", "48", "--mib", "64", "--steps", "192", "--warmup", "48", "--order", "cyclic", "--ro-every", "2"])
    assert out["mismatches"] == 0
    assert out["clean_evictions"] >= 60
    assert out["page_out_bytes"] <= 0.65 * out["page_in_bytes"]


def test_zipf_order_hits_resident_set(tmp_path):
    out = _swap_bench(tmp_path, {"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "2048m"},
                      ["--buffers", "48", "--mib", "64", "--steps", "200", "--warmup", "50", "--order", "zipf"])
    assert out["mismatches"] == 0
    assert out["page_in_bytes"] < 200 * (64 << 20) * 0.8     # hot buffers stay resident under LRU


def test_hard_cap_without_oversubscribe_refuses(tmp_path):
    env = dict(os.environ)
    env.update(v.hook_env(limit_mib=1024, cache_path=str(tmp_path / "hc.cache")))
    r = subprocess.run([os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", "32", "--mib", "64", "--steps", "4"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 3 and "rc=-1" in r.stdout        # cuMemAlloc_v2 -> (CUresult)-1 past the 1 GiB cap


def _gemm_loop(env_extra, n=4096, seconds=6):
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env.update(env_extra)
    r = subprocess.run([os.path.join(LIBDIR, "gemm_loop"), str(n), str(seconds)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-300:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    out["stderr"] = r.stderr
    return out


def _limiter_stats(stderr):
    import re
    m = re.search(r"limiter: limit=(\d+)% active=(\d) launches=(\d+) stamps=(\d+) groups=(\d+) busy_ms=([\d.]+) throttle_ms=([\d.]+)", stderr)
    assert m, stderr[-800:]
    return {"limit": int(m.group(1)), "active": int(m.group(2)), "launches": int(m.group(3)), "busy_ms": float(m.group(6)), "throttle_ms": float(m.group(7))}


def test_sm_limit_reaches_a_cublas_application_through_cudart(tmp_path):
    """BASELINE.json configs[3] on a cuBLAS SGEMM loop: a cudart application reaches the driver through dlsym /
    cuGetProcAddress, i.e. through the hook's symbol routing. The device-stamped busy time of the intercepted launches must
    sit at the quota (the independent duty-cycle measurement is test_sm_limit_on_a_driver_api_launch_loop: the app's own
    event pairs are blind to host-side throttling, and bare GEMM rates on this 1 kW part depend on power state)."""
    hooked = dict(v.hook_env(sm_limit=30, cache_path=str(tmp_path / "lim.cache")), GPU_CORE_UTILIZATION_POLICY="force", VGPU_PRINT_STATS="1")
    lim = _gemm_loop(hooked)
    st = _limiter_stats(lim["stderr"])
    assert st["limit"] == 30 and st["active"] == 1 and st["launches"] >= lim["gemms"]
    assert 0.24 <= st["busy_ms"] / 1e3 / lim["wall_s"] <= 0.36, (st, lim["wall_s"])
    off = dict(v.hook_env(sm_limit=30, cache_path=str(tmp_path / "off.cache")), GPU_CORE_UTILIZATION_POLICY="disable", VGPU_PRINT_STATS="1")
    free = _gemm_loop(off, seconds=3)
    assert _limiter_stats(free["stderr"])["launches"] == 0 and free["gemms"] / free["wall_s"] > 2.0 * lim["gemms"] / lim["wall_s"]


def _sample_gpu_util(stop, out):
    """NVML's utilisation counter of GPU 0 (what nvidia-smi shows), every 100 ms until `stop` is set."""
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        out.append(pynvml.nvmlDeviceGetUtilizationRates(h).gpu)
        time.sleep(0.1)
    pynvml.nvmlShutdown()


def test_cfg4_sgemm_8192_for_30s_nvml_utilisation_bare_reference_new(tmp_path):
    """BASELINE.json configs[3] AS WRITTEN: cuBLAS SGEMM 8192^3 for 30 s with CUDA_DEVICE_SM_LIMIT=30, judged by a counter
    the limiter does not produce: NVML's GPU utilisation sampled from outside the process — for the bare loop, the
    reference hook (rate_limiter@0x4591a / utilization_watcher@0x46710) and the new limiter. The table goes to
    gpurun_out/cfg4_table.json (copied to profiles/ by the developer)."""
    import threading
    table = {}

    def run(name, env):
        stop, samples = threading.Event(), []
        th = threading.Thread(target=_sample_gpu_util, args=(stop, samples))
        th.start()
        try:
            out = _gemm_loop(env, n=8192, seconds=30)
        finally:
            stop.set(); th.join()
        body = sorted(samples[20:-5] or samples)             # skip start-up and tear-down
        table[name] = {"mean_util": round(sum(body) / len(body), 1), "p95_util": body[int(0.95 * (len(body) - 1))], "p05_util": body[int(0.05 * (len(body) - 1))],
                       "samples": len(body), "gemms": out["gemms"], "tflops": out["tflops"], "app_duty": out["duty"]}

    run("bare", {})
    run("new_limit30", dict(v.hook_env(sm_limit=30, cache_path=str(tmp_path / "n.cache")), GPU_CORE_UTILIZATION_POLICY="force"))
    if have_reference():
        os.makedirs("/tmp/vgpulock", exist_ok=True)
        try:
            run("reference_limit30", {"LD_PRELOAD": os.path.join(OREF, "dlsym_shim.so") + ":" + os.path.join(OREF, "libvgpu.so"), "CUDA_DEVICE_SM_LIMIT": "30",
                                      "GPU_CORE_UTILIZATION_POLICY": "force", "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "r.cache"), "LIBCUDA_LOG_LEVEL": "0"})
        except AssertionError as e:
            table["reference_limit30"] = {"error": str(e)[-300:]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(table, open("gpurun_out/cfg4_table.json", "w"), indent=1)
    print(json.dumps(table))
    assert table["bare"]["mean_util"] >= 85
    assert 20 <= table["new_limit30"]["mean_util"] <= 42, table      # the quota, as NVML sees it from outside
    assert table["new_limit30"]["gemms"] < 0.45 * table["bare"]["gemms"]


def test_cudart_application_is_accounted_through_cugetprocaddress(tmp_path):
    """cudaMalloc in gemm_loop (3 x 64 MiB + cuBLAS workspace) must show up in the region: proves the dlsym /
    cuGetProcAddress routing catches a runtime-API program, not only directly linked driver-API calls."""
    cache = str(tmp_path / "rt.cache")
    env = dict(os.environ)
    env.update(v.hook_env(limit_mib=100, cache_path=cache))
    r = subprocess.run([os.path.join(LIBDIR, "gemm_loop"), "4096", "1"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode != 0, r.stdout      # 3 x 64 MiB does not fit a 100 MiB cap (cudart reports the hook's code as an error)
    env["VGPU_STRICT_CUDA_ERRORS"] = "1"
    r = subprocess.run([os.path.join(LIBDIR, "gemm_loop"), "4096", "1"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode != 0 and "out of memory" in (r.stdout + r.stderr).lower()


def _launch_loop(env_extra, mib, seconds=5):
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env.update(env_extra)
    r = subprocess.run([os.path.join(LIBDIR, "launch_loop"), CUBIN, str(mib), str(seconds)], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=200)
    assert r.returncode == 0, r.stderr[-1500:] + r.stdout[-300:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mib,quota", [(2048, 30), (2048, 60), (16, 30)])
def test_sm_limit_on_a_driver_api_launch_loop(tmp_path, mib, quota):
    """Independent duty-cycle check: launch RATE under the quota relative to the bare (GPU-bound) rate of the same loop,
    for ~0.7 ms kernels and for ~10 us kernels (stamps amortised over groups of launches). The reference hook on the same
    loop: no limiting at all at 30 % (94 % utilisation) and a multi-second stall at 60 %
    (profiles/r01_cfg4_limiter_comparison.txt) — its delta() overflows int32 on B200."""
    bare = _launch_loop({}, mib, seconds=3)
    env = dict(v.hook_env(sm_limit=quota, cache_path=str(tmp_path / "ll.cache")), GPU_CORE_UTILIZATION_POLICY="force")
    lim = _launch_loop(env, mib)
    ratio = (lim["launches"] / lim["wall_s"]) / (bare["launches"] / bare["wall_s"])
    assert 0.75 * quota / 100 <= ratio <= 1.25 * quota / 100, (ratio, bare, lim)


def _wide_probe(env_extra, seconds=4):
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env.update(env_extra)
    r = subprocess.run([os.path.join(LIBDIR, "wide_probe"), CUBIN, str(seconds)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=200)
    assert r.returncode == 0, r.stderr[-1500:] + r.stdout[-300:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_async_pool_and_vmm_allocations_count_against_the_quota(tmp_path):
    """SURVEY.md §8(f) #4 on the real driver: cuMemAllocAsync and cuMemCreate are charged and refused like cuMemAlloc
    (the reference forwards both unaccounted, so a PyTorch caching allocator escapes its quota there)."""
    M = 1 << 20
    bare = _wide_probe({}, seconds=1)
    assert bare["async"]["rc"] == 0 and bare["async"]["rc_big"] == 0 and bare["vmm"]["rc_big"] == 0 and bare["vmm"]["data_ok"] == 1
    out = _wide_probe(v.hook_env(limit_mib=2048, cache_path=str(tmp_path / "w.cache")), seconds=1)
    assert out["async"] == {"rc": 0, "rc_big": 2, "rc_free": 0, "charged": 512 * M, "returned": 512 * M}
    assert out["vmm"]["rc"] == 0 and out["vmm"]["rc_big"] == 2 and out["vmm"]["data_ok"] == 1
    assert out["vmm"]["charged"] == out["vmm"]["returned"] and 512 * M <= out["vmm"]["charged"] < 520 * M
    assert out["graph"]["word0"] == out["graph"]["expect_word0"]
    off = _wide_probe(dict(v.hook_env(limit_mib=2048, cache_path=str(tmp_path / "r.cache")), VGPU_REFERENCE_COVERAGE="1"), seconds=1)
    assert off["async"]["rc_big"] == 0 and off["vmm"]["rc_big"] == 0 and off["async"]["charged"] == 0


def test_sm_limit_holds_a_replayed_cuda_graph_to_its_quota(tmp_path):
    """A captured decode-style loop: capture must survive the limiter (no stamp kernels inside the capture), and the
    replays are billed — graph launch RATE under a 30 % quota vs the bare rate. The reference has no hook on
    cuGraphLaunch at all."""
    bare = _wide_probe({}, seconds=3)
    env = dict(v.hook_env(sm_limit=30, cache_path=str(tmp_path / "g.cache")), GPU_CORE_UTILIZATION_POLICY="force")
    lim = _wide_probe(env, seconds=5)
    assert lim["graph"]["word0"] == lim["graph"]["expect_word0"]
    ratio = (lim["graph"]["launches"] / lim["graph"]["wall_s"]) / (bare["graph"]["launches"] / bare["graph"]["wall_s"])
    assert 0.22 <= ratio <= 0.38, (ratio, bare, lim)


def test_captured_kernels_keep_their_swappable_operands_resident(tmp_path):
    """Swap mode + stream capture: the captured kernels' buffer lives in the swap arena; capture pins it resident (a
    replay cannot fault), so every replay sees mapped memory and the arithmetic is exact."""
    env = v.hook_env(limit_mib=4096, oversubscribe=True, cache_path=str(tmp_path / "s.cache"))
    out = _wide_probe(env, seconds=1)
    assert out["graph"]["word0"] == out["graph"]["expect_word0"] and out["graph"]["launches"] > 0
    assert out["async"]["rc"] == 0 and out["vmm"]["data_ok"] == 1


def test_node_monitor_sees_the_swap_counters_of_a_running_container(tmp_path):
    """The hook publishes its swap engine's counters into the extension block of the region file while the application
    runs; the host-side monitor (another process, C ABI only) reads residency and page traffic from there."""
    import time
    cache = str(tmp_path / "live.cache")
    env = dict(os.environ)
    env.update(v.hook_env(limit_mib=2048, oversubscribe=True, cache_path=cache))
    p = subprocess.Popen([os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", "64", "--mib", "64", "--steps", "400", "--warmup", "8",
                          "--order", "cyclic"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    seen, region = [], None
    try:
        deadline = time.time() + 120
        while p.poll() is None and time.time() < deadline:
            if region is None and os.path.exists(cache):
                try:
                    region = v.Region(cache)
                except OSError:
                    region = None
            if region is not None:
                c = region.swap_counters(0)
                if c and c["processes"]:
                    seen.append(c)
            time.sleep(0.05)
        out, err = p.communicate(timeout=120)
    finally:
        if p.poll() is None:
            p.kill()
    assert p.returncode == 0, err[-2000:]
    final = json.loads(out.strip().splitlines()[-1])
    assert final["mismatches"] == 0
    assert len(seen) > 5
    assert max(c["page_out_bytes"] for c in seen) > 10 * (64 << 20) and max(c["page_in_bytes"] for c in seen) > 10 * (64 << 20)
    assert all(c["resident_bytes"] <= 2048 << 20 for c in seen)                  # the quota bounds residency ...
    assert max(c["live_bytes"] for c in seen) == 64 * (64 << 20)                 # ... not live bytes
    assert [c["page_out_bytes"] for c in seen] == sorted(c["page_out_bytes"] for c in seen)
    assert region.swap_counters(0)["processes"] == 0                              # the exiting process took its record along
    region.close()
