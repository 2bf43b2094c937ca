#!/usr/bin/env python
"""bench.py — vmem swap throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            this repo (one rank per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's swap path (CUDA UVM paging on host cores)

Workload (BASELINE.json configs[2], SURVEY.md §8d cfg 3): a container with an 8 GiB gpumem quota holds 1152 x 64 MiB
buffers (8 GiB + 64 GiB oversubscribed) and touches them cyclically with a read-modify-write kernel — the LRU worst
case, every touch pages 64 MiB in and 64 MiB out. One "step" = 16 touches = 1 GiB in + 1 GiB out over the host link.

  value : page traffic GB/s (in + out) with the loop driven straight through the engine's C ABI in this process
  e2e   : the same loop as an UNMODIFIED driver-API program (swap_bench) under LD_PRELOAD=libvgpu.so — the
          reference-facing boundary; h2d/d2h bytes per step are the page-in/page-out bytes the hook moved
  roofline      : the TMA pack/unpack kernel (dominant kernel) against measured HBM copy bandwidth
  link_roofline : page traffic against the pinned-memcpy bandwidth of this box, measured in the same run
Multi-GPU: replicas only — every GPU enforces its own container, no collective on the data path;
torch.distributed is used for the barrier and the max-over-ranks reduction of the timing.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "k8s-device-plugin_b200")
LIBDIR = os.path.join(PKG, "lib")
CUBIN = os.path.join(PKG, "build", "vgpu_kernels.cubin")
OREF = os.path.join(ROOT, "oracle", "_ref")
MiB, GiB = 1 << 20, 1 << 30
BUF_MIB = 64
TOUCHES_PER_STEP = 16
QUOTA_MIB = 8192


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "1000"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa(torch, local):
    """Run this rank (and the app it spawns) on the CPUs of the GPU's NUMA node, so the pinned page pool is allocated
    next to the PCIe root the GPU hangs off (first-touch). Returns the node or None when sysfs does not say."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def host_budget_bytes(n_ranks):
    """Pinned host memory each rank may use for its page pool (the GPU box's RAM is shared by all ranks)."""
    avail = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            avail = int(line.split()[1]) * 1024
    return int(avail * 0.70 / max(n_ranks, 1))


def choose_workload(n_ranks):
    over = 64 * GiB
    budget = host_budget_bytes(n_ranks)
    # the e2e arm runs after the in-process arm has released its pool, so each arm may use the whole per-rank budget
    while over + 2 * GiB > budget and over > 8 * GiB:
        over //= 2
    nbuf = (QUOTA_MIB * MiB + over) // (BUF_MIB * MiB)
    return nbuf, over


def workload_name(nbuf, over):
    """config.workload — one string for both arms (the reference arm times a bounded sample of the same workload)."""
    return (f"{nbuf}x{BUF_MIB}MiB alloc+touch loop, {QUOTA_MIB}MiB gpumem quota, {over >> 30} GiB oversubscribed, cyclic RMW touch, "
            f"step={TOUCHES_PER_STEP} touches")


def measure_link(torch, barrier=None, concurrent=False):
    """Pinned-memcpy bandwidth of this box (the end-to-end roofline): 1 GiB, each direction and both at once. Alone
    (N=1): best of 5. With several ranks every copy starts behind a barrier, so all GPUs pull on the host at the same
    time — GPUs behind one PCIe switch share its uplink — and the MEDIAN of 5 is kept: the ceiling the replicas
    actually share, not the one a lone GPU sees."""
    n = 1 * GiB
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    got = {"h2d": [], "d2h": [], "bidir": []}
    sync = barrier if (barrier and concurrent) else torch.cuda.synchronize
    for _ in range(5):
        for key, fn in (("h2d", lambda: d1.copy_(h1, non_blocking=True)), ("d2h", lambda: h1.copy_(d1, non_blocking=True))):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sync()
            with torch.cuda.stream(s1):
                a.record(); fn(); b.record()
            b.synchronize()
            got[key].append(n / a.elapsed_time(b) / 1e6)
        sync()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        got["bidir"].append(2 * n / (time.perf_counter() - t0) / 1e9)
    del h1, h2, d1, d2
    pick = (lambda v: sorted(v)[len(v) // 2]) if concurrent else max
    return {k: pick(v) for k, v in got.items()}


def run_engine_arm(torch, v, nbuf, steps, warmup, barrier):
    """value: the alloc+touch loop through the C ABI (no intercept layer), timed with CUDA events on the touch stream."""
    L = v.lib()
    st = torch.cuda.current_stream().cuda_stream
    stp = C.c_void_p(st)
    sw = v.Swap(dev=torch.cuda.current_device(), resident_cap=QUOTA_MIB * MiB, profile=True)
    nbytes, nwords = BUF_MIB * MiB, BUF_MIB * MiB // 8
    bufs = []
    for i in range(nbuf):
        p = sw.alloc(nbytes)
        bufs.append(p)
        sw.acquire([p], st)
        L.vgpu_wl_fill(p, nwords, i, stp)
        sw.release([p], st)
    torch.cuda.synchronize()
    touches = [0] * nbuf
    pos = 0

    def step():
        nonlocal pos
        for _ in range(TOUCHES_PER_STEP):
            i = pos % nbuf
            sw.acquire([bufs[i]], st)
            L.vgpu_wl_touch(bufs[i], nwords, stp)
            sw.release([bufs[i]], st)
            touches[i] += 1
            pos += 1

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    sw.drain()
    s0 = sw.stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")        # lets `ncu --nvtx --nvtx-include "timed/"` list exactly these launches
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sw.drain()
    torch.cuda.nvtx.range_pop()
    barrier()
    ms = e0.elapsed_time(e1)
    s1 = sw.stats()
    # integrity of everything that went through the engine (outside the timed region)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for i, p in enumerate(bufs):
        sw.acquire([p], st)
        L.vgpu_wl_verify(p, nwords, i, touches[i], cnt.data_ptr(), stp)
        sw.release([p], st)
    torch.cuda.synchronize()
    bad = int(cnt.item())
    d = {k: s1[k] - s0[k] for k in s1 if isinstance(s1[k], (int, float))}
    sw.close()
    return ms, d, bad


def spawn_app(gpu, nbuf, steps, warmup, mode, wait_stdin, extra_args=(), ballast_mib=0):
    """The unmodified driver-API app. mode: 'new' (LD_PRELOAD=libvgpu.so), 'refhook' (reference binary), 'managed'."""
    import k8s_device_plugin_b200 as v
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    cache = f"/tmp/vgpu_bench_{os.getpid()}_{gpu}_{mode}.cache"
    if os.path.exists(cache):
        os.remove(cache)
    args = [os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", str(nbuf), "--mib", str(BUF_MIB),
            "--steps", str(steps * TOUCHES_PER_STEP), "--warmup", str(warmup * TOUCHES_PER_STEP), "--order", "cyclic",
            "--wait-stdin", "1" if wait_stdin else "0"] + list(extra_args)
    if mode == "new":
        env.update(v.hook_env(limit_mib=QUOTA_MIB, oversubscribe=True, cache_path=cache))
        env["LIBCUDA_LOG_LEVEL"] = "1"
        args += ["--profile", "0"]      # the headline arm carries no profiling events / in-kernel stamps; arm 1 provides the roofline
    elif mode == "refhook":
        os.makedirs("/tmp/vgpulock", exist_ok=True)
        env["LD_PRELOAD"] = os.path.join(OREF, "dlsym_shim.so") + ":" + os.path.join(OREF, "libvgpu.so")
        env.update({"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "250000m", "CUDA_DEVICE_MEMORY_SHARED_CACHE": cache,
                    "LIBCUDA_LOG_LEVEL": "0"})
        args += ["--ballast-mib", str(ballast_mib)]
    else:
        args += ["--managed", "1", "--ballast-mib", str(ballast_mib)]
    return subprocess.Popen(args, env=env, stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def finish_app(p, wait_stdin, barrier, timeout=3600):
    if wait_stdin:
        for line in p.stderr:              # wait for the populate + warm-up phases of this rank
            if line.startswith("READY"):
                break
        barrier()
        try:
            p.stdin.write("go\n"); p.stdin.flush()
        except Exception:
            pass
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        p.kill()
        p.communicate()
        raise RuntimeError(f"swap_bench did not finish within {timeout} s")
    if p.returncode != 0:
        raise RuntimeError(f"swap_bench failed rc={p.returncode}: {err[-2000:]} {out[-500:]}")
    return json.loads(out.strip().splitlines()[-1])


def intercept_overhead(gpu):
    """Second half of BASELINE.json's metric: ns per cuMemAlloc+cuMemFree(1 MiB) and per empty cuLaunchKernel, bare vs
    under LD_PRELOAD=libvgpu.so (hard-cap mode, 8 GiB limit), best of 3 alternating runs each."""
    import k8s_device_plugin_b200 as v
    exe = os.path.join(LIBDIR, "intercept_bench")
    best = {"bare": None, "hooked": None}
    for _ in range(3):
        for mode in ("bare", "hooked"):
            env = dict(os.environ)
            env.pop("LD_PRELOAD", None)
            env["CUDA_VISIBLE_DEVICES"] = str(gpu)
            if mode == "hooked":
                cache = f"/tmp/vgpu_bench_{os.getpid()}_ib.cache"
                if os.path.exists(cache):
                    os.remove(cache)
                env.update(v.hook_env(limit_mib=QUOTA_MIB, cache_path=cache))
                env["LIBCUDA_LOG_LEVEL"] = "0"
            try:
                r = subprocess.run([exe, CUBIN, "3000", "100000"], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=120)
                j = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:
                continue
            b = best[mode]
            best[mode] = j if b is None else {k: min(b[k], j[k]) if isinstance(j[k], float) else j[k] for k in j}
    if not best["bare"] or not best["hooked"]:
        return None
    out = {}
    for k in ("alloc_free_1mib_ns", "launch_empty_ns"):
        out[k] = {"bare": best["bare"][k], "hooked": best["hooked"][k],
                  "overhead_pct": round(100.0 * (best["hooked"][k] - best["bare"][k]) / best["bare"][k], 2)}
    return out


def reference_arm(args):
    """The reference's swap path = CUDA UVM demand paging (cuMemAllocManaged, cuMemoryAllocate libvgpu.so@0x315da),
    executed by the NVIDIA UVM driver's fault-servicing threads on the host cores. A ballast allocation pins all but
    8 GiB of the GPU so that UVM sees the same resident budget as the quota; bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.cuda.init()
    free_b, total_b = torch.cuda.mem_get_info(0)
    ballast_mib = max(0, (free_b - QUOTA_MIB * MiB - 1536 * MiB) // MiB)
    nbuf = (QUOTA_MIB + 8192) // BUF_MIB            # 8 GiB resident + 8 GiB oversubscribed: bounded sample
    # exactly the K timed steps and W warm-up steps asked for (a step = 16 touches of 64 MiB, ~0.1 s under UVM on this class
    # of box); only absurd requests are clamped, and the line says what ran
    steps, warmup = max(1, min(args.steps, 256)), max(1, min(args.warmup, 16))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    full_nbuf, full_over = choose_workload(world)
    kind, res, note = "reference", None, ""
    if os.path.exists(os.path.join(OREF, "libvgpu.so")):
        holder = None
        try:
            # the ballast lives in an UNHOOKED helper process: inside the reference-hooked process every large cuMemAlloc
            # becomes managed memory (cuMemoryAllocate allocmode 0) and would not pin anything
            henv = dict(os.environ, CUDA_VISIBLE_DEVICES="0", SWAP_BENCH_HOLD_MIB=str(ballast_mib))
            henv.pop("LD_PRELOAD", None)
            holder = subprocess.Popen([os.path.join(LIBDIR, "swap_bench")], env=henv, stdin=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            for line in holder.stderr:
                if line.startswith("READY"):
                    break
            p = spawn_app(0, nbuf, steps, warmup, "refhook", False, ballast_mib=0)
            res = finish_app(p, False, lambda: None, timeout=240 + 2 * steps)
            if res["event_ms"] / max(res["steps"], 1) < 0.5 * BUF_MIB * MiB / 60e6:   # faster than the link allows: nothing was paged
                raise RuntimeError("no paging happened under the reference hook")
            note = ("lib/nvidia/libvgpu.so binary preloaded (oracle/dlsym_shim.so first; CUDA_OVERSUBSCRIBE=true -> cuMemAllocManaged); "
                    "ballast held by an unhooked helper process")
        except Exception as e:
            res = None
            note = f"reference binary run unusable on this box ({str(e)[:120]}); "
        finally:
            if holder:
                try:
                    holder.stdin.close()
                    holder.wait(timeout=30)
                except Exception:
                    holder.kill()
    if res is None:
        p = spawn_app(0, nbuf, steps, warmup, "managed", False, ballast_mib=ballast_mib)
        res = finish_app(p, False, lambda: None)
        kind = "port"
        note += "direct cuMemAllocManaged (the call the reference hook makes in allocmode 0)"
    touched = res["steps"] * BUF_MIB * MiB
    gbs = 2 * touched / (res["event_ms"] / 1e3) / 1e9   # cyclic + RMW: every touch misses, and evicts a dirty buffer
    line = {
        "metric": "vmem_swap_GBps", "value": round(gbs, 3), "unit": "GB/s", "impl": "reference", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(res["event_ms"] / steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(full_nbuf, full_over), "inputs": "larger than L2 (each step streams 1 GiB in + 1 GiB out)",
                   "parallelism": f"rank 0 of {world} (host-side paging by the UVM driver: one measurement per box)",
                   "sample": f"bounded: {nbuf}x{BUF_MIB}MiB managed buffers over {QUOTA_MIB}MiB of physical memory (ballast {ballast_mib} MiB), same cyclic RMW touch"},
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": os.cpu_count(), "kind": kind,
                         "sample": f"{steps * TOUCHES_PER_STEP} touches of {BUF_MIB} MiB; paging by the UVM driver's fault threads on host cores + copy engines; {note}"},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mismatches": res["mismatches"],
    }
    print(json.dumps(line))


def aggregate(dist, device, ms, nbytes):
    """Whole-job throughput of N replicas: units all ranks processed / max-over-ranks time (GB/s)."""
    import torch
    if dist is None:
        return ms, nbytes, nbytes / (ms / 1e3) / 1e9
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    b = torch.tensor([float(nbytes)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(t.item()), float(b.item()), float(b.item()) / (float(t.item()) / 1e3) / 1e9


def dry_run_cpu(args):
    """No GPU: gloo backend, synthetic per-rank measurements (rank r: 100+r ms, (r+1) GB) through the same
    barrier / max-time / sum-bytes path the GPU run uses."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    d = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
        dist.barrier()
        d = dist
    t_max, total, value = aggregate(d, "cpu", 100.0 + rank, (rank + 1) * 1e9)
    if d:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"metric": "vmem_swap_GBps", "value": round(value, 6), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak",
                          "data": "dry-run", "total_bytes": total, "t_max_ms": t_max}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=72)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--dry-run-cpu", action="store_true", help="exercise the multi-rank plumbing with gloo and synthetic per-rank numbers (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return reference_arm(args)
    if args.dry_run_cpu:
        return dry_run_cpu(args)

    import torch
    import k8s_device_plugin_b200 as v
    v.lib()                                           # no CUDA extension -> ImportError, never a fallback
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.zeros(1, device="cuda")
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if not dist:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    numa = bind_to_gpu_numa(torch, local)
    nbuf, over = choose_workload(world)
    link = measure_link(torch, barrier, concurrent=world > 1)
    link_mean = {k: reduce(v_, "SUM") / world for k, v_ in link.items()}
    link_min = reduce(link["bidir"], "MIN")
    sampler = ClockSampler(local)
    sampler.start()

    # ---- arm 1: engine through the C ABI
    ms, d, bad = run_engine_arm(torch, v, nbuf, args.steps, args.warmup, barrier)
    page_bytes = d["page_in_bytes"] + d["page_out_bytes"]
    t_max, total_bytes, value = aggregate(dist, "cuda", ms, page_bytes)
    kern_ms = d["pack_ms"] + d["unpack_ms"]
    span_ms = d.get("pack_span_ms", 0.0) + d.get("unpack_span_ms", 0.0)
    kern_bytes = 2 * (d["pack_bytes"] + d["unpack_bytes"])      # algorithmic: read + write of every byte moved
    kern_launches = d["pack_launches"] + d["unpack_launches"]
    launches = kern_launches + d["scan_launches"] + args.steps * TOUCHES_PER_STEP
    clocks = sampler.finish()

    # ---- arm 2: unmodified app under LD_PRELOAD (reference-facing boundary)
    p = spawn_app(local, nbuf, args.steps, args.warmup, "new", True)
    e2e = finish_app(p, True, barrier)
    e2e_ms, e2e_bytes, e2e_value = aggregate(dist, "cuda", e2e["event_ms"], e2e["page_in_bytes"] + e2e["page_out_bytes"])
    e2e_bad = reduce(e2e["mismatches"], "SUM")

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's swap path on this box's host cores
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4", "--warmup", "1"],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # reported, never silently replaced
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {str(e)[:200]}"}

    overhead = intercept_overhead(local) if (rank == 0 and world == 1) else None

    if rank == 0:
        peak, how = hbm_peak()
        achieved = kern_bytes / (kern_ms / 1e3) / 1e9 if kern_ms > 0 else None
        line = {
            "metric": "vmem_swap_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t_max / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(nbuf, over),
                       "inputs": "larger than L2 (each step streams 1 GiB in + 1 GiB out)", "parallelism": f"replicas x{world}", "numa_node": numa},
            "e2e": {"value": round(e2e_value, 3), "unit": "GB/s",
                    "h2d_bytes_per_step": int(e2e["page_in_bytes"] // args.steps), "d2h_bytes_per_step": int(e2e["page_out_bytes"] // args.steps),
                    "via": "LD_PRELOAD=libvgpu.so on an unmodified driver-API app (cuMemAlloc_v2/cuLaunchKernel intercept)",
                    "wall_ms": e2e["wall_ms"], "mismatches": int(e2e_bad)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "vgpu_pack_tma (pack + unpack launches inside the timed region)",
                         "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4) if achieved else None, "peak_source": how,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same 32 MiB-chunk launch, ncu --set full
                         # (profiles/r01b_pack_tma_full.md: 33.567 MB read + 0.456 MB written back so far; the 33.55 MB written
                         # by the kernel are still dirty in the 126 MB L2 when it ends)
                         "traffic": 34023936,
                         "launches": int(kern_launches), "avg_launch_us": round(kern_ms * 1e3 / max(kern_launches, 1), 2),
                         "bytes_per_launch": int(kern_bytes // max(kern_launches, 1)),
                         "note": "achieved = CUDA-event brackets on the engine's kernel streams inside the timed region; they include the host's "
                                 "event->launch gap (the streams are idle between chunks) and SM sharing with the application's touch kernel. "
                                 "device_span = the same launches timed by in-kernel %globaltimer stamps (min CTA start .. max CTA end); "
                                 "isolated figures and ncu captures: profiles/",
                         "device_span": {"achieved": round(kern_bytes / (span_ms / 1e3) / 1e9, 1) if span_ms > 0 else None,
                                         "frac": round(kern_bytes / (span_ms / 1e3) / 1e9 / peak, 4) if span_ms > 0 else None,
                                         "avg_launch_us": round(span_ms * 1e3 / max(kern_launches, 1), 2)}},
            "link_roofline": {"bound": "host-link", "achieved": round(value / world, 3), "peak": round(link_mean["bidir"], 2), "unit": "GB/s",
                              "frac": round(value / world / link_mean["bidir"], 4) if link_mean["bidir"] else None,
                              "h2d_peak": round(link_mean["h2d"], 2), "d2h_peak": round(link_mean["d2h"], 2),
                              "peak_min_over_ranks": round(link_min, 2),
                              "peak_source": ("pinned 1 GiB cudaMemcpyAsync both directions at once, measured in this run" if world == 1 else
                                              f"per-GPU mean over {world} ranks copying 1 GiB each way AT THE SAME TIME (barrier-started, median of 5): "
                                              "the host link the replicas share, measured in this run")},
            "cpu_baseline": cpu,
            "intercept_overhead": overhead,
            "mismatches": bad,
            "engine": dict({k: d[k] for k in ("faults", "evictions", "scans", "scan_launches", "phys_creates", "phys_reuses")},
                           # rank 0's calling-thread time per step inside admissions: where the step goes when it is not DMA
                           host_ms_per_step={k[5:-3]: round(d[k] / 1e6 / args.steps, 3) for k in
                                             ("host_admit_ns", "host_scan_ns", "host_packsync_ns", "host_vmm_ns", "host_ring_ns") if k in d}),
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
