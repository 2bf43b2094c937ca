#!/usr/bin/env python
"""bench.py — vmem swap throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            this repo (one rank per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's swap path (CUDA UVM paging on host cores)
    python bench.py --config cfg5 ...                          BASELINE.json configs[4]: quota = 50 % of the HBM, + 8 GiB

Workload (BASELINE.json configs[2], SURVEY.md §8d cfg 3): a container with an 8 GiB gpumem quota holds 1152 x 64 MiB
buffers (8 GiB + 64 GiB oversubscribed) and touches them cyclically with a read-modify-write kernel — the LRU worst
case, every touch pages 64 MiB in and 64 MiB out. One "step" = 16 touches = 1 GiB in + 1 GiB out over the host link.

  value : page traffic GB/s (in + out) with the loop driven straight through the engine's C ABI in this process, timed with
          profiling OFF (no event brackets, no in-kernel stamps). Bytes = 2 x 64 MiB x (touches of the timed region that
          found their buffer paged out) — the workload's definition, the same one the reference arm uses — NOT the engine's
          counters, which run ahead of the last touch by the prefetch window (they are reported under `engine.bytes`)
  e2e   : the same loop as an UNMODIFIED driver-API program (swap_bench) under LD_PRELOAD=libvgpu.so — the
          reference-facing boundary; h2d/d2h bytes per step are the page-in/page-out bytes the hook moved
  roofline      : vgpu_pack_tma (the engine's HBM-bound kernel) against the measured HBM copy bandwidth. In the headline
                  pass the pager moves the bytes with plain DMA between each buffer's own range and its pinned block (no
                  kernel: `roofline.headline_pass` says how many bytes went which way); the kernel runs on the latency
                  path (unpredicted misses), so it is timed in a SEPARATE profiled pass of this same process: the Zipf
                  read-mostly pass (in situ, in-kernel %globaltimer spans) and back-to-back launches on resident memory
                  (CUDA events on the launching stream). `roofline.link` = page traffic against the pinned-memcpy
                  bandwidth of this box measured in the same run: the bound that matters end to end.
  secondary     : Zipf(1.1) order, every second buffer advised read-mostly and only read: clean evictions need no copy
Multi-GPU: replicas only — every GPU enforces its own container, no collective on the data path;
torch.distributed is used for the barrier and the max-over-ranks reduction of the timing. The reference arm runs one
reference container per GPU at the same time, like ours.
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


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def pack_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE vgpu_pack_tma launch from the committed ncu --set full capture
    (profiles/r02_pack_traffic.json, written by scripts/summarize_ncu.py from the .ncu-rep), or None when no capture of
    this round is committed. Never a literal."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_pack_traffic.json")))
        return j
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "500"],
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


class Workload:
    """cfg3 (default): 8 GiB quota + 64 GiB oversubscribed. cfg5: quota = 50 % of the HBM, working set = quota + 8 GiB
    (SURVEY.md §8d cfg 5). The pinned pool holds one block per live buffer, so the whole working set must fit the budget."""

    def __init__(self, name, n_ranks, hbm_total):
        self.name = name
        if name == "cfg5":
            self.quota_mib = int(hbm_total * 0.5) // MiB // BUF_MIB * BUF_MIB
            over = 8 * GiB
        else:
            self.quota_mib = 8192
            over = 64 * GiB
        budget = host_budget_bytes(n_ranks)
        # the e2e arm runs after the in-process arm has released its pool, so each arm may use the whole per-rank budget
        while self.quota_mib * MiB + over + 2 * GiB > budget and over > 2 * GiB:
            over //= 2
        self.over = over
        self.nbuf = (self.quota_mib * MiB + over) // (BUF_MIB * MiB)

    def describe(self):
        """config.workload — one string for both arms."""
        return (f"{self.nbuf}x{BUF_MIB}MiB alloc+touch loop, {self.quota_mib}MiB gpumem quota, {self.over >> 30} GiB oversubscribed, cyclic RMW touch, "
                f"step={TOUCHES_PER_STEP} touches")

    def config(self, world):
        return {"workload": self.describe(), "inputs": "larger than L2 (each step streams 1 GiB in + 1 GiB out)",
                "parallelism": f"replicas x{world}", "name": self.name}


def measure_link(torch, barrier=None, concurrent=False):
    """Pinned-memcpy bandwidth of this box (the end-to-end roofline): each direction alone (1 GiB) and both at once (4 x
    1 GiB per direction as 32 MiB copies back to back on two streams, timed with CUDA events from a common start to the
    later of the two ends). Alone (N=1): best of 5. With several ranks every copy starts behind a barrier, so all GPUs
    pull on the host at the same time — GPUs behind one PCIe switch share its uplink — and the MEDIAN of 5 is kept: the
    ceiling the replicas actually share, not the one a lone GPU sees."""
    n, piece = 1 * GiB, 32 * MiB
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
        a, b1, b2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 4
        with torch.cuda.stream(s1):
            a.record()
        s2.wait_event(a)
        for _r in range(reps):
            for o in range(0, n, piece):
                with torch.cuda.stream(s1):
                    d1[o:o + piece].copy_(h1[o:o + piece], non_blocking=True)
                with torch.cuda.stream(s2):
                    h2[o:o + piece].copy_(d2[o:o + piece], non_blocking=True)
        with torch.cuda.stream(s1):
            b1.record()
        with torch.cuda.stream(s2):
            b2.record()
        torch.cuda.synchronize()
        got["bidir"].append(2 * reps * n / max(a.elapsed_time(b1), a.elapsed_time(b2)) / 1e6)
    del h1, h2, d1, d2
    pick = (lambda v: sorted(v)[len(v) // 2]) if concurrent else max
    return {k: pick(v) for k, v in got.items()}


def zipf_order(nbuf, count, seed=0x5EED, s=1.1):
    import numpy as np
    w = 1.0 / np.arange(1, nbuf + 1) ** s
    rng = np.random.default_rng(seed)
    return rng.choice(nbuf, size=count, p=w / w.sum()).tolist()


def run_engine_arm(torch, v, wl, steps, warmup, barrier):
    """value: the alloc+touch loop through the C ABI (no intercept layer), timed with CUDA events on the touch stream and
    profiling OFF. Then, on the same populated engine, the secondary Zipf read-mostly pass with profiling ON (in situ
    pack/unpack spans), and the final integrity check of every word."""
    L = v.lib()
    st = torch.cuda.current_stream().cuda_stream
    stp = C.c_void_p(st)
    nbuf = wl.nbuf
    sw = v.Swap(dev=torch.cuda.current_device(), resident_cap=wl.quota_mib * MiB, profile=False)
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
    torch.cuda.synchronize()                   # NOT drained: the pager's pipeline is as full at the start of the timed region as at its end
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
    d = {k: s1[k] - s0[k] for k in s1 if isinstance(s1[k], (int, float))}
    d["vmm_max_ns"] = s1["vmm_max_ns"]
    # Page traffic of the timed region BY THE WORKLOAD'S DEFINITION: every touch that found its buffer paged out needed its
    # 64 MiB brought in and — the set being full of dirty buffers — 64 MiB written back. The engine's own byte counters
    # (reported under `engine.bytes`) also include whatever the prefetch window has in flight when the region ends, which
    # the timing (an event behind the LAST TOUCH) does not wait for; the reference arm is computed the same way.
    d["paged_bytes"] = 2 * d["faults"] * nbytes
    d["touches"] = steps * TOUCHES_PER_STEP
    d["host_slabs"], d["host_slabs_local"] = s1["host_slabs"], s1["host_slabs_local"]        # totals, not deltas

    # ---- secondary pass (outside the headline timing): Zipf order, odd buffers read-mostly and only read; profiling on so
    # that the pack/unpack launches of the latency path carry event brackets and in-kernel spans
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for i in range(1, nbuf, 2):
        sw.advise_read_mostly(bufs[i])
    order = zipf_order(nbuf, max(steps, 8) * TOUCHES_PER_STEP + 4 * TOUCHES_PER_STEP)
    sw.set_profile(True)

    def ztouch(i):
        sw.acquire([bufs[i]], st)
        if i % 2:
            L.vgpu_wl_verify(bufs[i], nwords, i, touches[i], cnt.data_ptr(), stp)     # read-only touch (and an integrity check)
        else:
            L.vgpu_wl_touch(bufs[i], nwords, stp)
            touches[i] += 1
        sw.release([bufs[i]], st)

    for i in order[:4 * TOUCHES_PER_STEP]:
        ztouch(i)
    torch.cuda.synchronize()
    sw.drain()
    z0 = sw.stats()
    barrier()
    z_e0, z_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    z_e0.record()
    for i in order[4 * TOUCHES_PER_STEP:]:
        ztouch(i)
    z_e1.record()
    torch.cuda.synchronize()
    sw.drain()
    barrier()
    z_ms = z_e0.elapsed_time(z_e1)
    z1 = sw.stats()
    z = {k: z1[k] - z0[k] for k in z1 if isinstance(z1[k], (int, float))}
    z["touches"] = len(order) - 4 * TOUCHES_PER_STEP
    sw.set_profile(False)

    # integrity of everything that went through the engine (outside the timed regions)
    for i, p in enumerate(bufs):
        sw.acquire([p], st)
        L.vgpu_wl_verify(p, nwords, i, touches[i], cnt.data_ptr(), stp)
        sw.release([p], st)
    torch.cuda.synchronize()
    bad = int(cnt.item())
    sw.close()
    return ms, d, bad, z_ms, z


def pack_kernel_isolated(torch, v):
    """vgpu_pack_tma on resident memory, timed with CUDA events on the launching stream: 64 back-to-back launches of one
    32 MiB chunk each (what the engine launches; back to back so that no event->launch gap is inside the bracket) and 3
    launches of 1 GiB (steady state of the kernel). Algorithmic bytes = 2 per byte moved."""
    n = 1 * GiB
    src = torch.empty(n, dtype=torch.uint8, device="cuda").fill_(3)
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    chunk = 32 * MiB
    out = {}
    for name, segs, reps in (("chunk_32MiB", [[(src.data_ptr() + (k % 32) * chunk, dst.data_ptr() + (k % 32) * chunk, chunk)] for k in range(64)], 1),
                             ("launch_1GiB", [[(src.data_ptr(), dst.data_ptr(), n)]], 3)):
        for sg in segs[:2]:
            v.pack(sg, st)                                   # warm-up
        torch.cuda.synchronize()
        best = None
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for sg in segs:
                v.pack(sg, st)
            b.record()
            b.synchronize()
            ms = a.elapsed_time(b)
            best = ms if best is None else min(best, ms)
        moved = sum(s[2] for sg in segs for s in sg)
        out[name] = {"launches": len(segs), "avg_launch_us": round(best * 1e3 / len(segs), 2), "achieved_gbs": round(2 * moved / (best / 1e3) / 1e9, 1)}
    ok = bool(torch.equal(src, dst))
    del src, dst
    out["byte_exact"] = ok
    return out


def spawn_app(gpu, wl, steps, warmup, mode, wait_stdin, extra_args=(), ballast_mib=0, extra_env=None):
    """The unmodified driver-API app. mode: 'new' (LD_PRELOAD=libvgpu.so), 'refhook' (reference binary), 'managed'."""
    import k8s_device_plugin_b200 as v
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    cache = f"/tmp/vgpu_bench_{os.getpid()}_{gpu}_{mode}.cache"
    if os.path.exists(cache):
        os.remove(cache)
    args = [os.path.join(LIBDIR, "swap_bench"), "--cubin", CUBIN, "--buffers", str(wl.nbuf), "--mib", str(BUF_MIB),
            "--steps", str(steps * TOUCHES_PER_STEP), "--warmup", str(warmup * TOUCHES_PER_STEP), "--order", "cyclic",
            "--wait-stdin", "1" if wait_stdin else "0"] + list(extra_args)
    if mode == "new":
        env.update(v.hook_env(limit_mib=wl.quota_mib, oversubscribe=True, cache_path=cache))
        env["LIBCUDA_LOG_LEVEL"] = "1"
        args += ["--profile", "0"]      # the headline arm carries no profiling events / in-kernel stamps
    elif mode == "refhook":
        os.makedirs("/tmp/vgpulock", exist_ok=True)
        env["LD_PRELOAD"] = os.path.join(OREF, "dlsym_shim.so") + ":" + os.path.join(OREF, "libvgpu.so")
        env.update({"CUDA_OVERSUBSCRIBE": "true", "CUDA_DEVICE_MEMORY_LIMIT_0": "250000m", "CUDA_DEVICE_MEMORY_SHARED_CACHE": cache,
                    "LIBCUDA_LOG_LEVEL": "0"})
        args += ["--ballast-mib", str(ballast_mib)]
    else:
        args += ["--managed", "1", "--ballast-mib", str(ballast_mib)]
    if extra_env:
        env.update(extra_env)
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
                env.update(v.hook_env(limit_mib=8192, cache_path=cache))
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


def setup_dist(torch, nccl=True):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if nccl:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return rank, world, local, dist


def reference_arm(args):
    """The reference's swap path = CUDA UVM demand paging (cuMemAllocManaged, cuMemoryAllocate libvgpu.so@0x315da),
    executed by the NVIDIA UVM driver's fault-servicing threads on the host cores. Same workload as our arm (the same
    buffer set, the same cyclic RMW touches); a ballast allocation pins all but the quota of the GPU so that UVM sees
    the same resident budget. At N > 1 every rank runs one reference container on its own GPU at the same time, and
    the line is the aggregate over ranks like ours."""
    import torch
    rank, world, local, dist = setup_dist(torch, nccl=True)
    torch.cuda.set_device(local)
    torch.zeros(1, device="cuda")

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    bind_to_gpu_numa(torch, local)
    free_b, total_b = torch.cuda.mem_get_info(local)
    wl = Workload(args.config, world, total_b)
    ballast_mib = max(0, (free_b - wl.quota_mib * MiB - 1536 * MiB) // MiB)
    # exactly the K timed steps and W warm-up steps asked for; only absurd requests are clamped, and the line says what ran
    steps, warmup = max(1, min(args.steps, 256)), max(1, min(args.warmup, 16))
    kind, res, note = "reference", None, ""
    if os.path.exists(os.path.join(OREF, "libvgpu.so")):
        holder = None
        try:
            # the ballast lives in an UNHOOKED helper process: inside the reference-hooked process every large cuMemAlloc
            # becomes managed memory (cuMemoryAllocate allocmode 0) and would not pin anything
            henv = dict(os.environ, CUDA_VISIBLE_DEVICES=str(local), SWAP_BENCH_HOLD_MIB=str(ballast_mib))
            henv.pop("LD_PRELOAD", None)
            holder = subprocess.Popen([os.path.join(LIBDIR, "swap_bench")], env=henv, stdin=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            for line in holder.stderr:
                if line.startswith("READY"):
                    break
            p = spawn_app(local, wl, steps, warmup, "refhook", True, ballast_mib=0)
            res = finish_app(p, True, barrier, timeout=900 + 4 * steps)
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
    ok = 1.0 if res is not None else 0.0
    if dist:      # all ranks take the same path, or the barriers inside finish_app would not pair up
        t = torch.tensor([ok], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if float(t.item()) < 1.0:
            res = None
    if res is None:
        if world > 1 and ok == 1.0:
            note = "reference binary ran here but not on every rank of this box (it looks its pid up on NVML device 0, and NVML sees all GPUs outside a container); all ranks: "
        p = spawn_app(local, wl, steps, warmup, "managed", True, ballast_mib=ballast_mib)
        res = finish_app(p, True, barrier, timeout=900 + 4 * steps)
        kind = "port"
        note += "direct cuMemAllocManaged (the call the reference hook makes in allocmode 0)"
    touched = res["steps"] * BUF_MIB * MiB
    t_max, total_bytes, gbs = aggregate(dist, "cuda", res["event_ms"], 2 * touched)   # cyclic + RMW: every touch misses and evicts a dirty buffer
    mism = int(reduce_scalar(torch, dist, res["mismatches"], "SUM"))
    if rank == 0:
        line = {
            "metric": "vmem_swap_GBps", "value": round(gbs, 3), "unit": "GB/s", "impl": "reference", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(t_max / steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": wl.config(world),
            "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": f"{steps * TOUCHES_PER_STEP} touches of {BUF_MIB} MiB per GPU over the full buffer set ({wl.nbuf} managed buffers, "
                                       f"{wl.quota_mib} MiB left physical by a {ballast_mib} MiB ballast); paging by the UVM driver's fault threads on host cores + copy engines; {note}"},
            "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mismatches": mism,
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def reduce_scalar(torch, dist, x, op, device="cuda"):
    if not dist:
        return x
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


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
        line = {"metric": "vmem_swap_GBps", "value": round(value, 6), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "data": "dry-run", "total_bytes": total, "t_max_ms": t_max}
        if args.impl == "reference":       # the reference arm aggregates over ALL ranks too (N containers against N containers)
            line["impl"] = "reference"
        print(json.dumps(line))


def engine_summary(d, steps):
    """Counters of the timed region: where the bytes went and where the threads' time went (per step, rank 0)."""
    per = lambda k: round(d.get(k, 0) / 1e6 / steps, 3)
    return {
        "touches": d.get("touches"), "faults": d["faults"], "evictions": d["evictions"], "scans": d["scans"], "scan_launches": d["scan_launches"],
        "phys_creates": d["phys_creates"], "phys_reuses": d["phys_reuses"],
        "bytes": {"paged_by_definition": d.get("paged_bytes"), "counted_in": d["page_in_bytes"], "counted_out": d["page_out_bytes"], "direct_in": d["direct_in_bytes"], "direct_out": d["direct_out_bytes"], "via_pack_kernel": max(0, d["page_out_bytes"] - d["direct_out_bytes"]),
                  "via_unpack_kernel": max(0, d["page_in_bytes"] - d["direct_in_bytes"])},
        "prefetch": {"issued": d["prefetch_issued"], "hits": d["prefetch_hits"], "wasted": d["prefetch_wasted"]},
        "clean_evictions": d["clean_evictions"], "demand_waits": d["demand_waits"],
        # APPLICATION (admitting) thread per step: total inside admissions, blocked for the pager, inside VMM calls (0: the pager owns them)
        "host_ms_per_step": {"admit": per("host_admit_ns"), "wait": per("host_wait_ns"), "vmm": per("host_vmm_ns")},
        "pager_ms_per_step": {"busy": per("pager_busy_ns"), "vmm": per("pager_vmm_ns"), "unmap": per("pager_unmap_ns"), "setaccess": per("pager_setaccess_ns"),
                              "scan": per("pager_scan_ns"), "packsync": per("pager_packsync_ns"), "ringwait": per("pager_ring_ns")},
        "vmm": {"calls": d["vmm_calls"], "slow_calls_over_2ms": d["vmm_slow_calls"], "slow_ms": round(d["vmm_slow_ns"] / 1e6, 1),
                "worst_ms": round(d.get("vmm_max_ns", 0) / 1e6, 1)},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=72)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="graft")
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg5"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-cfg5", action="store_true", help="N > 1 only: do not append the cfg5 (50 %% HBM quota) pass")
    ap.add_argument("--dry-run-cpu", action="store_true", help="exercise the multi-rank plumbing with gloo and synthetic per-rank numbers (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.dry_run_cpu:
        return dry_run_cpu(args)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import k8s_device_plugin_b200 as v
    v.lib()                                           # no CUDA extension -> ImportError, never a fallback
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.zeros(1, device="cuda")
    rank, world, local, dist = setup_dist(torch)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        return reduce_scalar(torch, dist, x, op)

    numa = bind_to_gpu_numa(torch, local)
    _, total_b = torch.cuda.mem_get_info(local)
    wl = Workload(args.config, world, total_b)
    link = measure_link(torch, barrier, concurrent=world > 1)
    link_mean = {k: reduce(v_, "SUM") / world for k, v_ in link.items()}
    link_min = reduce(link["bidir"], "MIN")
    sampler = ClockSampler(local)
    sampler.start()

    # ---- arm 1: engine through the C ABI (headline pass, then the profiled Zipf read-mostly pass)
    ms, d, bad, z_ms, z = run_engine_arm(torch, v, wl, args.steps, args.warmup, barrier)
    page_bytes = d["paged_bytes"]
    t_max, total_bytes, value = aggregate(dist, "cuda", ms, page_bytes)
    z_tmax, z_total, z_value = aggregate(dist, "cuda", z_ms, z["page_in_bytes"] + z["page_out_bytes"])
    launches = d["pack_launches"] + d["unpack_launches"] + d["scan_launches"] + args.steps * TOUCHES_PER_STEP
    clocks = sampler.finish()
    local_frac = reduce(d["host_slabs_local"] / d["host_slabs"] if d.get("host_slabs") else 1.0, "MIN")
    iso = pack_kernel_isolated(torch, v) if rank == 0 else None

    # ---- arm 2: unmodified app under LD_PRELOAD (reference-facing boundary)
    p = spawn_app(local, wl, args.steps, args.warmup, "new", True)
    e2e = finish_app(p, True, barrier)
    e2e_paged = 2 * e2e["faults"] * BUF_MIB * MiB               # same definition as `value`: touches that missed x (64 MiB in + 64 MiB out)
    e2e_ms, e2e_bytes, e2e_value = aggregate(dist, "cuda", e2e["event_ms"], e2e_paged)
    e2e_bad = reduce(e2e["mismatches"], "SUM")

    # ---- N > 1: BASELINE.json configs[4] (every container at 50 % of its GPU's memory + swap) as an extra, shorter pass
    cfg5 = None
    if world > 1 and args.config == "cfg3" and not args.skip_cfg5:
        wl5 = Workload("cfg5", world, total_b)
        k5 = max(2, min(args.steps, 8))
        ms5, d5, bad5, _, _ = run_engine_arm(torch, v, wl5, k5, 3, barrier)
        t5, b5, v5 = aggregate(dist, "cuda", ms5, d5["paged_bytes"])
        bad5 = reduce(bad5, "SUM")
        cfg5 = {"value": round(v5, 3), "unit": "GB/s", "per_gpu": round(v5 / world, 3), "steps": k5, "workload": wl5.describe(),
                "frac_of_link_peak": round(v5 / world / link_mean["bidir"], 4) if link_mean["bidir"] else None, "mismatches": int(bad5)}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's swap path on this box's host cores
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "4", "--warmup", "3", "--config", args.config],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1800)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:  # reported, never silently replaced
            cpu = {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {str(e)[:200]}"}

    overhead = intercept_overhead(local) if (rank == 0 and world == 1) else None

    if rank == 0:
        peak, how = hbm_peak()
        # the pack/unpack launches of the profiled Zipf pass: CUDA-event brackets and in-kernel spans, algorithmic bytes = 2/byte
        zk_bytes = 2 * (z["pack_bytes"] + z["unpack_bytes"])
        zk_launches = z["pack_launches"] + z["unpack_launches"]
        zk_ms = z["pack_ms"] + z["unpack_ms"]
        zk_span = z["pack_span_ms"] + z["unpack_span_ms"]
        chunk = iso["chunk_32MiB"]
        traffic = pack_traffic()
        # The bound is MEASURED in this run, with every rank copying at the same time: pinned memcpy in both directions at once
        # (two streams, 4 GiB each way as 32 MiB copies, CUDA events). On a lone GPU the pager's deep queues sustain a little
        # more than this two-stream probe (frac can exceed 1: the probe is a floor of the link's capacity, the hard ceiling
        # for a workload that moves as much in as out is `symmetric_bound` = 2 x the slower one-direction peak); with 4-8
        # GPUs pulling on the shared PCIe uplinks and memory controllers the both-at-once figure is the tighter bound.
        link_peak = link_mean["bidir"]
        sym = 2.0 * min(link_mean["h2d"], link_mean["d2h"])
        link_obj = {"bound": "host-link", "achieved": round(value / world, 3), "peak": round(link_peak, 2), "unit": "GB/s",
                    "frac": round(value / world / link_peak, 4) if link_peak else None,
                    "peak_min_over_ranks": round(link_min, 2),
                    "h2d_peak": round(link_mean["h2d"], 2), "d2h_peak": round(link_mean["d2h"], 2),
                    "symmetric_bound": round(sym, 2), "frac_of_symmetric_bound": round(value / world / sym, 4) if sym else None,
                    "peak_source": ("pinned memcpy both directions at once (4 GiB each way as 32 MiB copies on two streams, CUDA events), best of 5, measured in this run"
                                    if world == 1 else
                                    f"per-GPU mean over {world} ranks copying both directions AT THE SAME TIME (barrier-started, median of 5): the host link "
                                    "the replicas share, measured in this run")}
        eng = engine_summary(d, args.steps)
        eng["pinned_slabs_on_gpu_numa_node_min_over_ranks"] = round(local_frac, 3)
        eng["numa_node_rank0"] = numa
        line = {
            "metric": "vmem_swap_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t_max / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": wl.config(world),
            "e2e": {"value": round(e2e_value, 3), "unit": "GB/s",
                    "h2d_bytes_per_step": int(e2e_paged // 2 // args.steps), "d2h_bytes_per_step": int(e2e_paged // 2 // args.steps),
                    "faults": int(e2e["faults"]), "touches": args.steps * TOUCHES_PER_STEP,
                    "hook_counted_bytes": {"page_in": int(e2e["page_in_bytes"]), "page_out": int(e2e["page_out_bytes"])},
                    "via": "LD_PRELOAD=libvgpu.so on an unmodified driver-API app (cuMemAlloc_v2/cuLaunchKernel intercept)",
                    "wall_ms": e2e["wall_ms"], "mismatches": int(e2e_bad),
                    "app_thread_vmm_ms": e2e["host_ms"]["vmm"], "vmm_slow": e2e.get("vmm_slow"), "engine": eng},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "vgpu_pack_tma",
                         # CUDA events on the launching stream, 64 back-to-back launches of one 32 MiB chunk (the engine's launch shape)
                         "achieved": chunk["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": round(chunk["achieved_gbs"] / peak, 4),
                         "peak_source": how, "launches": chunk["launches"], "avg_launch_us": chunk["avg_launch_us"], "bytes_per_launch": 2 * 32 * MiB,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": (traffic or {}).get("source", "no ncu --set full capture of this round committed yet"),
                         "steady_state_1GiB_launch": iso["launch_1GiB"], "byte_exact": iso["byte_exact"],
                         "in_situ": {"pass": "Zipf read-mostly pass of this run (latency path: pack to staging, unpack from staging), profiling on",
                                     "launches": int(zk_launches), "event_bracket_avg_us": round(zk_ms * 1e3 / max(zk_launches, 1), 2),
                                     "device_span_avg_us": round(zk_span * 1e3 / max(zk_launches, 1), 2),
                                     "device_span_achieved": round(zk_bytes / (zk_span / 1e3) / 1e9, 1) if zk_span > 0 else None,
                                     "device_span_frac": round(zk_bytes / (zk_span / 1e3) / 1e9 / peak, 4) if zk_span > 0 else None},
                         "headline_pass": {"pack_launches": int(d["pack_launches"]), "unpack_launches": int(d["unpack_launches"]),
                                           "bytes_by_plain_dma": int(d["direct_in_bytes"] + d["direct_out_bytes"]),
                                           "bytes_counted": int(d["page_in_bytes"] + d["page_out_bytes"])},
                         "note": "the headline (cyclic) pass is link-bound and moves its bytes by DMA, not by this kernel: see `link`; the kernel "
                                 "is the latency path's (unpredicted misses) and is timed here outside the headline region",
                         "link": link_obj},
            "link_roofline": link_obj,
            "secondary": {"name": "zipf(1.1) order, every second buffer advised read-mostly and only read", "value": round(z_value, 3), "unit": "GB/s",
                          "touches_per_gpu": z["touches"], "page_in_bytes": int(z["page_in_bytes"]), "page_out_bytes": int(z["page_out_bytes"]),
                          "clean_evictions": int(z["clean_evictions"]), "evictions": int(z["evictions"]), "faults": int(z["faults"]),
                          "ms_per_touch": round(z_tmax / max(z["touches"], 1), 3)},
            "cfg5": cfg5,
            "cpu_baseline": cpu,
            "intercept_overhead": overhead,
            "mismatches": bad,
            "engine": eng,
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
