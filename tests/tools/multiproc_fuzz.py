"""Deterministic multi-process differential fuzz: the processes of ONE container (one region file) are driven op by op
over pipes by this coordinator, so the interleaving is exactly the same under the reference binary and under the new hook —
no wall-clock schedule. Random allocations / frees / queries in up to four concurrent processes, normal exits (exit
handler), SIGKILLs (slot left behind, reclaimed by a sibling's next quota breach — rm_quitted_process) and respawns, under
a limit that is crossed often. Every output line (return code + the container-wide counter words) must be identical.
    python tests/tools/multiproc_fuzz.py <first seed> <last seed> [wide] [mixed]  (build container: needs oracle/_ref/libvgpu.so)"""
import os
import random
import signal
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import FAKE, HOOK_SO, OREF, REF_SO, SHIM_SO  # noqa: E402

M = 1 << 20


def gen(seed, nsteps=260, nproc=4, wide=False):
    """The schedule: (slot, action) pairs; action is a trace line, "spawn", "exit" or "kill"."""
    rng = random.Random(seed)
    sched, alive, live, nid = [], set(), {}, 0
    for _ in range(nsteps):
        if not alive or (len(alive) < nproc and rng.random() < 0.08):
            s = rng.choice([i for i in range(nproc) if i not in alive]); alive.add(s); live[s] = []; sched.append((s, "spawn")); continue
        s = rng.choice(sorted(alive)); r = rng.random()
        if r < 0.05 and len(alive) > 1:
            sched.append((s, "exit")); alive.discard(s)
        elif r < 0.10 and len(alive) > 1:
            sched.append((s, "kill")); alive.discard(s)
        elif r < 0.55:
            kind = rng.choice("AAAMP")
            if kind == "P":
                sched.append((s, f"P {nid} {rng.choice([100, 4096, 10000])} {rng.choice([1, 64, 500])}"))
            else:
                sched.append((s, f"{kind} {nid} {rng.choice([4096, 1 * M, 3 * M, 9 * M, 17 * M, 33 * M])}"))
            live[s].append(nid); nid += 1
        elif r < 0.80 and live[s]:
            sched.append((s, f"F {live[s].pop(rng.randrange(len(live[s])))}"))
        elif r < 0.84 and wide:
            sched.append((s, rng.choice([f"D {rng.randrange(3)}", f"B {rng.randrange(3)}", f"E {rng.randrange(8)} {rng.randrange(3)}", "N", "h 4096", "m"])))
        elif r < 0.88:
            sched.append((s, "I"))
        elif r < 0.92:
            sched.append((s, "T"))
        elif r < 0.96:
            sched.append((s, "L 1 1 1"))
        else:
            sched.append((s, f"X {hex(0x7f0000000000 + rng.randrange(1 << 30))}"))
    return sched


def run(mode, sched, cache, limit="160m", extra=None):
    env = {"PATH": os.environ.get("PATH", ""), "LD_LIBRARY_PATH": FAKE, "LIBCUDA_LOG_LEVEL": "0", "TRACE_FLUSH": "1", "FAKE_GPU_CTX_MIB": "16",
           "CUDA_DEVICE_MEMORY_LIMIT_0": limit, "CUDA_DEVICE_MEMORY_SHARED_CACHE": cache,
           "LD_PRELOAD": HOOK_SO if mode == "new" else SHIM_SO + ":" + REF_SO}
    env.update(extra or {})
    for kv in filter(None, os.environ.get("MPFUZZ_ENV", "").split(",")):      # e.g. MPFUZZ_ENV=CUDA_OVERSUBSCRIBE=true,VGPU_SWAP_LIMIT_MODE=virtual,FAKE_GPU_EXEC=1
        k, _, val = kv.partition("="); env[k] = val
    # mode "mixed": a container whose even process slots run under the new hook and odd ones under the reference binary
    preload = lambda slot: (HOOK_SO if slot % 2 == 0 else SHIM_SO + ":" + REF_SO) if mode == "mixed" else env["LD_PRELOAD"]
    os.makedirs("/tmp/vgpulock", exist_ok=True)
    procs, out = {}, []
    try:
        for s, act in sched:
            if act == "spawn":
                p = subprocess.Popen([os.path.join(OREF, "trace_replay"), "/dev/stdin"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, env=dict(env, LD_PRELOAD=preload(s)), text=True, bufsize=1)
                procs[s] = p
                out.append(f"[{s}] " + p.stdout.readline().rstrip("\n"))
            elif act == "exit":
                p = procs.pop(s); p.stdin.close(); p.wait(timeout=60); out.append(f"[{s}] exit rc={p.returncode}")
            elif act == "kill":
                p = procs.pop(s); p.send_signal(signal.SIGKILL); p.wait(timeout=60); out.append(f"[{s}] killed")
            else:
                p = procs[s]; p.stdin.write(act + "\n"); p.stdin.flush()
                line = p.stdout.readline().rstrip("\n")
                # the op counter is per process; keep the rest (rc + container-wide counters)
                out.append(f"[{s}] {act.split()[0]} " + line.split(" ", 2)[2] if line.count(" ") >= 2 else f"[{s}] {act} -> {line!r}")
    finally:
        for p in procs.values():
            try:
                p.stdin.close(); p.wait(timeout=30)
            except Exception:
                p.kill()
    return out


def compare(seed, workdir, wide=False, mixed=False):
    """wide: three GPUs with their own limits, device switches, context creation, NVML queries and host-side calls mixed in."""
    sched = gen(seed, wide=wide)
    extra = {"FAKE_GPU_COUNT": "3", "CUDA_DEVICE_MEMORY_LIMIT_1": "96m", "CUDA_DEVICE_MEMORY_LIMIT_2": "300m"} if wide else None
    res = {}
    for mode in ("new", "reference") + (("mixed",) if mixed else ()):
        cache = os.path.join(workdir, f"{mode}.cache")
        if os.path.exists(cache):
            os.remove(cache)
        res[mode] = run(mode, sched, cache, extra=extra)
    diffs = [(i, sched[i], a, b) for i, (a, b) in enumerate(zip(res["new"], res["reference"])) if a != b]
    if mixed:       # a mixed container must behave like an all-reference one as well
        diffs += [(i, sched[i], a, b) for i, (a, b) in enumerate(zip(res["mixed"], res["reference"])) if a != b]
    return sched, res, diffs


if __name__ == "__main__":
    work = tempfile.mkdtemp(prefix="vgpu_mpfuzz_")
    bad = 0
    wide, mixed = "wide" in sys.argv[3:], "mixed" in sys.argv[3:]
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        sched, res, diffs = compare(seed, work, wide, mixed)
        if diffs or len(res["new"]) != len(res["reference"]):
            bad += 1
            i, act, a, b = diffs[0] if diffs else (-1, None, len(res["new"]), len(res["reference"]))
            print("seed", seed, "first diff at step", i, act); print("  new:", a); print("  ref:", b)
    print("done, bad =", bad)
