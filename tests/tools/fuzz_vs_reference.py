"""Differential fuzzing of the hook against the reference BINARY on the fake driver (build container only: needs
oracle/_ref/libvgpu.so): random three-GPU traces — device switches, every allocation family, frees of live / stale /
foreign / cross-device pointers, cuMemGetInfo, cuDeviceTotalMem, NVML queries, launches, context creation / destruction,
host-side allocations and registrations, pointer queries, async and VMM allocations under VGPU_REFERENCE_COVERAGE=1 — compared line by line.   python tests/tools/fuzz_vs_reference.py <first seed> <last seed>
This is how the cross-device-free crediting and the wrapping NVML free figure were found (tests/test_hook_parity_cpu.py
keeps three seeds in the suite)."""
import sys, random, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))     # tests/ (conftest)
from conftest import run_replay
def gen(seed, nops=900):
    rng = random.Random(seed)
    lines, live, nid = [], {0: [], 1: [], 2: []}, 0
    cur = 0
    made = set()
    for _ in range(nops):
        r = rng.random()
        if r < 0.08:
            cur = rng.randrange(3); lines.append(f"D {cur}")
        elif r < 0.50:
            size = rng.choice([1, 256, 4096, 1 << 20, (2 << 20) - 1, 2 << 20, (2 << 20) + 1, 5 << 20, 17 << 20, 33 << 20, 70 << 20])
            kind = rng.choice("AAAMP")
            if kind == "P":
                lines.append(f"P {nid} {rng.choice([1, 100, 4096, 10000])} {rng.choice([1, 64, 500, 3000])}")
            else:
                lines.append(f"{kind} {nid} {size}")
            live[cur].append(nid); nid += 1
        elif r < 0.80 and live[cur]:
            lines.append(f"F {live[cur].pop(rng.randrange(len(live[cur])))}")
        elif r < 0.84:
            lines.append(f"X {hex(0x7f0000000000 + rng.randrange(1 << 30))}")
        elif r < 0.88:
            lines.append("I")
        elif r < 0.90:
            lines.append("T")
        elif r < 0.92:
            lines.append("N")
        elif r < 0.93:
            lines.append("L 1 1 1")
        elif r < 0.94:
            k = rng.choice("YC"); lines.append(f"{k} {nid} {rng.choice([2<<20, 4<<20, 64<<20])}"); live.setdefault(('x',cur), []).append((k, nid)); nid += 1
        elif r < 0.95 and live.get(('x',cur)):
            k, i = live[('x',cur)].pop(); lines.append(f"{'Z' if k=='Y' else 'R'} {i}")
        elif r < 0.955:
            # context family: more retains, cuCtxCreate_v2 on any device (becomes current), destroys. Created contexts
            # are never made current again later (the reference exit()s on cuCtxSetCurrent of a duplicate create)
            k = rng.choice("BEe")
            if k == "B":
                lines.append(f"B {rng.randrange(3)}")
            elif k == "E":
                slot = rng.randrange(8); made.add(slot); cur = rng.randrange(3); lines.append(f"E {slot} {cur}")
            elif made:
                slot = rng.choice(sorted(made)); made.discard(slot); lines.append(f"e {slot}")
        elif r < 0.965:
            lines.append(rng.choice(["h 4096", "a 65536", "r 8192", "m"]))      # host-side calls: quota check only
        elif r < 0.97 and live[cur]:
            lines.append(f"Q {rng.choice(live[cur])}")                            # pointer query on a live buffer
        elif r < 0.98 and nid:
            lines.append(f"F {rng.randrange(nid)}")     # maybe double free / stale id
        else:
            other = [d for d in (0, 1, 2) if d != cur and live[d]]
            if other:
                d = rng.choice(other); lines.append(f"F {live[d].pop(rng.randrange(len(live[d])))}")
    return lines
import tempfile
D = tempfile.mkdtemp(prefix='vgpu_fuzz_')
TR, NC, RC = os.path.join(D, 't.txt'), os.path.join(D, 'n.cache'), os.path.join(D, 'r.cache')
os.makedirs('/tmp/vgpulock', exist_ok=True)
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    lines = gen(seed)
    open(TR,'w').write("\n".join(lines)+"\n")
    env={"CUDA_DEVICE_MEMORY_LIMIT_0":"96m","CUDA_DEVICE_MEMORY_LIMIT_1":"64m","CUDA_DEVICE_MEMORY_LIMIT_2":"200m","FAKE_GPU_COUNT":"3","FAKE_GPU_CTX_MIB":"16","VGPU_REFERENCE_COVERAGE":"1"}
    for kv in filter(None, os.environ.get("FUZZ_ENV", "").split(",")):      # e.g. FUZZ_ENV=CUDA_DEVICE_MEMORY_LIMIT_1=0,MEMORY_OVERRIDE=1
        k, _, val = kv.partition("="); env[k] = val
    for f in (NC, RC):
        if os.path.exists(f): os.remove(f)
    new=run_replay(TR,'new',dict(env,CUDA_DEVICE_MEMORY_SHARED_CACHE=NC)).splitlines()
    ref=run_replay(TR,'reference',dict(env,CUDA_DEVICE_MEMORY_SHARED_CACHE=RC)).splitlines()
    for i,(a,b) in enumerate(zip(new,ref)):
        if a!=b:
            print("seed", seed, "first diff at", i, "op:", lines[i-1]); print("  new:", a); print("  ref:", b); bad += 1; break
    else:
        if len(new)!=len(ref): print("seed", seed, "length", len(new), len(ref)); bad += 1
print("done, bad =", bad)
