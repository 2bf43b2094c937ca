"""Reads the '[vgpu-b200 trace]' JSON lines (VGPU_SWAP_TRACE) and prints, per missing admission, the host phases and
the busy/idle pattern of the two DMA queues. Usage: python scripts/analyze_trace.py gpurun_out/trace1.err"""
import json
import sys

recs = [json.loads(l.split("] ", 1)[1]) for l in open(sys.argv[1]) if l.startswith("[vgpu-b200 trace]")]
if not recs:
    sys.exit("no trace lines")
t0 = recs[0]["host_us"]["begin"]
print(f"{'i':>3} {'begin':>8} | host: {'packs':>6} {'stage':>6} {'sync+unmap':>10} {'map+unpack':>10} {'total':>6} | d2h chunks (start-end)            | h2d chunks (start-end)")
for r in recs:
    h = r["host_us"]
    ph = (h["packs"] - h["begin"], h["staged"] - h["packs"], h["unmapped"] - h["staged"], h["end"] - h["unmapped"], h["end"] - h["begin"])
    d2h = " ".join(f"{a - t0:.0f}-{b - t0:.0f}" for a, b in r["d2h_us"])
    h2d = " ".join(f"{a - t0:.0f}-{b - t0:.0f}" for a, b in r["h2d_us"])
    print(f"{r['i']:>3} {h['begin'] - t0:>8.0f} | {ph[0]:>6.0f} {ph[1]:>6.0f} {ph[2]:>10.0f} {ph[3]:>10.0f} {ph[4]:>6.0f} | {d2h:<34} | {h2d}")
for name in ("d2h_us", "h2d_us"):
    iv = sorted((a, b) for r in recs for a, b in r[name])
    busy = sum(b - a for a, b in iv)
    span = iv[-1][1] - iv[0][0]
    gaps = [iv[i + 1][0] - iv[i][1] for i in range(len(iv) - 1)]
    print(f"{name}: busy {busy / span:.1%} of {span / 1e3:.1f} ms; chunk avg {busy / len(iv):.0f} us; gaps avg {sum(gaps) / len(gaps):.0f} us max {max(gaps):.0f} us")
period = (recs[-1]["host_us"]["begin"] - recs[0]["host_us"]["begin"]) / (len(recs) - 1)
print(f"admission period {period:.0f} us")
