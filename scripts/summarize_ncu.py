"""Turns gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and prof_*.ncu-rep into the tracked summaries
under profiles/. Usage: python scripts/summarize_ncu.py <round tag>"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def launches(name="launches.csv", title=None):
    p = os.path.join(OUT, name)
    if not os.path.exists(p):
        return None
    text = open(p).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values()) or 1
    lines = [title or f"# ncu launch list ({tag}) — scripts/ncu_target.py (engine-only swap loop, 32x64 MiB under 1 GiB)", "",
             "cold-cache, serialised per-launch times: compare SHARES, not absolutes", "",
             "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.2f} | {ns / total:.1%} |")
    return "\n".join(lines) + "\n"


def full(rep, kernel):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    r = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return f"could not read {rep}: {r.stderr[:300]}\n"
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg"]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full: {kernel} ({tag})", "", "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |",
             "|---|---|" + "---:|" * (len(rows) - 2)]
    for m in want:
        if m in idx:
            lines.append(f"| {m} | {units[idx[m]]} | " + " | ".join(row[idx[m]] for row in rows[2:]) + " |")
    return "\n".join(lines) + "\n"


a = launches()
if a:
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write(a)
    print(a)
c = launches("launches_bench.csv", f"# ncu launch list ({tag}) — `python bench.py --steps 2 --warmup 3`, launches inside the NVTX range \"timed\" only "
             "(`ncu --nvtx --nvtx-include \"timed/\" --metrics gpu__time_duration.sum --clock-control none`)")
if c:
    open(os.path.join(ROOT, "profiles", f"{tag}_launches_bench.md"), "w").write(c)
    print(c)
b = full("prof_pack.ncu-rep", "vgpu_pack_tma")
if b:
    open(os.path.join(ROOT, "profiles", f"{tag}_pack_tma_full.md"), "w").write(b)
    print(b)
    # bench.py's roofline.traffic: DRAM bytes of ONE launch, read from the capture (never typed in)
    import json
    r = subprocess.run(["ncu", "-i", os.path.join(OUT, "prof_pack.ncu-rep"), "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]

    def to_bytes(row, m):
        val = float(row[idx[m]].replace(",", ""))
        return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[m]], 1)

    per = []
    for row in rows[2:]:
        dur_unit = units[idx["gpu__time_duration.sum"]]
        dur = float(row[idx["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(dur_unit, 1)
        per.append({"duration_us": round(dur, 2), "dram_read_bytes": int(to_bytes(row, "dram__bytes_read.sum")), "dram_write_bytes": int(to_bytes(row, "dram__bytes_write.sum"))})
    small = [p for p in per if p["duration_us"] < 100]
    big = [p for p in per if p["duration_us"] >= 100]
    out = {"source": f"profiles/{tag}_pack_tma_full.md (ncu --set full --clock-control none, scripts/ncu_pack_target.py)", "launches": per}
    if small:
        out["dram_bytes_per_launch"] = small[-1]["dram_read_bytes"] + small[-1]["dram_write_bytes"]
        out["algorithmic_bytes_per_launch"] = 2 * 32 * 1024 * 1024
        out["note"] = "32 MiB chunk launch: reads 33.55 MB from DRAM; most of the 33.55 MB it writes are still dirty in the 126 MB L2 when the kernel ends"
    if big:
        out["launch_1GiB"] = {"dram_bytes": big[-1]["dram_read_bytes"] + big[-1]["dram_write_bytes"], "algorithmic_bytes": 2 * 1024 ** 3, "duration_us": big[-1]["duration_us"]}
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_pack_traffic.json"), "w"), indent=1)
    print(json.dumps(out))
