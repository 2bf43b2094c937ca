"""The N>1 path of bench.py on CPU: two gloo ranks through the same barrier / max-over-ranks / sum-of-units code the
GPU run uses (replicas only — there is no data-path collective to test)."""
import json
import os
import socket
import subprocess
import sys

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_aggregation_with_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "10", "--warmup", "3", "--dry-run-cpu"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "rank 0 alone prints ONE JSON line"
    j = json.loads(lines[0])
    assert j["n_gpus"] == 2 and j["scaling"] == "weak"
    assert j["t_max_ms"] == 101.0 and j["total_bytes"] == 3e9          # max over ranks, sum over ranks
    assert abs(j["value"] - 3e9 / 0.101 / 1e9) < 1e-3


def test_single_rank_dry_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dry-run-cpu"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["n_gpus"] == 1 and abs(j["value"] - 10.0) < 1e-9


def test_reference_arm_aggregates_over_all_ranks_like_ours():
    """vs_reference at N > 1 must be N containers against N containers: the reference arm runs on every rank and rank 0
    prints the aggregate (max time over ranks, sum of bytes), exactly like our arm."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "10", "--warmup", "3",
           "--dry-run-cpu"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["total_bytes"] == 3e9 and j["t_max_ms"] == 101.0
