"""Pins the CPU restatement (oracle/vgpu_oracle.c) to the reference: against golden fixtures produced by EXECUTING the
reference's shipped hook binary (tests/golden/make_golden.py), and — where the binary is present (build container) —
against the binary live. The reference's own tests hold no vectors for this path."""
import ctypes as C
import gzip
import hashlib
import json
import os

import pytest
from conftest import GOLDEN, OREF, have_reference, run_replay
from trace_gen import gen_trace

ORA = C.CDLL(os.path.join(OREF, "libvgpu_oracle.so"))
ORA.vo_parse_limit.restype = C.c_uint64
ORA.vo_parse_limit.argtypes = [C.c_char_p]
ORA.vo_delta.restype = C.c_int32
ORA.vo_delta.argtypes = [C.c_int32] * 6
ORA.vo_total_cuda_cores.restype = C.c_int32
ORA.vo_total_cuda_cores.argtypes = [C.c_int32, C.c_int32]


def _write(tmp_path, text):
    p = tmp_path / "trace.txt"
    p.write_text(text)
    return str(p)


def test_kat_limit_parse_matches_reference_binary():
    kat = json.load(open(os.path.join(GOLDEN, "ref_kat.json")))
    for text, want in kat["limit"]:
        assert ORA.vo_parse_limit(text.encode()) == want, text
    # SURVEY.md Appendix E vectors, now confirmed by the binary
    assert ORA.vo_parse_limit(b"8192m") == 8589934592
    assert ORA.vo_parse_limit(b"1000k") == 1024000
    assert ORA.vo_parse_limit(None) == 0


def test_kat_delta_matches_reference_binary():
    kat = json.load(open(os.path.join(GOLDEN, "ref_kat.json")))
    assert ORA.vo_total_cuda_cores(148, 2048) == kat["total_cores"] == 9699328
    for up, cur, share, want in kat["delta"]:
        assert ORA.vo_delta(148, 2048, 9699328, up, cur, share) == want, (up, cur, share)
    for up, cur, share, want in kat["delta_v100"]:
        assert ORA.vo_delta(80, 2048, 80 * 2048 * 32, up, cur, share) == want, (up, cur, share)
    # the int32 overflow on B200 that makes the share RISE while over quota (Appendix E)
    assert ORA.vo_delta(148, 2048, 9699328, 30, 100, 5000000) == 7037212


@pytest.mark.parametrize("name,n,seed,kinds", [("ref_trace_2k.out.gz", 2000, 0xB200, "A"), ("ref_trace_mixed.out.gz", 1500, 7, "AAMP")])
def test_oracle_stream_equals_golden_reference_stream(tmp_path, name, n, seed, kinds):
    want = gzip.open(os.path.join(GOLDEN, name), "rt").read()
    got = run_replay(_write(tmp_path, gen_trace(n, seed=seed, kinds=kinds)), "oracle", {"CUDA_DEVICE_MEMORY_LIMIT_0": "8192m"})
    assert got == want
    assert "rc=-1" in want and ("rc=2" in want or kinds == "A")  # the fixtures do exercise quota breaches


def test_oracle_100k_trace_hashes_equal_reference(tmp_path):
    h = json.load(open(os.path.join(GOLDEN, "ref_hashes.json")))
    big = _write(tmp_path, gen_trace(100000, seed=0xB200))
    got = run_replay(big, "oracle", {"CUDA_DEVICE_MEMORY_LIMIT_0": "8192m"})
    assert hashlib.sha256(got.encode()).hexdigest() == h["cfg2_100k_limit8192m"]
    env = dict(os.environ)
    env.pop("CUDA_DEVICE_MEMORY_LIMIT_0", None)
    got = run_replay(big, "oracle", {"CUDA_DEVICE_MEMORY_LIMIT_0": ""})
    assert hashlib.sha256(got.encode()).hexdigest() == h["cfg2_100k_unlimited"]


@pytest.mark.skipif(not have_reference(), reason="reference binary only exists in the build container")
def test_oracle_equals_reference_binary_live(tmp_path):
    t = _write(tmp_path, gen_trace(3000, seed=1234, kinds="AAAMP"))
    env = {"CUDA_DEVICE_MEMORY_LIMIT_0": "4g", "CUDA_DEVICE_MEMORY_SHARED_CACHE": str(tmp_path / "ref.cache"), "FAKE_GPU_CTX_MIB": "300"}
    assert run_replay(t, "reference", env) == run_replay(t, "oracle", env)


@pytest.mark.skipif(not have_reference(), reason="reference binary only exists in the build container")
def test_reference_deadlocks_without_the_shim_when_nvml_calls_dlsym(tmp_path):
    """Why the reference hook cannot run on driver 580 without oracle/dlsym_shim.c: libnvidia-ml resolves cu* symbols
    with dlsym() inside nvmlInit, the hook's dlsym override answers cu* names through pthread_once(preInit), and the
    hook calls nvmlInit from inside preInit. Reproduced with the fake driver (FAKE_NVML_DLSYM=1)."""
    import subprocess
    from conftest import FAKE, REF_SO, SHIM_SO
    t = _write(tmp_path, "A 0 4096\n")
    base = dict(os.environ, LD_LIBRARY_PATH=FAKE, LIBCUDA_LOG_LEVEL="0", FAKE_NVML_DLSYM="1", CUDA_DEVICE_MEMORY_SHARED_CACHE=str(tmp_path / "r.cache"))
    os.makedirs("/tmp/vgpulock", exist_ok=True)
    # a _dl_sym-only shim (what SURVEY.md §0.5 proposed): still deadlocks
    mini = tmp_path / "mini.c"
    mini.write_text('#define _GNU_SOURCE\n#include <dlfcn.h>\nvoid *_dl_sym(void *h, const char *n, void *w) { (void)w; return dlvsym(h, n, "GLIBC_2.2.5"); }\n')
    subprocess.run(["gcc", "-shared", "-fPIC", "-o", str(tmp_path / "mini.so"), str(mini), "-ldl"], check=True)
    with pytest.raises(subprocess.TimeoutExpired):
        subprocess.run([os.path.join(OREF, "trace_replay"), t], env=dict(base, LD_PRELOAD=f"{tmp_path / 'mini.so'}:{REF_SO}"),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=8)
    r = subprocess.run([os.path.join(OREF, "trace_replay"), t], env=dict(base, LD_PRELOAD=f"{SHIM_SO}:{REF_SO}"),
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=60, text=True)
    assert r.returncode == 0 and " rc=0 " in r.stdout.splitlines()[1]


def test_fuzz_limit_parser_and_delta_against_the_reference_binary():
    """3000 random limit strings and 3000 random delta() arguments evaluated INSIDE the reference binary
    (get_limit_from_env@0x40d00, delta@0x45c7b) — the CPU restatement and the product's vgpu_parse_limit agree on all."""
    import random
    import subprocess
    from conftest import FAKE, OREF, REF_SO, have_reference
    if not have_reference():
        pytest.skip("reference binary only exists in the build container")
    import k8s_device_plugin_b200 as v
    rng = random.Random(20260921)
    alphabet = "0123456789" * 4 + "abcdefxXkKmMgG" + " +-._"
    texts = []
    for _ in range(3000):
        kind = rng.random()
        if kind < 0.5:
            t = str(rng.choice([0, 1, 7, 1023, 4096, 8192, 10 ** 6, 2 ** 31, 2 ** 40, 2 ** 54, 2 ** 63, 2 ** 64 - 1, rng.randrange(10 ** 12)])) + rng.choice(["", "k", "K", "m", "M", "g", "G", "t", "b", "mm", "Mi"])
        elif kind < 0.7:
            t = rng.choice(["0x", "0X", "0"]) + "".join(rng.choice("0123456789abcdefABCDEF") for _ in range(rng.randint(1, 12))) + rng.choice(["", "k", "m", "g"])
        else:
            t = "".join(rng.choice(alphabet) for _ in range(rng.randint(1, 14)))
        if "\n" not in t and "\0" not in t:
            texts.append(t)
    deltas = [(rng.randint(0, 100), rng.randint(0, 100), rng.choice([0, 1, 1000, rng.randrange(9699328 + 1), 9699328])) for _ in range(3000)]
    feed = "".join(f"L {t}\n" for t in texts) + "".join(f"D {a} {b} {c}\n" for a, b, c in deltas)
    env = dict(os.environ, LD_LIBRARY_PATH=FAKE + ":" + os.environ.get("LD_LIBRARY_PATH", ""), LIBCUDA_LOG_LEVEL="0")
    r = subprocess.run([os.path.join(OREF, "ref_kat"), REF_SO, "--stdin"], input=feed, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    got = r.stdout.split()
    assert len(got) == len(texts) + len(deltas)
    for t, want in zip(texts, got):
        assert ORA.vo_parse_limit(t.encode()) == int(want), repr(t)
        assert v.lib().vgpu_parse_limit(t.encode()) == int(want), repr(t)
    for (a, b, c), want in zip(deltas, got[len(texts):]):
        assert ORA.vo_delta(148, 2048, 9699328, a, b, c) == int(want), (a, b, c)
