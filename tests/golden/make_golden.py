"""Regenerates the committed golden fixtures by EXECUTING THE REFERENCE BINARY (lib/nvidia/libvgpu.so, copied to
oracle/_ref/ by oracle/Makefile) on the deterministic fake driver (oracle/fake_driver/fake_gpu.c). Only runs in the
build container (needs /root/reference); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py

Fixtures:
  ref_kat.json            delta() and get_limit_from_env() called INSIDE the reference binary (oracle/ref_kat.c)
  ref_trace_2k.out.gz     full replay stream of gen_trace(2000, seed=0xB200), limit 8192m
  ref_trace_mixed.out.gz  full stream of gen_trace(1500, seed=7, kinds="AAMP"), limit 8192m (managed + pitch paths)
  ref_exported_symbols.txt  the cu*/nvml* names the reference hook exports (nm -D), the drop-in symbol surface
  ref_hashes.json         sha256 of the replay streams of the 100 000-op cfg-2 trace (limit 8192m and unlimited) and
                          of a 20 000-op mixed trace
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import FAKE, OREF, REF_SO, SHIM_SO, run_replay  # noqa: E402
from trace_gen import gen_trace  # noqa: E402


def ref_stream(trace_text, limit):
    with tempfile.TemporaryDirectory() as td:
        tp = os.path.join(td, "t.txt")
        open(tp, "w").write(trace_text)
        env = {"CUDA_DEVICE_MEMORY_SHARED_CACHE": os.path.join(td, "ref.cache")}
        if limit:
            env["CUDA_DEVICE_MEMORY_LIMIT_0"] = limit
        return run_replay(tp, "reference", env, timeout=3600)


def main():
    assert os.path.exists(REF_SO), "reference binary not available (run `make -C oracle` in the build container)"
    env = dict(os.environ, LD_PRELOAD=SHIM_SO, LD_LIBRARY_PATH=FAKE, LIBCUDA_LOG_LEVEL="0")
    kat = subprocess.run([os.path.join(OREF, "ref_kat"), REF_SO], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True, check=True).stdout
    json.loads(kat)
    open(os.path.join(HERE, "ref_kat.json"), "w").write(kat)

    small = ref_stream(gen_trace(2000, seed=0xB200), "8192m")
    gzip.open(os.path.join(HERE, "ref_trace_2k.out.gz"), "wt").write(small)
    mixed = ref_stream(gen_trace(1500, seed=7, kinds="AAMP"), "8192m")
    gzip.open(os.path.join(HERE, "ref_trace_mixed.out.gz"), "wt").write(mixed)

    hashes = {}
    big = gen_trace(100000, seed=0xB200)
    hashes["cfg2_100k_limit8192m"] = hashlib.sha256(ref_stream(big, "8192m").encode()).hexdigest()
    hashes["cfg2_100k_unlimited"] = hashlib.sha256(ref_stream(big, None).encode()).hexdigest()
    hashes["mixed_20k_limit8192m"] = hashlib.sha256(ref_stream(gen_trace(20000, seed=7, kinds="AAMP"), "8192m").encode()).hexdigest()
    hashes["_sha256_of_reference_binary"] = hashlib.sha256(open(REF_SO, "rb").read()).hexdigest()
    json.dump(hashes, open(os.path.join(HERE, "ref_hashes.json"), "w"), indent=1)
    syms = subprocess.run(["nm", "-D", "--defined-only", REF_SO], stdout=subprocess.PIPE, text=True, check=True).stdout
    import re
    names = sorted({l.split()[2] for l in syms.splitlines() if len(l.split()) == 3 and l.split()[1] == "T" and re.match(r"^(cu[A-Z]|nvml[A-Z])", l.split()[2])})
    open(os.path.join(HERE, "ref_exported_symbols.txt"), "w").write("\n".join(names) + "\n")
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
