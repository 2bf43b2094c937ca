import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

ORACLE = os.path.join(ROOT, "oracle")
OREF = os.path.join(ORACLE, "_ref")
FAKE = os.path.join(OREF, "fake")
PKG = os.path.join(ROOT, "k8s-device-plugin_b200")
LIBDIR = os.path.join(PKG, "lib")
HOOK_SO = os.path.join(LIBDIR, "libvgpu.so")
CORE_SO = os.path.join(LIBDIR, "libvgpu_core.so")
REF_SO = os.path.join(OREF, "libvgpu.so")
SHIM_SO = os.path.join(OREF, "dlsym_shim.so")
CUBIN = os.path.join(PKG, "build", "vgpu_kernels.cubin")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _build_once():
    # the driver runs build() before the tests; when a developer runs pytest directly make sure the natives exist
    need = [HOOK_SO, CORE_SO, os.path.join(OREF, "oracle_replay"), os.path.join(OREF, "trace_replay"), os.path.join(OREF, "hook_stress"),
            os.path.join(FAKE, "libcuda.so.1"), os.path.join(OREF, "libvgpu_oracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


_build_once()


def run_replay(trace_path, mode, env_extra=None, fake=True, timeout=600):
    """mode: 'reference' | 'new' | 'oracle' | 'bare'. Returns stdout text of the replayer."""
    env = dict(os.environ)
    env.pop("LD_PRELOAD", None)
    env["LIBCUDA_LOG_LEVEL"] = "0"
    if fake:
        env["LD_LIBRARY_PATH"] = FAKE + ":" + env.get("LD_LIBRARY_PATH", "")
    if env_extra:
        env.update(env_extra)
    if mode == "oracle":
        cmd = [os.path.join(OREF, "oracle_replay"), trace_path]
    else:
        cmd = [os.path.join(OREF, "trace_replay"), trace_path]
        if mode == "reference":
            os.makedirs("/tmp/vgpulock", exist_ok=True)
            env["LD_PRELOAD"] = SHIM_SO + ":" + REF_SO
        elif mode == "new":
            env["LD_PRELOAD"] = HOOK_SO
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{mode} replay failed rc={r.returncode}: {r.stderr[-2000:]}")
    return r.stdout


def have_reference():
    return os.path.exists(REF_SO) and os.path.exists(SHIM_SO)


@pytest.fixture
def tmp_cache(tmp_path):
    return str(tmp_path / "vgpu.cache")
