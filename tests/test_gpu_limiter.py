"""The device-timestamped token bucket through the C ABI: achieved duty cycle vs quota on long and on short kernels
(the reference's quota semantics — CUDA_DEVICE_SM_LIMIT percent — not its delta() arithmetic)."""
import ctypes as C
import time

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402


def _run(percent, nbytes, seconds=3.0):
    torch.zeros(1, device="cuda")
    L = v.lib()
    h = C.c_void_p()
    assert L.vgpu_limiter_create(percent, C.byref(h)) == 0
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    stp = C.c_void_p(st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        L.vgpu_limiter_before_launch(h, stp)
        assert L.vgpu_wl_touch(buf.data_ptr(), nbytes // 8, stp) == 0
        L.vgpu_limiter_after_launch(h, stp)
        n += 1
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    s = v.LimiterStats()
    assert L.vgpu_limiter_stats(h, C.byref(s)) == 0
    L.vgpu_limiter_destroy(h)
    return n, wall, s.as_dict()


def _kernel_ms(nbytes):
    """device time of one touch kernel, by CUDA events, no limiter involved"""
    buf = torch.zeros(nbytes // 8, dtype=torch.int64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        v.lib().vgpu_wl_touch(buf.data_ptr(), nbytes // 8, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        v.lib().vgpu_wl_touch(buf.data_ptr(), nbytes // 8, st)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 10


def test_long_kernels_are_held_to_30_percent():
    t_kernel = _kernel_ms(4 << 30)          # ~1.5 ms per touch of 4 GiB
    n, wall, s = _run(30, 4 << 30)
    duty = s["busy_ns"] / 1e9 / wall
    assert 0.24 <= duty <= 0.36, (duty, n, s)
    assert s["throttle_ns"] > 0.4 * wall * 1e9
    # the %globaltimer stamps must agree with CUDA events on what a kernel costs (independent calibration) ...
    assert abs(s["busy_ns"] / 1e6 / n - t_kernel) < 0.25 * t_kernel, (s["busy_ns"] / 1e6 / n, t_kernel)
    # ... and so must the launch count: n kernels of t_kernel each inside `wall` seconds
    assert 0.22 <= n * t_kernel / 1e3 / wall <= 0.38, (n, t_kernel, wall)


def test_short_kernels_are_held_to_50_percent_with_amortised_stamps():
    n, wall, s = _run(50, 8 << 20)          # ~10 us kernels: stamps must be amortised over groups
    duty = s["busy_ns"] / 1e9 / wall
    assert 0.35 <= duty <= 0.6, (duty, n, s)
    assert s["stamps"] < 1.2 * s["launches"]


def test_limit_100_is_a_no_op():
    n, wall, s = _run(100, 64 << 20, seconds=1.0)
    assert s["stamps"] == 0 and s["throttle_ns"] == 0
