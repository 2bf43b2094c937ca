"""GPU parity of the sm_100a kernels, called through the C ABI, against the CPU oracle (oracle/vgpu_oracle.c):
bit-exact bytes for pack/unpack, identical victim sets for the LRU scan. The reference has no counterpart for these
(its swap is UVM, cuMemoryAllocate libvgpu.so@0x315da) — the oracle restates DESIGN.md's definitions."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402
from conftest import OREF  # noqa: E402


class OSeg(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("dst_off", C.c_uint64), ("bytes", C.c_uint64)]


class OEntry(C.Structure):
    _fields_ = [("base", C.c_uint64), ("size", C.c_uint64), ("last_touch", C.c_uint64), ("state", C.c_uint32), ("host_slot", C.c_uint32)]


@pytest.fixture(scope="module")
def ora():
    o = C.CDLL(os.path.join(OREF, "libvgpu_oracle.so"))
    o.vo_select_victims.restype = C.c_int64
    o.vo_select_victims.argtypes = [C.POINTER(OEntry), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    o.vo_pack.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(OSeg), C.c_uint64]
    return o


@pytest.fixture(scope="module", autouse=True)
def ctx():
    assert torch.cuda.is_available()
    torch.zeros(1, device="cuda:0")
    assert os.path.exists(v.CORE_SO)
    v.lib()


def _oracle_pack(ora, hsrc, total, segs):
    want = np.zeros(total, dtype=np.uint8)
    arr = (OSeg * len(segs))(*[OSeg(s, d, n) for s, d, n in segs])
    ora.vo_pack(want.ctypes.data_as(C.c_void_p), hsrc.ctypes.data_as(C.c_void_p), arr, len(segs))
    return want


def _run_pack(ora, src_bytes, dst_bytes, segs, seed=0):
    rng = np.random.default_rng(seed)
    hsrc = rng.integers(0, 256, size=src_bytes, dtype=np.uint8)
    src = torch.from_numpy(hsrc).cuda()
    dst = torch.zeros(dst_bytes, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    v.pack([(src.data_ptr() + s, dst.data_ptr() + d, n) for s, d, n in segs], st)
    torch.cuda.synchronize()
    want = _oracle_pack(ora, hsrc, dst_bytes, segs)
    got = dst.cpu().numpy()
    assert np.array_equal(got, want), f"first mismatch at {int(np.argmax(got != want))}"


def test_pack_single_aligned_segment_sizes(ora):
    for n in (16, 256, 4096, 32768 - 16, 32768, 32768 + 16, 1 << 20, (1 << 20) + 4096 + 16, 64 << 20):
        _run_pack(ora, n + 64, n + 64, [(16, 32, n)], seed=n & 0xFFFF)


def test_pack_empty_and_zero_length(ora):
    _run_pack(ora, 4096, 4096, [])
    _run_pack(ora, 4096, 4096, [(0, 0, 0), (16, 16, 160), (512, 1024, 0)])


def test_pack_many_small_aligned_segments_compact(ora):
    # compaction: 400 scattered 16-byte-aligned pieces gathered back to back (more than one launch: 96 segs each)
    rng = np.random.default_rng(3)
    segs, dpos = [], 0
    for i in range(400):
        n = int(rng.integers(1, 300)) * 16
        s = i * 8192 + int(rng.integers(0, 200)) * 16
        segs.append((s, dpos, n))
        dpos += n
    _run_pack(ora, 400 * 8192 + 8192, dpos + 64, segs)


def test_pack_ragged_unaligned_segments(ora):
    segs = [(3, 5, 1), (100, 201, 12345), (70001, 90000, 33), (200000, 131072 + 7, 65536 + 5), (400001, 300001, 40000)]
    _run_pack(ora, 1 << 20, 1 << 20, segs)
    # same phase (src % 16 == dst % 16) takes the vector body with scalar head/tail
    _run_pack(ora, 1 << 20, 1 << 20, [(7, 23, 100000), (500003, 600003, 77777)])


def test_pack_mixed_aligned_and_unaligned_in_one_call(ora):
    segs = [(0, 0, 1 << 20), ((1 << 20) + 1, (1 << 20) + 4, 999), (2 << 20, 3 << 20, 524288), ((3 << 20) + 8, (2 << 20) + 24, 4096)]
    _run_pack(ora, 4 << 20, 4 << 20, segs)


def test_unpack_is_pack_with_roles_swapped_roundtrip(ora):
    # encode -> decode round trip at a size the CPU oracle would not be asked for: 1 GiB scattered into chunks and back
    n = 1 << 30
    a = torch.empty(n // 8, dtype=torch.int64, device="cuda")
    a.random_()
    staging = torch.zeros_like(a)
    back = torch.zeros_like(a)
    st = torch.cuda.current_stream().cuda_stream
    chunk = 32 << 20
    order = list(range(n // chunk))
    order = order[1::2] + order[0::2]
    v.pack([(a.data_ptr() + i * chunk, staging.data_ptr() + k * chunk, chunk) for k, i in enumerate(order)], st)
    v.pack([(staging.data_ptr() + k * chunk, back.data_ptr() + i * chunk, chunk) for k, i in enumerate(order)], st)
    torch.cuda.synchronize()
    assert torch.equal(a, back)
    assert not torch.equal(a, staging)


def _scan_case(ora, n, need, seed, touch_max=1000, resident_frac=0.7, pinned_frac=0.1):
    rng = np.random.default_rng(seed)
    tbl = (OEntry * n)()
    for i in range(n):
        r = rng.random()
        state = v.ENTRY_RESIDENT if r < resident_frac else (v.ENTRY_PAGED_OUT if r < 0.9 else v.ENTRY_FREE)
        if state == v.ENTRY_RESIDENT and rng.random() < pinned_frac:
            state |= v.ENTRY_PINNED
        tbl[i] = OEntry(0x7F0000000000 + i * (2 << 20), int(rng.integers(1, 64 << 20)), int(rng.integers(0, touch_max + 1)), state, 0)
    out = (C.c_uint32 * n)()
    freed = C.c_uint64(0)
    cnt = ora.vo_select_victims(tbl, n, need, out, C.byref(freed))
    raw = np.frombuffer(tbl, dtype=np.uint8).copy()
    d_tbl = torch.from_numpy(raw).cuda()
    got, gfreed, ins = v.victim_scan(d_tbl.data_ptr(), n, need, touch_max, torch.cuda.current_stream().cuda_stream)
    if cnt < 0:
        assert ins, "oracle says insufficient"
        want = sorted(i for i in range(n) if tbl[i].state == v.ENTRY_RESIDENT)
        assert got == want
    else:
        assert not ins
        assert got == list(out[:cnt]), (n, need, seed)
        assert gfreed == freed.value
    return cnt


def test_victim_scan_matches_oracle_small_tables(ora):
    for n, need in ((1, 1), (2, 10), (7, 1 << 20), (33, 200 << 20), (100, 64 << 20), (1000, 3 << 30)):
        for seed in range(3):
            _scan_case(ora, n, need, seed)


def test_victim_scan_single_launch_path_and_its_boundary(ora):
    # <= 8192 rows: one 1024-thread CTA does the whole scan (vgpu_victim_small); above: the multi-launch path
    for n, need in ((1024, 8 << 30), (3000, 30 << 30), (8191, 100 << 30), (8192, 1 << 30), (8193, 100 << 30), (20000, 300 << 30)):
        _scan_case(ora, n, need, seed=n)
    _scan_case(ora, 8192, 1 << 62, seed=4)                       # insufficient on the single-launch path
    _scan_case(ora, 8192, 90 << 30, seed=8, touch_max=(1 << 40) - 1)   # wide clock: 40 + 13 key bits, five digits
    _scan_case(ora, 4096, 50 << 30, seed=9, touch_max=0)         # all ties: index order only


def test_victim_scan_ties_broken_by_index(ora):
    # every candidate has the same last_touch: the order is purely by row index
    _scan_case(ora, 500, 1 << 30, seed=5, touch_max=0)
    _scan_case(ora, 500, 1 << 30, seed=6, touch_max=1)


def test_victim_scan_insufficient_and_exact_total(ora):
    assert _scan_case(ora, 64, 1 << 62, seed=1) < 0          # cannot be met: all candidates returned, flag set
    # need equal to the exact candidate total selects everything without the flag
    rng = np.random.default_rng(11)
    n = 200
    tbl = (OEntry * n)()
    total = 0
    for i in range(n):
        sz = int(rng.integers(1, 1 << 20))
        tbl[i] = OEntry(i, sz, int(rng.integers(0, 50)), v.ENTRY_RESIDENT, 0)
        total += sz
    d_tbl = torch.from_numpy(np.frombuffer(tbl, dtype=np.uint8).copy()).cuda()
    got, freed, ins = v.victim_scan(d_tbl.data_ptr(), n, total, 50)
    assert not ins and got == list(range(n)) and freed == total
    got, freed, ins = v.victim_scan(d_tbl.data_ptr(), n, total + 1, 50)
    assert ins and got == list(range(n))


def test_victim_scan_large_table_and_wide_clock(ora):
    _scan_case(ora, 200000, 500 << 30, seed=2, touch_max=(1 << 40) - 1)
    _scan_case(ora, 1 << 20, 4 << 40, seed=3, touch_max=123456789)


def test_victim_scan_persistent_kernel_boundaries_and_repeated_use(ora):
    """8192 < rows <= 148 x 8192: ONE cooperative launch, every CTA keeps its slice in registers across the digit passes
    (vgpu_victim_persist). Boundaries of the path (8193 rows = 2 CTAs, 1 212 416 = 148 full CTAs, one more row = back to the
    multi-launch path), an unmeetable need (all candidates, flag set), a wide clock (five digit passes), all ties, and many
    scans in a row on the same scanner state (the kernel must leave its histograms and barrier clean)."""
    for n, need, kw in ((8193, 60 << 30, {}), (16384, 1 << 62, {}), (50000, 700 << 30, {"touch_max": (1 << 40) - 1}),
                        (148 * 8192, 3 << 40, {"touch_max": 999}), (148 * 8192 + 1, 3 << 40, {"touch_max": 999}),
                        (30000, 200 << 30, {"touch_max": 0}), (300000, 1, {}), (300000, 2 << 40, {"resident_frac": 0.05})):
        _scan_case(ora, n, need, seed=n % 97, **kw)
    for rep in range(12):
        _scan_case(ora, 20000 + rep * 1000, (rep + 1) << 33, seed=rep)


def test_victim_scan_prefix_property_at_scale():
    # size-independent property on 1 M rows: the chosen set is a prefix of the (touch, index) order and is minimal
    n = 1 << 20
    rng = np.random.default_rng(9)
    arr = np.zeros((n, 4), dtype=np.uint64)
    arr[:, 1] = rng.integers(1, 1 << 22, size=n)
    arr[:, 2] = rng.integers(0, 1 << 30, size=n)
    state = np.where(rng.random(n) < 0.6, v.ENTRY_RESIDENT, v.ENTRY_PAGED_OUT).astype(np.uint64)
    arr[:, 3] = state
    d = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).cuda()
    need = int(arr[state == v.ENTRY_RESIDENT, 1].sum() // 3)
    got, freed, ins = v.victim_scan(d.data_ptr(), n, need, (1 << 30) - 1)
    assert not ins
    got = np.array(got)
    assert np.all(np.diff(got) > 0) and np.all(state[got] == v.ENTRY_RESIDENT)
    assert freed == int(arr[got, 1].sum()) and freed >= need
    keys = arr[:, 2].astype(object) * n + np.arange(n)
    chosen_max = max(keys[got])
    cand = np.nonzero(state == v.ENTRY_RESIDENT)[0]
    others = np.setdiff1d(cand, got)
    assert min(keys[others]) > chosen_max                      # prefix of the LRU order
    last = got[np.argmax([keys[i] for i in got])]
    assert freed - int(arr[last, 1]) < need                    # minimal
