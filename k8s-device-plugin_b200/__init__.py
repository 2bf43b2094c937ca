"""B200-native vGPU enforcement library — Python binding of the C ABI (include/vgpu.h).

The product is the native code under ``csrc/`` (LD_PRELOAD hook ``lib/libvgpu.so`` and ``lib/libvgpu_core.so``);
this module only loads ``libvgpu_core.so`` with ctypes so tests, ``bench.py`` and host tools can call the C ABI.
There is no Python or CPU fallback: if the shared library is missing, importing the binding raises.
"""
import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_PKG_DIR, "lib")
CORE_SO = os.environ.get("VGPU_CORE_SO") or os.path.join(LIB_DIR, "libvgpu_core.so")   # (override: sanitizer builds, tests/tools/asan_check.sh)
HOOK_SO = os.path.join(LIB_DIR, "libvgpu.so")

MAX_DEVICES = 16
UUID_LEN = 96
REGION_SIZE = 0xC4748
REGION_MAGIC = 19920718
MEM_CONTEXT, MEM_MODULE, MEM_BUFFER = 0, 1, 2
ENTRY_FREE, ENTRY_RESIDENT, ENTRY_PAGED_OUT, ENTRY_PINNED = 0, 1, 2, 4


class DeviceMemory(C.Structure):
    _fields_ = [("context_size", C.c_uint64), ("module_size", C.c_uint64), ("buffer_size", C.c_uint64),
                ("offset", C.c_uint64), ("total", C.c_uint64)]


class ProcUsage(C.Structure):
    _fields_ = [("pid", C.c_int32), ("hostpid", C.c_int32), ("status", C.c_int32), ("_pad", C.c_int32),
                ("used", DeviceMemory * MAX_DEVICES)]


class SwapRecord(C.Structure):          # vgpu_swap_record_t (include/vgpu_region.h)
    _fields_ = [("pid", C.c_int32), ("dev", C.c_int32), ("page_out_bytes", C.c_uint64), ("page_in_bytes", C.c_uint64),
                ("evictions", C.c_uint64), ("faults", C.c_uint64), ("resident_bytes", C.c_uint64), ("live_bytes", C.c_uint64),
                ("host_bytes", C.c_uint64)]


class RegionSnapshot(C.Structure):
    _fields_ = [("initialized", C.c_int32), ("proc_num", C.c_int32), ("utilization_switch", C.c_int32),
                ("recent_kernel", C.c_int32), ("priority", C.c_int32), ("_pad", C.c_int32),
                ("device_num", C.c_uint64), ("limit", C.c_uint64 * MAX_DEVICES), ("sm_limit", C.c_uint64 * MAX_DEVICES),
                ("usage_total", C.c_uint64 * MAX_DEVICES), ("uuids", (C.c_char * UUID_LEN) * MAX_DEVICES)]


class Seg(C.Structure):
    _fields_ = [("src", C.c_uint64), ("dst", C.c_uint64), ("bytes", C.c_uint64)]


class Entry(C.Structure):
    _fields_ = [("base", C.c_uint64), ("size", C.c_uint64), ("last_touch", C.c_uint64), ("state", C.c_uint32),
                ("host_slot", C.c_uint32)]


class SwapConfig(C.Structure):
    _fields_ = [("resident_cap", C.c_uint64), ("virtual_cap", C.c_uint64), ("host_pool_cap", C.c_uint64),
                ("chunk_bytes", C.c_uint64), ("ring_slots", C.c_uint32), ("profile", C.c_uint32),
                ("prefetch_bytes", C.c_uint64), ("copy_bytes", C.c_uint64)]


class SwapStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "page_out_bytes", "page_in_bytes", "evictions", "faults", "admissions", "pack_launches", "unpack_launches",
        "scan_launches", "scans", "resident_bytes", "live_bytes", "host_bytes", "entries", "phys_creates",
        "phys_reuses", "pack_bytes", "unpack_bytes")] + [("pack_ms", C.c_double), ("unpack_ms", C.c_double)] + [(n, C.c_uint64) for n in (
        "scan_cache_hits", "host_admit_ns", "host_wait_ns", "host_vmm_ns", "pager_vmm_ns", "pager_scan_ns", "pager_packsync_ns",
        "pager_ring_ns", "pager_busy_ns", "vmm_calls")] + [
        ("pack_span_ms", C.c_double), ("unpack_span_ms", C.c_double)] + [(n, C.c_uint64) for n in (
        "direct_out_bytes", "direct_in_bytes", "prefetch_issued", "prefetch_hits", "prefetch_wasted", "demand_waits",
        "clean_evictions", "host_slabs", "host_slabs_local", "pager_unmap_ns", "pager_setaccess_ns", "pager_issue_ns",
        "pager_poll_ns", "pager_lock_ns")] + [("pager_step_ns", C.c_uint64 * 5)] + [(n, C.c_uint64) for n in (
        "vmm_slow_calls", "vmm_slow_ns", "vmm_max_ns", "inplace_uses")]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n == "pager_step_ns" else getattr(self, n)) for n, _ in self._fields_}


class LimiterStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("launches", "stamps", "groups", "busy_ns", "throttle_ns", "wall_ns")] + [
        ("limit_percent", C.c_int32), ("_pad", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "_pad"}


# every symbol include/vgpu.h declares: (restype, argtypes)
_P = C.c_void_p
_U64, _U32, _I32, _INT = C.c_uint64, C.c_uint32, C.c_int32, C.c_int
ABI = {
    "vgpu_version": (C.c_char_p, []),
    "vgpu_parse_limit": (_U64, [C.c_char_p]),
    "vgpu_region_open": (_INT, [C.c_char_p, _INT, C.POINTER(_U64), C.POINTER(_U64), _INT, C.POINTER(_P)]),
    "vgpu_region_close": (None, [_P]),
    "vgpu_region_snapshot": (_INT, [_P, C.POINTER(RegionSnapshot)]),
    "vgpu_region_proc": (_INT, [_P, _INT, C.POINTER(ProcUsage)]),
    "vgpu_region_claim": (_INT, [_P, _I32]),
    "vgpu_region_release": (None, [_P, _I32]),
    "vgpu_region_try_add": (_INT, [_P, _I32, _INT, _U64, _INT, _INT]),
    "vgpu_region_sub": (None, [_P, _I32, _INT, _U64, _INT]),
    "vgpu_region_usage": (_U64, [_P, _INT]),
    "vgpu_region_set_feedback": (_INT, [_P, _I32, _I32]),
    "vgpu_region_set_hostpid": (_INT, [_P, _I32, _I32]),
    "vgpu_region_set_uuid": (_INT, [_P, _INT, C.c_char_p]),
    "vgpu_region_swap_counters": (_INT, [_P, _INT, _P]),
    "vgpu_region_raw": (_P, [_P]),
    "vgpu_monitor_observe": (_INT, [C.POINTER(_P), _INT]),
    "vgpu_pack": (_INT, [C.POINTER(Seg), C.c_size_t, _P]),
    "vgpu_pack_config": (_INT, [_U32, _U32, _U32]),
    "vgpu_victim_scan": (_INT, [_U64, _U32, _U64, _U64, _P, C.POINTER(_U32), _U32, C.POINTER(_U32), C.POINTER(_U64),
                                C.POINTER(_INT)]),
    "vgpu_wl_fill": (_INT, [_U64, _U64, _U64, _P]),
    "vgpu_wl_touch": (_INT, [_U64, _U64, _P]),
    "vgpu_wl_verify": (_INT, [_U64, _U64, _U64, _U64, _U64, _P]),
    "vgpu_wl_touch_indirect": (_INT, [_U64, _U32, _U64, _P]),
    "vgpu_swap_create": (_INT, [_INT, C.POINTER(SwapConfig), C.POINTER(_P)]),
    "vgpu_swap_destroy": (None, [_P]),
    "vgpu_swap_alloc": (_INT, [_P, _U64, C.POINTER(_U64)]),
    "vgpu_swap_free": (_INT, [_P, _U64]),
    "vgpu_swap_acquire": (_INT, [_P, C.POINTER(_U64), _INT, _P]),
    "vgpu_swap_release": (_INT, [_P, C.POINTER(_U64), _INT, _P]),
    "vgpu_swap_release_ro": (_INT, [_P, C.POINTER(_U64), _INT, _P]),
    "vgpu_swap_advise_read_mostly": (_INT, [_P, _U64, _INT]),
    "vgpu_swap_pin": (_INT, [_P, _U64, _INT]),
    "vgpu_swap_prefetch": (_INT, [_P, _U64, _INT]),
    "vgpu_swap_stats": (_INT, [_P, C.POINTER(SwapStats)]),
    "vgpu_swap_drain": (_INT, [_P]),
    "vgpu_swap_set_profile": (_INT, [_P, _INT]),
    "vgpu_swap_table": (_INT, [_P, C.POINTER(Entry), _U32, C.POINTER(_U32)]),
    "vgpu_limiter_create": (_INT, [_INT, C.POINTER(_P)]),
    "vgpu_limiter_destroy": (None, [_P]),
    "vgpu_limiter_before_launch": (None, [_P, _P]),
    "vgpu_limiter_after_launch": (None, [_P, _P]),
    "vgpu_limiter_stats": (_INT, [_P, C.POINTER(LimiterStats)]),
    "vgpu_runtime_swap_stats": (_INT, [_INT, C.POINTER(SwapStats)]),
    "vgpu_runtime_limiter_stats": (_INT, [C.POINTER(LimiterStats)]),
    "vgpu_runtime_set_swap_profile": (_INT, [_INT, _INT]),
    "vgpu_runtime_swap_pin": (_INT, [_U64, _INT]),
    "vgpu_runtime_context_size": (_U64, []),
    "vgpu_runtime_check_memory_type": (_INT, [_U64]),
}

_lib = None


def lib():
    """The loaded C-ABI library. Raises (never falls back) when the native build is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(CORE_SO):
            raise ImportError(f"{CORE_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or `make -C {os.path.join(_PKG_DIR, 'csrc')}`); there is no Python/CPU fallback")
        L = C.CDLL(CORE_SO, mode=C.RTLD_LOCAL)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)  # AttributeError == a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class VgpuError(RuntimeError):
    def __init__(self, what, code):
        super().__init__(f"{what} failed with CUresult {code}")
        self.code = code


def _check(what, code):
    if code != 0:
        raise VgpuError(what, code)


def parse_limit(value):
    return lib().vgpu_parse_limit(value.encode() if value is not None else None)


class Region:
    """Shared-region handle (host tools / monitor side)."""

    def __init__(self, path, create=False, mem_limits=None, sm_limits=None, priority=1):
        h = _P()
        ml = (_U64 * MAX_DEVICES)(*mem_limits) if mem_limits is not None else None
        sl = (_U64 * MAX_DEVICES)(*sm_limits) if sm_limits is not None else None
        rc = lib().vgpu_region_open(path.encode(), int(create), ml, sl, priority, C.byref(h))
        if rc != 0:
            raise OSError(f"cannot open shared region {path}")
        self._h = h

    def close(self):
        if self._h:
            lib().vgpu_region_close(self._h)
            self._h = None

    def snapshot(self):
        s = RegionSnapshot()
        _check("vgpu_region_snapshot", lib().vgpu_region_snapshot(self._h, C.byref(s)))
        return s

    def proc(self, i):
        p = ProcUsage()
        _check("vgpu_region_proc", lib().vgpu_region_proc(self._h, i, C.byref(p)))
        return p

    def claim(self, pid):
        return lib().vgpu_region_claim(self._h, pid)

    def release(self, pid):
        lib().vgpu_region_release(self._h, pid)

    def try_add(self, pid, dev, nbytes, kind=MEM_BUFFER, enforce=True):
        return bool(lib().vgpu_region_try_add(self._h, pid, dev, nbytes, kind, int(enforce)))

    def sub(self, pid, dev, nbytes, kind=MEM_BUFFER):
        lib().vgpu_region_sub(self._h, pid, dev, nbytes, kind)

    def usage(self, dev):
        return lib().vgpu_region_usage(self._h, dev)

    def set_uuid(self, dev, uuid):
        lib().vgpu_region_set_uuid(self._h, dev, uuid.encode())

    def swap_counters(self, dev):
        """Swap-engine counters of the container on `dev` (extension block of the region file), or None when the file has
        none (written by the reference hook)."""
        rec = SwapRecord()
        if lib().vgpu_region_swap_counters(self._h, dev, C.byref(rec)) != 0:
            return None
        return {k: getattr(rec, k) for k, _ in SwapRecord._fields_ if k not in ("pid", "dev")} | {"processes": rec.pid}

    def set_feedback(self, recent_kernel=None, utilization_switch=None):
        keep = -(2 ** 31)
        lib().vgpu_region_set_feedback(self._h, keep if recent_kernel is None else recent_kernel,
                                       keep if utilization_switch is None else utilization_switch)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def monitor_observe(regions):
    """One feedback pass of the node monitor over Region objects (reference Observe, feedback.go:197-255)."""
    arr = (_P * max(len(regions), 1))(*[r._h for r in regions])
    return lib().vgpu_monitor_observe(arr, len(regions))


def pack(segments, stream=0):
    """segments: iterable of (src_devptr, dst_devptr, nbytes); enqueues the copies on `stream` (a CUstream handle)."""
    segs = list(segments)
    arr = (Seg * len(segs))(*[Seg(s, d, n) for s, d, n in segs])
    _check("vgpu_pack", lib().vgpu_pack(arr, len(segs), _P(stream)))


def victim_scan(d_table, n, need, max_touch, stream=0):
    """Returns (sorted row indices, freed bytes, insufficient flag)."""
    out = (_U32 * max(n, 1))()
    cnt, freed, ins = _U32(0), _U64(0), _INT(0)
    _check("vgpu_victim_scan", lib().vgpu_victim_scan(d_table, n, need, max_touch, _P(stream), out, max(n, 1),
                                                      C.byref(cnt), C.byref(freed), C.byref(ins)))
    return list(out[:cnt.value]), freed.value, bool(ins.value)


class Swap:
    """The swap engine through the C ABI (the same object the hook creates under CUDA_OVERSUBSCRIBE=true)."""

    def __init__(self, dev=0, resident_cap=0, virtual_cap=0, host_pool_cap=0, chunk_bytes=0, ring_slots=0, profile=False,
                 prefetch_bytes=0, copy_bytes=0):
        """prefetch_bytes: 0 = default window, None = prefetch off."""
        cfg = SwapConfig(resident_cap, virtual_cap, host_pool_cap, chunk_bytes, ring_slots, int(profile),
                         (1 << 64) - 1 if prefetch_bytes is None else prefetch_bytes, copy_bytes)
        h = _P()
        _check("vgpu_swap_create", lib().vgpu_swap_create(dev, C.byref(cfg), C.byref(h)))
        self._h = h

    def alloc(self, nbytes):
        p = _U64(0)
        _check("vgpu_swap_alloc", lib().vgpu_swap_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        _check("vgpu_swap_free", lib().vgpu_swap_free(self._h, ptr))

    def acquire(self, ptrs, stream=0):
        a = (_U64 * len(ptrs))(*ptrs)
        _check("vgpu_swap_acquire", lib().vgpu_swap_acquire(self._h, a, len(ptrs), _P(stream)))

    def release(self, ptrs, stream=0):
        a = (_U64 * len(ptrs))(*ptrs)
        _check("vgpu_swap_release", lib().vgpu_swap_release(self._h, a, len(ptrs), _P(stream)))

    def release_ro(self, ptrs, stream=0):
        """release() for work that only read the buffers: they stay clean."""
        a = (_U64 * len(ptrs))(*ptrs)
        _check("vgpu_swap_release_ro", lib().vgpu_swap_release_ro(self._h, a, len(ptrs), _P(stream)))

    def advise_read_mostly(self, ptr, on=True):
        _check("vgpu_swap_advise_read_mostly", lib().vgpu_swap_advise_read_mostly(self._h, ptr, int(on)))

    def prefetch(self, ptr, to_device=True):
        """cuMemPrefetchAsync's twin: queue a page-in (to_device) or make the buffer the first victim (to the host)."""
        _check("vgpu_swap_prefetch", lib().vgpu_swap_prefetch(self._h, ptr, int(to_device)))

    def pin(self, ptr, on=True):
        _check("vgpu_swap_pin", lib().vgpu_swap_pin(self._h, ptr, int(on)))

    def stats(self):
        s = SwapStats()
        _check("vgpu_swap_stats", lib().vgpu_swap_stats(self._h, C.byref(s)))
        return s.as_dict()

    def drain(self):
        _check("vgpu_swap_drain", lib().vgpu_swap_drain(self._h))

    def set_profile(self, on=True):
        _check("vgpu_swap_set_profile", lib().vgpu_swap_set_profile(self._h, int(on)))

    def table(self):
        n = _U32(0)
        lib().vgpu_swap_table(self._h, None, 0, C.byref(n))
        arr = (Entry * max(n.value, 1))()
        _check("vgpu_swap_table", lib().vgpu_swap_table(self._h, arr, n.value, C.byref(n)))
        return list(arr[:n.value])

    def close(self):
        if self._h:
            lib().vgpu_swap_destroy(self._h)
            self._h = None


def hook_env(limit_mib=None, sm_limit=None, oversubscribe=False, cache_path=None, extra=None):
    """Environment a container gets from Allocate (reference server.go:343-361) for running a process under the hook."""
    env = {"LD_PRELOAD": HOOK_SO}
    if limit_mib is not None:
        env["CUDA_DEVICE_MEMORY_LIMIT_0"] = f"{int(limit_mib)}m"
    if sm_limit is not None:
        env["CUDA_DEVICE_SM_LIMIT"] = str(int(sm_limit))
    if oversubscribe:
        env["CUDA_OVERSUBSCRIBE"] = "true"
    if cache_path:
        env["CUDA_DEVICE_MEMORY_SHARED_CACHE"] = cache_path
    if extra:
        env.update(extra)
    return env
