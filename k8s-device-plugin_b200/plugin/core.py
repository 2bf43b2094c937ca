"""ctypes binding of include/vgpu_plugin.h (native host logic in csrc/plugin_core.cc)."""
import ctypes as C
import os
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(os.path.dirname(_HERE), "lib", "libvgpu_core.so")
MAX_STR = 128


class _CD(C.Structure):
    _fields_ = [("uuid", C.c_char * MAX_STR), ("type", C.c_char * MAX_STR), ("usedmem", C.c_int32), ("usedcores", C.c_int32)]


class _ND(C.Structure):
    _fields_ = [("id", C.c_char * MAX_STR), ("count", C.c_int32), ("devmem", C.c_int32), ("devcore", C.c_int32),
                ("type", C.c_char * MAX_STR), ("numa", C.c_int32), ("health", C.c_int32)]


class _KV(C.Structure):
    _fields_ = [("key", C.c_char * MAX_STR), ("value", C.c_char * 512)]


class _Mount(C.Structure):
    _fields_ = [("container_path", C.c_char * 512), ("host_path", C.c_char * 512), ("read_only", C.c_int32)]


class _AllocIn(C.Structure):
    _fields_ = [("devices", C.POINTER(_CD)), ("n_devices", C.c_int), ("n_requested_ids", C.c_int), ("host_hook_path", C.c_char_p),
                ("pod_uid", C.c_char_p), ("container_name", C.c_char_p), ("cache_uuid", C.c_char_p), ("device_memory_scaling", C.c_double),
                ("disable_core_limit", C.c_int), ("container_sets_disable_control", C.c_int), ("license_present", C.c_int),
                ("device_list_envvar", C.c_char_p)]


class _AllocOut(C.Structure):
    _fields_ = [("envs", _KV * 32), ("n_envs", C.c_int), ("mounts", _Mount * 8), ("n_mounts", C.c_int), ("cache_host_dir", C.c_char * 512)]


@dataclass
class ContainerDevice:          # util.ContainerDevice (pkg/util/types.go:85-91)
    UUID: str
    Type: str
    Usedmem: int
    Usedcores: int


@dataclass
class NodeDevice:               # api.DeviceInfo (pkg/api/device_register.go:13-22)
    Id: str
    Count: int
    Devmem: int
    Devcore: int
    Type: str
    Numa: int
    Health: bool


class CodecError(ValueError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} missing: run __graft_entry__.build()")
        L = C.CDLL(_SO)
        for n in ("vgpu_codec_encode_node_devices", "vgpu_codec_decode_node_devices", "vgpu_codec_encode_container_devices",
                  "vgpu_codec_decode_container_devices", "vgpu_codec_encode_pod_single_device", "vgpu_codec_decode_pod_single_device",
                  "vgpu_codec_next_device_request", "vgpu_codec_erase_next_device_request", "vgpu_plugin_device_id", "vgpu_plugin_allocate"):
            getattr(L, n).restype = C.c_int
        L.vgpu_plugin_registered_mem.restype = C.c_int32
        L.vgpu_plugin_registered_mem.argtypes = [C.c_uint64, C.c_double]
        L.vgpu_plugin_registered_cores.restype = C.c_int32
        L.vgpu_plugin_registered_cores.argtypes = [C.c_double]
        _lib = L
    return _lib


def _cd_arr(devs):
    arr = (_CD * max(len(devs), 1))()
    for i, d in enumerate(devs):
        arr[i] = _CD(d.UUID.encode(), d.Type.encode(), d.Usedmem, d.Usedcores)
    return arr


def _cd_list(arr, n):
    return [ContainerDevice(arr[i].uuid.decode(), arr[i].type.decode(), arr[i].usedmem, arr[i].usedcores) for i in range(n)]


def _chk(rc, what):
    if rc != 0:
        raise CodecError(f"{what}: rc={rc}")


def encode_node_devices(devs):
    arr = (_ND * max(len(devs), 1))()
    for i, d in enumerate(devs):
        arr[i] = _ND(d.Id.encode(), d.Count, d.Devmem, d.Devcore, d.Type.encode(), d.Numa, int(d.Health))
    buf = C.create_string_buffer(65536)
    _chk(lib().vgpu_codec_encode_node_devices(arr, len(devs), buf, len(buf)), "encode_node_devices")
    return buf.value.decode()


def decode_node_devices(s):
    arr = (_ND * 64)()
    n = C.c_int(0)
    _chk(lib().vgpu_codec_decode_node_devices(s.encode(), arr, 64, C.byref(n)), "node annotations not decode successfully")
    return [NodeDevice(arr[i].id.decode(), arr[i].count, arr[i].devmem, arr[i].devcore, arr[i].type.decode(), arr[i].numa, bool(arr[i].health))
            for i in range(n.value)]


def encode_container_devices(devs):
    buf = C.create_string_buffer(65536)
    _chk(lib().vgpu_codec_encode_container_devices(_cd_arr(devs), len(devs), buf, len(buf)), "encode_container_devices")
    return buf.value.decode()


def decode_container_devices(s):
    arr = (_CD * 256)()
    n = C.c_int(0)
    _chk(lib().vgpu_codec_decode_container_devices(s.encode(), arr, 256, C.byref(n)), "pod annotation format error")
    return _cd_list(arr, n.value)


def encode_pod_single_device(containers):
    flat = [d for c in containers for d in c]
    counts = (C.c_int * max(len(containers), 1))(*[len(c) for c in containers])
    buf = C.create_string_buffer(65536)
    _chk(lib().vgpu_codec_encode_pod_single_device(_cd_arr(flat), counts, len(containers), buf, len(buf)), "encode_pod_single_device")
    return buf.value.decode()


def decode_pod_single_device(s):
    arr = (_CD * 512)()
    counts = (C.c_int * 128)()
    n = C.c_int(0)
    _chk(lib().vgpu_codec_decode_pod_single_device(s.encode(), arr, 512, counts, 128, C.byref(n)), "pod annotation format error")
    out, k = [], 0
    for c in range(n.value):
        out.append(_cd_list([arr[k + i] for i in range(counts[c])], counts[c]))
        k += counts[c]
    return out


def next_device_request(annotation):
    """GetNextDeviceRequest: (container index, devices) of the first container that still has devices to allocate."""
    arr = (_CD * 64)()
    n, idx = C.c_int(0), C.c_int(-1)
    rc = lib().vgpu_codec_next_device_request(annotation.encode(), C.byref(idx), arr, 64, C.byref(n))
    if rc == -3:
        raise LookupError("device request not found")
    _chk(rc, "next_device_request")
    return idx.value, _cd_list(arr, n.value)


def erase_next_device_request(annotation):
    buf = C.create_string_buffer(65536)
    _chk(lib().vgpu_codec_erase_next_device_request(annotation.encode(), buf, len(buf)), "erase_next_device_request")
    return buf.value.decode()


def device_id(uuid, i):
    buf = C.create_string_buffer(256)
    _chk(lib().vgpu_plugin_device_id(uuid.encode(), i, buf, len(buf)), "device_id")
    return buf.value.decode()


def registered_mem(total_bytes, scaling=1.0):
    return lib().vgpu_plugin_registered_mem(total_bytes, scaling)


def registered_cores(scaling=1.0):
    return lib().vgpu_plugin_registered_cores(scaling)


def allocate(devices, n_requested_ids, host_hook_path, pod_uid, container_name, cache_uuid=None, device_memory_scaling=1.0,
             disable_core_limit=False, container_sets_disable_control=False, license_present=False, device_list_envvar="NVIDIA_VISIBLE_DEVICES"):
    """Returns (envs dict in insertion order, mounts list of (container_path, host_path, read_only), cache_host_dir)."""
    arr = _cd_arr(devices)
    i = _AllocIn(arr, len(devices), n_requested_ids, host_hook_path.encode(), pod_uid.encode(), container_name.encode(),
                 cache_uuid.encode() if cache_uuid else None, device_memory_scaling, int(disable_core_limit),
                 int(container_sets_disable_control), int(license_present), device_list_envvar.encode())
    o = _AllocOut()
    rc = lib().vgpu_plugin_allocate(C.byref(i), C.byref(o))
    if rc == -4:
        raise ValueError("device allocate number not matched")
    _chk(rc, "allocate")
    envs = {o.envs[k].key.decode(): o.envs[k].value.decode() for k in range(o.n_envs)}
    mounts = [(o.mounts[k].container_path.decode(), o.mounts[k].host_path.decode(), bool(o.mounts[k].read_only)) for k in range(o.n_mounts)]
    return envs, mounts, o.cache_host_dir.decode()
