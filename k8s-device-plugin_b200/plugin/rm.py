"""The device-plugin shell around Allocate — SURVEY.md §8(f) #3.

Reference: cmd/device-plugin/nvidia/main.go:170-245 (start: FS watcher on the kubelet socket, SIGHUP restart, 30 s retry),
vgpucfg.go:80-108 (per-node overrides from /config/config.json), nvinternal/rm/health.go:42-190 (Xid health check over an
NVML event set), nvinternal/plugin/register.go:96-200 (getApiDevices / RegistrInAnnotation / WatchAndRegister: 30 s on
success, 5 s on error).

NVML is reached through nvidia-ml-py (the Python binding of the same libnvidia-ml.so.1 the Go code uses via go-nvml);
everything above it takes small interfaces so the CPU tests drive it with scripted events.
"""
import json
import os
import signal
import threading
import time
from dataclasses import dataclass

from . import api
from .server import GpuDevice

ENV_DISABLE_HEALTHCHECKS = "DP_DISABLE_HEALTHCHECKS"      # health.go:34
MAX_SUCCESSIVE_EVENT_ERRORS = 3
# health.go:66-72 — application errors: the GPU itself is still healthy
APPLICATION_ERROR_XIDS = (13, 31, 43, 45, 68)
EVENT_XID_CRITICAL = 0x8                                   # nvmlEventTypeXidCriticalError
EVENT_DOUBLE_BIT_ECC = 0x2
EVENT_SINGLE_BIT_ECC = 0x1


def additional_xids(value):
    """getAdditionalXids (health.go:192-213): comma-separated, invalid entries ignored."""
    out = []
    for part in (value or "").split(","):
        part = part.strip()
        if part.isdigit():
            out.append(int(part))
    return out


def skipped_xids(env_value):
    """None = health checks disabled entirely ("all" / contains "xids"); otherwise the set of Xids that do not mark a GPU
    unhealthy."""
    v = (env_value or "").lower()
    if v == "all":
        v = "xids"
    if "xids" in v:
        return None
    return set(APPLICATION_ERROR_XIDS) | set(additional_xids(v))


@dataclass
class Event:
    """What eventSet.Wait returns, reduced to what checkHealth reads."""
    etype: int = 0
    xid: int = 0
    uuid: str | None = None        # None: the device UUID could not be determined
    error: str | None = None       # "timeout" or an NVML error string


class EventSource:
    def register(self, uuid):      # -> None, or an error string (device then marked unhealthy)
        return None

    def wait(self, timeout_ms):
        raise NotImplementedError

    def close(self):
        pass


def check_health(stop, devices, unhealthy, source, disable=None, wait_ms=5000):
    """checkHealth (health.go:42-190). stop: threading.Event; devices: GpuDevice list; unhealthy: callable(device)."""
    skipped = skipped_xids(os.environ.get(ENV_DISABLE_HEALTHCHECKS) if disable is None else disable)
    if skipped is None:
        return
    by_uuid = {}
    for d in devices:
        by_uuid[d.ID] = d
        err = source.register(d.ID)
        if err is not None:
            unhealthy(d)
    try:
        while not stop.is_set():
            e = source.wait(wait_ms)
            if e.error == "timeout":
                continue
            if e.error is not None:                      # any other wait error: every device is suspect
                for d in devices:
                    unhealthy(d)
                continue
            if e.etype != EVENT_XID_CRITICAL:
                continue
            if e.xid in skipped:
                continue
            if e.uuid is None:
                for d in devices:
                    unhealthy(d)
                continue
            d = by_uuid.get(e.uuid)
            if d is None:
                continue
            unhealthy(d)
    finally:
        source.close()


class NvmlEventSource(EventSource):
    def __init__(self):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.set = pynvml.nvmlEventSetCreate()

    def register(self, uuid):
        nv = self.nv
        try:
            h = nv.nvmlDeviceGetHandleByUUID(uuid)
            supported = nv.nvmlDeviceGetSupportedEventTypes(h)
            nv.nvmlDeviceRegisterEvents(h, (EVENT_XID_CRITICAL | EVENT_DOUBLE_BIT_ECC | EVENT_SINGLE_BIT_ECC) & supported, self.set)
        except nv.NVMLError as e:
            return str(e)
        return None

    def wait(self, timeout_ms):
        nv = self.nv
        try:
            d = nv.nvmlEventSetWait_v2(self.set, timeout_ms)
        except nv.NVMLError_Timeout:
            return Event(error="timeout")
        except nv.NVMLError as e:
            return Event(error=str(e))
        try:
            uuid = nv.nvmlDeviceGetUUID(d.device)
            uuid = uuid.decode() if isinstance(uuid, bytes) else uuid
        except nv.NVMLError:
            uuid = None
        return Event(etype=int(d.eventType), xid=int(d.eventData), uuid=uuid)

    def close(self):
        try:
            self.nv.nvmlEventSetFree(self.set)
            self.nv.nvmlShutdown()
        except Exception:
            pass


def parse_nvidia_numa_info(idx, topo):
    """parseNvidiaNumaInfo (register.go:45-93) on the text of `nvidia-smi topo -m`, quirks included: only lines containing
    "GPU" are looked at; the header must be the very first line (its "NUMA Affinity" column index is used for the rows,
    after collapsing double tabs); a row belongs to GPU idx when its first word CONTAINS the decimal idx (so 1 also matches
    GPU10..GPU19 — the last matching row wins); "N/A" -> 0 at once; anything else is strconv.Atoi (ValueError here)."""
    result, col = 0, 0
    for index, line in enumerate(topo.split("\n")):
        if "GPU" not in line:
            continue
        words = line.replace("\t\t", "\t").split("\t")
        if index == 0:
            for ci, header in enumerate(words):
                if "NUMA Affinity" in header:
                    col = ci
            continue
        if str(idx) in words[0]:
            if words[col] == "N/A":           # IndexError where Go would panic on a short row
                return 0
            result = _go_atoi(words[col])
    return result


def _go_atoi(text):
    """strconv.Atoi: optional sign, decimal digits only (no spaces, no underscores)."""
    body = text[1:] if text[:1] in "+-" else text
    if not body or not all("0" <= ch <= "9" for ch in body):
        raise ValueError(f'strconv.Atoi: parsing "{text}": invalid syntax')
    return int(text)


def numa_node_of(idx, bus_id, run=None):
    """NUMA node of GPU idx. sysfs first (no process to start, no text to parse); where the PCI device is not visible in
    sysfs (some container runtimes), the reference's way: `nvidia-smi topo -m` (register.go:35-43)."""
    bus = bus_id.lower()
    for cand in (bus, bus[4:] if len(bus) > 12 else bus):      # NVML pads the domain to 8 hex digits, sysfs uses 4
        try:
            return max(int(open(f"/sys/bus/pci/devices/{cand}/numa_node").read()), 0)
        except (OSError, ValueError):
            continue
    try:
        import subprocess
        out = (run or (lambda: subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                              timeout=30).stdout))()
        return parse_nvidia_numa_info(idx, out)
    except (OSError, ValueError, IndexError, Exception):
        return 0


def nvml_devices():
    """getApiDevices' NVML walk (register.go:96-162): every GPU with its UUID, memory, model and NUMA node."""
    import pynvml as nv
    nv.nvmlInit()
    try:
        out = []
        for i in range(nv.nvmlDeviceGetCount()):
            h = nv.nvmlDeviceGetHandleByIndex(i)
            s = lambda x: x.decode() if isinstance(x, bytes) else x
            numa = numa_node_of(i, s(nv.nvmlDeviceGetPciInfo(h).busId))
            out.append(GpuDevice(ID=s(nv.nvmlDeviceGetUUID(h)), Health=api.HEALTHY, TotalMemory=int(nv.nvmlDeviceGetMemoryInfo(h).total),
                                 Model=s(nv.nvmlDeviceGetName(h)), Numa=numa))
        return out
    finally:
        nv.nvmlShutdown()


# ---------------------------------------------------------------- per-node overrides (vgpucfg.go:80-108)
def read_node_config(path, node_name, split_count=2, memory_scaling=1.0, cores_scaling=1.0):
    """/config/config.json: {"nodeconfig": [{"name", "devicememoryscaling", "devicecorescaling", "devicesplitcount"}]};
    positive values for the matching node override the flags. Returns (split_count, memory_scaling, cores_scaling)."""
    try:
        cfg = json.load(open(path))
    except (OSError, ValueError):
        return split_count, memory_scaling, cores_scaling
    for val in cfg.get("nodeconfig") or []:
        if val.get("name") == node_name:
            if (val.get("devicememoryscaling") or 0) > 0:
                memory_scaling = float(val["devicememoryscaling"])
            if (val.get("devicecorescaling") or 0) > 0:
                cores_scaling = float(val["devicecorescaling"])
            if (val.get("devicesplitcount") or 0) > 0:
                split_count = int(val["devicesplitcount"])
    return split_count, memory_scaling, cores_scaling


# ---------------------------------------------------------------- registration loop (register.go:164-200)
def register_in_annotation(plugin, kube, now=None):
    annos = plugin.node_annotations(now=time.ctime() if now is None else now)
    kube.get_node(plugin.node_name)                       # util.GetNode first: a missing node is an error before the patch
    kube.patch_node_annotations(plugin.node_name, annos)
    return annos


def watch_and_register(plugin, kube, stop, ok_interval=30.0, err_interval=5.0, log=None):
    while not stop.is_set():
        try:
            register_in_annotation(plugin, kube)
            delay = ok_interval
        except Exception as e:
            if log:
                log(f"Failed to register annotation: {e}")
            delay = err_interval
        stop.wait(delay)


# ---------------------------------------------------------------- restart loop (main.go:170-245)
class PluginManager:
    """start(): serve + register every plugin that has devices; a failure to reach the kubelet schedules a retry in 30 s;
    the kubelet socket being re-created (kubelet restart) or SIGHUP restarts everything; other signals stop."""

    def __init__(self, make_plugins, kubelet_socket, retry_s=30.0, poll_s=0.5):
        self.make_plugins = make_plugins
        self.kubelet_socket = kubelet_socket
        self.retry_s = retry_s
        self.poll_s = poll_s
        self.plugins = []
        self.restarts = 0
        self._events = []
        self._cv = threading.Condition()

    def notify(self, what):                 # "restart" | "exit"
        with self._cv:
            self._events.append(what)
            self._cv.notify_all()

    def install_signal_handlers(self):
        signal.signal(signal.SIGHUP, lambda *_: self.notify("restart"))
        for s in (signal.SIGINT, signal.SIGTERM, signal.SIGQUIT):
            signal.signal(s, lambda *_: self.notify("exit"))

    def _socket_id(self):
        try:
            st = os.stat(self.kubelet_socket)
            return (st.st_ino, st.st_ctime_ns)
        except OSError:
            return None

    def _start_plugins(self):
        self.plugins = list(self.make_plugins())
        started = 0
        for p in self.plugins:
            if not p.devices:
                continue
            try:
                p.Start()
            except Exception:
                return True               # could not contact the kubelet: retry all of them later
            started += 1
        return False

    def _stop_plugins(self):
        for p in self.plugins:
            p.Stop()

    def run(self, max_restarts=None):
        restarting = False
        while True:
            if restarting:
                self._stop_plugins()
                self.restarts += 1
                if max_restarts is not None and self.restarts > max_restarts:
                    return
            retry = self._start_plugins()
            restarting = True
            deadline = time.monotonic() + self.retry_s if retry else None
            sock = self._socket_id()
            while True:
                with self._cv:
                    if not self._events:
                        self._cv.wait(self.poll_s)
                    ev = self._events.pop(0) if self._events else None
                if ev == "exit":
                    self._stop_plugins()
                    return
                if ev == "restart":
                    break
                if deadline is not None and time.monotonic() >= deadline:
                    break
                now = self._socket_id()
                if now is not None and now != sock:      # inotify Create on kubelet.sock in the reference
                    break
                if now is None:
                    sock = None
