"""Scheduler extender + admission webhook — SURVEY.md §8(f) #1, the caller that writes the annotations Allocate consumes.

Reference: pkg/scheduler/scheduler.go (Scheduler, onAddPod:68, RegisterFromNodeAnnotatons:123, getNodesUsage:250, Bind:315,
Filter:361), nodes.go / pods.go (the two managers), webhook.go:46-83, routes/route.go:41-134, cmd/scheduler/main.go:48-84
(flags and routes), cmd/scheduler/metrics.go:55-200 (metric names), pkg/device/nvidia/device.go:59-177
(MutateAdmission / GenerateResourceRequests), pkg/k8sutil/pod.go:26-41 (Resourcereqs), pkg/util/nodelock/nodelock.go.

The scoring itself (device ordering, fit, score) is native: csrc/sched_core.cc behind include/vgpu_sched.h. This module
is the transport and the cluster bookkeeping around it; pods and nodes are plain Kubernetes JSON objects (dicts), the
wire format kube-scheduler and the apiserver speak. Cluster access goes through the small KubeClient interface below so
tests run against InMemoryKube; a deployment plugs in a client for the real apiserver.
"""
import base64
import ctypes as C
import json
import math
import re
import threading
import time
from dataclasses import dataclass, field
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from . import core
from .server import (ALLOCATED, ASSIGNED_NODE, BIND_ALLOCATING, BIND_PHASE, BIND_TIME, HANDSHAKE, NVIDIA_GPU_DEVICE, REGISTER,
                     RESOURCE_CORES, RESOURCE_MEM, RESOURCE_MEM_PERCENTAGE, RESOURCE_NAME, TO_ALLOCATE)

ASSIGNED_TIME = "4pd.io/vgpu-time"                 # util.AssignedTimeAnnotations (types.go:27)
NODE_LOCK_TIME = "4pd.io/mutex.lock"               # nodelock.NodeLockTime (nodelock.go:14)
MAX_LOCK_RETRY = 5
GPU_IN_USE = "nvidia.com/use-gputype"              # device.go:21-23
GPU_NO_USE = "nvidia.com/nouse-gputype"
NUMA_BIND = "nvidia.com/numa-bind"
TASK_PRIORITY_ENV = "CUDA_TASK_PRIORITY"           # api.TaskPriority
HANDSHAKE_TIME_FMT = "%Y.%m.%d %H:%M:%S"           # Go layout "2006.01.02 15:04:05"


# ---------------------------------------------------------------- native binding (include/vgpu_sched.h)
class _Usage(C.Structure):
    _fields_ = [("id", C.c_char * core.MAX_STR), ("type", C.c_char * core.MAX_STR), ("index", C.c_uint32), ("used", C.c_int32),
                ("count", C.c_int32), ("usedmem", C.c_int32), ("totalmem", C.c_int32), ("totalcore", C.c_int32),
                ("usedcores", C.c_int32), ("numa", C.c_int32), ("health", C.c_int32)]


class _Request(C.Structure):
    _fields_ = [("nums", C.c_int32), ("type", C.c_char * core.MAX_STR), ("memreq", C.c_int32), ("mem_percentage_req", C.c_int32),
                ("coresreq", C.c_int32)]


class _Annos(C.Structure):
    _fields_ = [("use_gputype", C.c_char_p), ("nouse_gputype", C.c_char_p), ("numa_bind", C.c_char_p)]


class _Assignment(C.Structure):
    _fields_ = [("idx", C.c_int32), ("container", C.c_int32), ("dev", core._CD)]


@dataclass
class DeviceUsage:                 # util.DeviceUsage (types.go:110-122)
    Id: str
    Index: int = 0
    Used: int = 0
    Count: int = 0
    Usedmem: int = 0
    Totalmem: int = 0
    Totalcore: int = 0
    Usedcores: int = 0
    Numa: int = 0
    Type: str = ""
    Health: bool = True


@dataclass
class DeviceInfo:                  # scheduler.DeviceInfo (nodes.go:28-37)
    ID: str
    Index: int = 0
    Count: int = 0
    Devmem: int = 0
    Devcore: int = 0
    Type: str = ""
    Numa: int = 0
    Health: bool = True


@dataclass
class NodeInfo:
    ID: str = ""
    Devices: list = field(default_factory=list)


@dataclass
class NodeUsage:
    Devices: list = field(default_factory=list)


@dataclass
class ContainerDeviceRequest:      # util.ContainerDeviceRequest (types.go:93-99)
    Nums: int = 0
    Type: str = ""
    Memreq: int = 0
    MemPercentagereq: int = 0
    Coresreq: int = 0


@dataclass
class PodInfo:                     # scheduler.podInfo (pods.go:28-35)
    Namespace: str
    Name: str
    Uid: str
    NodeID: str
    Devices: dict                  # util.PodDevices: {"NVIDIA": [[ContainerDevice, ...] per container]}


@dataclass
class NodeScore:
    nodeID: str
    devices: dict
    score: float


def _lib():
    L = core.lib()
    if not getattr(L, "_sched_ready", False):
        for n in ("vgpu_sched_check_type", "vgpu_sched_fit_in_certain_device", "vgpu_sched_fit_in_devices", "vgpu_sched_score_node"):
            getattr(L, n).restype = C.c_int
        L.vgpu_sched_sort_devices.restype = None
        L.vgpu_sched_charge.restype = None
        L._sched_ready = True
    return L


def _usage_arr(devs):
    arr = (_Usage * max(len(devs), 1))()
    for i, d in enumerate(devs):
        arr[i] = _Usage(d.Id.encode(), d.Type.encode(), d.Index, d.Used, d.Count, d.Usedmem, d.Totalmem, d.Totalcore, d.Usedcores,
                        d.Numa, int(d.Health))
    return arr


def _usage_list(arr, n):
    return [DeviceUsage(arr[i].id.decode(), arr[i].index, arr[i].used, arr[i].count, arr[i].usedmem, arr[i].totalmem, arr[i].totalcore,
                        arr[i].usedcores, arr[i].numa, arr[i].type.decode(), bool(arr[i].health)) for i in range(n)]


def _annos_struct(annos):
    annos = annos or {}
    enc = lambda k: annos[k].encode() if k in annos else None
    return _Annos(enc(GPU_IN_USE), enc(GPU_NO_USE), enc(NUMA_BIND))


def _req_struct(r):
    return _Request(r.Nums, r.Type.encode(), r.Memreq, r.MemPercentagereq, r.Coresreq)


class SchedulerPanic(RuntimeError):
    """The reference's Go code panics on this input (calcScore, score.go:211: index out of range)."""


def score_node(usage, reqs, annos, mode=0):
    """calcScore's per-node body. usage: NodeUsage (mutated like the reference: sorted, and charged on fit); reqs: one
    ContainerDeviceRequest per container (Nums == 0 = none). Returns (fit, score, {"NVIDIA": [[ContainerDevice]...]})."""
    L = _lib()
    n = len(usage.Devices)
    arr = _usage_arr(usage.Devices)
    rq = (_Request * max(len(reqs), 1))()
    for i, r in enumerate(reqs):
        rq[i] = _req_struct(r)
    cap = max(sum(r.Nums for r in reqs), 1)
    out = (_Assignment * cap)()
    n_out, sc = C.c_int(0), C.c_float(0)
    a = _annos_struct(annos)
    rc = L.vgpu_sched_score_node(arr, n, rq, len(reqs), C.byref(a), mode, out, cap, C.byref(n_out), C.byref(sc))
    usage.Devices[:] = _usage_list(arr, n)
    if rc == -5:
        raise SchedulerPanic("runtime error: index out of range (score.go:211: a device-less container after a fitted one)")
    if rc < 0:
        raise RuntimeError(f"vgpu_sched_score_node rc={rc}")
    per_ctr = {}
    for i in range(n_out.value):
        per_ctr.setdefault(out[i].container, []).append(
            core.ContainerDevice(out[i].dev.uuid.decode(), out[i].dev.type.decode(), out[i].dev.usedmem, out[i].dev.usedcores))
    devices = {}
    if per_ctr:
        if mode == 0:   # one slot per FITTED container, in order (fitInDevices appends, container index is not kept)
            devices[NVIDIA_GPU_DEVICE] = [per_ctr[c] for c in sorted(per_ctr)]
        else:           # one slot per container, empty for the device-less ones
            devices[NVIDIA_GPU_DEVICE] = [per_ctr.get(c, []) for c in range(len(reqs))]
    return rc == 1, float(sc.value), devices


# ---------------------------------------------------------------- resource.Quantity, the two accessors the reference uses
_QTY = re.compile(r"^([+-]?)(\d+)(?:\.(\d*))?(?:([eE])([+-]?\d+)|(Ki|Mi|Gi|Ti|Pi|Ei|n|u|m|k|M|G|T|P|E))?$")
_BIN = {"Ki": 2 ** 10, "Mi": 2 ** 20, "Gi": 2 ** 30, "Ti": 2 ** 40, "Pi": 2 ** 50, "Ei": 2 ** 60}
_DEC = {"n": -9, "u": -6, "m": -3, "": 0, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}


def _quantity_fraction(q):
    """(numerator, denominator) of a Kubernetes quantity given as str or number."""
    if isinstance(q, bool):
        raise ValueError(q)
    if isinstance(q, int):
        return q, 1
    m = _QTY.match(str(q).strip())
    if not m:
        raise ValueError(f"bad quantity {q!r}")
    sign, whole, frac, e, exp, suf = m.groups()
    frac = frac or ""
    num, den = int(whole + frac), 10 ** len(frac)
    if e:
        p = int(exp)
        num, den = (num * 10 ** p, den) if p >= 0 else (num, den * 10 ** -p)
    elif suf in _BIN:
        num *= _BIN[suf]
    else:
        p = _DEC[suf or ""]
        num, den = (num * 10 ** p, den) if p >= 0 else (num, den * 10 ** -p)
    if sign == "-":
        num = -num
    g = math.gcd(num, den) or 1
    return num // g, den // g


def quantity_as_int64(q):
    """Quantity.AsInt64(): (value, ok) — ok only for an integral value that fits int64."""
    try:
        num, den = _quantity_fraction(q)
    except ValueError:
        return 0, False
    if den != 1 or not -2 ** 63 <= num < 2 ** 63:
        return 0, False
    return num, True


def quantity_value(q):
    """Quantity.Value(): rounded up to the nearest integer."""
    num, den = _quantity_fraction(q)
    return -((-num) // den)


def _int32(x):
    return ((int(x) + 2 ** 31) % 2 ** 32) - 2 ** 31


@dataclass
class Config:                      # pkg/scheduler/config/config.go + device.go:43-49 flag defaults
    HttpBind: str = "127.0.0.1:8080"
    SchedulerName: str = ""
    DefaultMem: int = 0
    DefaultCores: int = 0
    MetricsBindAddress: str = ":9395"
    ResourceName: str = RESOURCE_NAME
    ResourceMem: str = RESOURCE_MEM
    ResourceMemPercentage: str = RESOURCE_MEM_PERCENTAGE
    ResourceCores: str = RESOURCE_CORES
    ResourcePriority: str = "vgputaskpriority"
    MultiContainer: bool = False   # False = the reference's container bookkeeping (vgpu_sched.h mode 0), True = mode 1


def _limit_or_request(ctr, name):
    res = ctr.get("resources") or {}
    for k in ("limits", "requests"):
        v = (res.get(k) or {}).get(name)
        if v is not None:
            return v
    return None


def generate_resource_requests(ctr, cfg):
    """NvidiaGPUDevices.GenerateResourceRequests (device.go:120-177)."""
    v = _limit_or_request(ctr, cfg.ResourceName)
    if v is None:
        return ContainerDeviceRequest()
    n, ok = quantity_as_int64(v)
    if not ok:
        return ContainerDeviceRequest()
    memnum = 0
    mem = _limit_or_request(ctr, cfg.ResourceMem)
    if mem is not None:
        val, ok = quantity_as_int64(mem)
        if ok:
            memnum = val
    mempnum = 101
    mem = _limit_or_request(ctr, cfg.ResourceMemPercentage)
    if mem is not None:
        val, ok = quantity_as_int64(mem)
        if ok:
            mempnum = _int32(val)
    if mempnum == 101 and memnum == 0:
        if cfg.DefaultMem != 0:
            memnum = cfg.DefaultMem
        else:
            mempnum = 100
    corenum = cfg.DefaultCores
    c = _limit_or_request(ctr, cfg.ResourceCores)
    if c is not None:
        val, ok = quantity_as_int64(c)
        if ok:
            corenum = _int32(val)
    return ContainerDeviceRequest(_int32(n), NVIDIA_GPU_DEVICE, _int32(memnum), _int32(mempnum), _int32(corenum))


def resource_reqs(pod, cfg):
    """k8sutil.Resourcereqs (pod.go:26-41): one request map per container, holding only requests with Nums > 0."""
    counts = []
    for ctr in (pod.get("spec") or {}).get("containers") or []:
        r = generate_resource_requests(ctr, cfg)
        counts.append({NVIDIA_GPU_DEVICE: r} if r.Nums > 0 else {})
    return counts


def is_pod_in_terminated_state(pod):
    return ((pod.get("status") or {}).get("phase")) in ("Failed", "Succeeded")


def _meta(obj):
    return obj.setdefault("metadata", {})


def _annotations(obj):
    return (obj.get("metadata") or {}).get("annotations") or {}


# ---------------------------------------------------------------- cluster access
class KubeClient:
    def get_pod(self, namespace, name):
        raise NotImplementedError

    def patch_pod_annotations(self, namespace, name, annotations):     # strategic merge of metadata.annotations
        raise NotImplementedError

    def bind_pod(self, namespace, name, uid, node):
        raise NotImplementedError

    def list_nodes(self):
        raise NotImplementedError

    def get_node(self, name):
        raise NotImplementedError

    def patch_node_annotations(self, name, annotations):
        raise NotImplementedError

    def update_node(self, node):                                        # full update (nodelock uses Update, not Patch)
        raise NotImplementedError


class InMemoryKube(KubeClient):
    def __init__(self, nodes=(), pods=()):
        self.nodes = {n["metadata"]["name"]: n for n in nodes}
        self.pods = {(p["metadata"].get("namespace", "default"), p["metadata"]["name"]): p for p in pods}
        self.bindings = []
        self.fail_patch = False

    def get_pod(self, namespace, name):
        return self.pods[(namespace, name)]

    def patch_pod_annotations(self, namespace, name, annotations):
        if self.fail_patch:
            raise RuntimeError("patch refused")
        p = self.pods[(namespace, name)]
        _meta(p).setdefault("annotations", {}).update(annotations)

    def bind_pod(self, namespace, name, uid, node):
        if node not in self.nodes:
            raise RuntimeError(f"nodes \"{node}\" not found")
        self.bindings.append((namespace, name, uid, node))
        self.pods[(namespace, name)].setdefault("spec", {})["nodeName"] = node

    def list_nodes(self):
        return list(self.nodes.values())

    def get_node(self, name):
        if name not in self.nodes:
            raise KeyError(f"nodes \"{name}\" not found")
        return self.nodes[name]

    def patch_node_annotations(self, name, annotations):
        _meta(self.nodes[name]).setdefault("annotations", {}).update(annotations)

    def update_node(self, node):
        self.nodes[node["metadata"]["name"]] = node


# ---------------------------------------------------------------- nodelock (pkg/util/nodelock/nodelock.go)
def _rfc3339(t):
    return time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(t))


def _parse_rfc3339(s):
    import calendar
    m = re.match(r"^(\d{4})-(\d\d)-(\d\d)T(\d\d):(\d\d):(\d\d)(?:\.\d+)?(Z|[+-]\d\d:\d\d)$", s)
    if not m:
        raise ValueError(f"parsing time {s!r}")
    t = calendar.timegm(tuple(int(x) for x in m.groups()[:6]) + (0, 0, 0))
    if m.group(7) != "Z":
        sign = 1 if m.group(7)[0] == "+" else -1
        t -= sign * (int(m.group(7)[1:3]) * 3600 + int(m.group(7)[4:6]) * 60)
    return t


def set_node_lock(kube, node_name, now=None):
    node = kube.get_node(node_name)
    if NODE_LOCK_TIME in _annotations(node):
        raise RuntimeError(f"node {node_name} is locked")
    new = json.loads(json.dumps(node))
    _meta(new).setdefault("annotations", {})[NODE_LOCK_TIME] = _rfc3339(time.time() if now is None else now)
    err = None
    for _ in range(1 + MAX_LOCK_RETRY):
        try:
            kube.update_node(new)
            return
        except Exception as e:   # conflict: re-read and retry
            err = e
            time.sleep(0.1)
            new = json.loads(json.dumps(kube.get_node(node_name)))
            _meta(new).setdefault("annotations", {})[NODE_LOCK_TIME] = _rfc3339(time.time() if now is None else now)
    raise RuntimeError(f"setNodeLock exceeds retry count {MAX_LOCK_RETRY}") from err


def release_node_lock(kube, node_name):
    node = kube.get_node(node_name)
    if NODE_LOCK_TIME not in _annotations(node):
        return
    new = json.loads(json.dumps(node))
    del new["metadata"]["annotations"][NODE_LOCK_TIME]
    err = None
    for _ in range(1 + MAX_LOCK_RETRY):
        try:
            kube.update_node(new)
            return
        except Exception as e:
            err = e
            time.sleep(0.1)
            new = json.loads(json.dumps(kube.get_node(node_name)))
            new["metadata"].get("annotations", {}).pop(NODE_LOCK_TIME, None)
    raise RuntimeError(f"releaseNodeLock exceeds retry count {MAX_LOCK_RETRY}") from err


def lock_node(kube, node_name, now=None):
    """LockNode (nodelock.go:86-109): take the lock, stealing one older than five minutes."""
    node = kube.get_node(node_name)
    stamp = _annotations(node).get(NODE_LOCK_TIME)
    if stamp is None:
        return set_node_lock(kube, node_name, now)
    lock_time = _parse_rfc3339(stamp)
    if (time.time() if now is None else now) - lock_time > 300:
        release_node_lock(kube, node_name)
        return set_node_lock(kube, node_name, now)
    raise RuntimeError(f"node {node_name} has been locked within 5 minutes")


# ---------------------------------------------------------------- the scheduler
class Scheduler:
    def __init__(self, kube=None, cfg=None):
        self.kube = kube
        self.cfg = cfg or Config()
        self.nodes = {}             # nodeManager.nodes
        self.pods = {}              # podManager.pods, by UID
        self.cachedstatus = {}
        self.overviewstatus = {}
        self._node_info_copy = {}   # RegisterFromNodeAnnotatons' nodeInfoCopy (keyed by handshake annotation)
        self._mu = threading.RLock()

    # -- nodeManager (nodes.go:52-116)
    def add_node(self, node_id, info):
        if info is None or not info.Devices:
            return
        with self._mu:
            if node_id in self.nodes:
                self.nodes[node_id].Devices = list(self.nodes[node_id].Devices) + list(info.Devices)
            else:
                self.nodes[node_id] = info

    def rm_node_device(self, node_id, info):
        with self._mu:
            cur = self.nodes.get(node_id)
            if cur is None or not cur.Devices:
                return
            gone = {d.ID for d in info.Devices}
            cur.Devices = [d for d in cur.Devices if d.ID not in gone and len(d.ID) > 0]

    def get_node(self, node_id):
        with self._mu:
            if node_id in self.nodes:
                return self.nodes[node_id]
        raise KeyError(f"node {node_id} not found")

    def list_nodes(self):
        return self.nodes

    # -- podManager (pods.go:46-72)
    def add_pod(self, pod, node_id, devices):
        with self._mu:
            uid = _meta(pod).get("uid", "")
            if uid not in self.pods:
                self.pods[uid] = PodInfo(_meta(pod).get("namespace", ""), _meta(pod).get("name", ""), uid, node_id, devices)

    def del_pod(self, pod):
        with self._mu:
            self.pods.pop(_meta(pod).get("uid", ""), None)

    def get_scheduled_pods(self):
        return self.pods

    # -- informer callbacks (scheduler.go:68-104)
    def on_add_pod(self, pod):
        annos = _annotations(pod)
        node_id = annos.get(ASSIGNED_NODE)
        if node_id is None:
            return
        if is_pod_in_terminated_state(pod):
            self.del_pod(pod)
            return
        self.add_pod(pod, node_id, decode_pod_devices({NVIDIA_GPU_DEVICE: ALLOCATED}, annos))

    def on_update_pod(self, _old, new):
        self.on_add_pod(new)

    def on_del_pod(self, pod):
        if ASSIGNED_NODE not in _annotations(pod):
            return
        self.del_pod(pod)

    # -- node registration from annotations (scheduler.go:123-241), one pass of its 15 s loop
    def register_from_node_annotations_once(self, now=None):
        now = time.time() if now is None else now
        node_names = []
        for val in self.kube.list_nodes():
            name = val["metadata"]["name"]
            node_names.append(name)
            annos = _annotations(val)
            if REGISTER not in annos:
                continue
            try:
                nodedevices = core.decode_node_devices(annos[REGISTER])
            except core.CodecError:
                continue
            if not nodedevices:
                continue
            handshake = annos.get(HANDSHAKE, "")
            if "Requesting" in handshake:
                former = _parse_handshake_time(handshake.split("_")[1])
                if now > former + 60 and name in self.nodes and self._node_info_copy.get(HANDSHAKE) is not None:
                    self.rm_node_device(name, self._node_info_copy[HANDSHAKE])
                    self.kube.patch_node_annotations(name, {HANDSHAKE: "Deleted_" + _handshake_time(now)})
                continue
            elif "Deleted" in handshake:
                continue
            else:
                self.kube.patch_node_annotations(name, {HANDSHAKE: "Requesting_" + _handshake_time(now)})
            info = NodeInfo(ID=name, Devices=[])
            for index, di in enumerate(nodedevices):
                found = False
                for cur in (self.nodes[name].Devices if name in self.nodes else []):
                    if cur.ID == di.Id:
                        found = True
                        cur.Devmem, cur.Devcore = di.Devmem, di.Devcore
                        break
                if not found:
                    info.Devices.append(DeviceInfo(di.Id, index, di.Count, di.Devmem, di.Devcore, di.Type, di.Numa, di.Health))
            self.add_node(name, info)
            self._node_info_copy[HANDSHAKE] = info
        self.get_nodes_usage(node_names, None)

    def inspect_all_nodes_usage(self):
        return self.overviewstatus

    # -- getNodesUsage (scheduler.go:250-313)
    def get_nodes_usage(self, node_names, task):
        L = _lib()
        overall, cache, failed = {}, {}, {}
        with self._mu:
            for node in self.nodes.values():
                overall[node.ID] = NodeUsage([DeviceUsage(d.ID, d.Index, 0, d.Count, 0, d.Devmem, d.Devcore, 0, d.Numa, d.Type, d.Health)
                                              for d in node.Devices])
            for p in self.pods.values():
                node = overall.get(p.NodeID)
                if node is None:
                    continue
                flat = [u for single in p.Devices.values() for ctr in single for u in ctr]
                if not flat:
                    continue
                arr = _usage_arr(node.Devices)
                L.vgpu_sched_charge(arr, len(node.Devices), core._cd_arr(flat), len(flat))
                node.Devices[:] = _usage_list(arr, len(node.Devices))
            self.overviewstatus = overall
            for node_id in node_names:
                if node_id not in self.nodes:
                    failed[node_id] = "node unregisterd"
                    continue
                cache[node_id] = overall[node_id]
            self.cachedstatus = cache
        return cache, failed

    # -- Filter (scheduler.go:361-407)
    def filter(self, args):
        pod = args.get("Pod") or args.get("pod") or {}
        node_names = args.get("NodeNames")
        if node_names is None:
            node_names = args.get("nodenames")
        nums = resource_reqs(pod, self.cfg)
        total = sum(k.Nums for n in nums for k in n.values())
        if total == 0:
            return filter_result(node_names=node_names)
        annos = _annotations(pod)
        self.del_pod(pod)
        usage, failed = self.get_nodes_usage(node_names or [], pod)
        scores = self.calc_score(usage, nums, annos, pod)
        if not scores:
            return filter_result(failed_nodes=failed)
        scores.sort(key=lambda s: s.score)                       # sort.Sort(nodeScores); the last (highest) wins
        m = scores[-1]
        annotations = {ASSIGNED_NODE: m.nodeID, ASSIGNED_TIME: str(int(time.time()))}
        annotations.update(encode_pod_devices({NVIDIA_GPU_DEVICE: TO_ALLOCATE}, m.devices))
        annotations.update(encode_pod_devices({NVIDIA_GPU_DEVICE: ALLOCATED}, m.devices))
        self.add_pod(pod, m.nodeID, m.devices)
        try:
            self.kube.patch_pod_annotations(_meta(pod).get("namespace", ""), _meta(pod).get("name", ""), annotations)
        except Exception:
            self.del_pod(pod)
            raise
        return filter_result(node_names=[m.nodeID])

    def calc_score(self, nodes, nums, annos, task):
        """calcScore (score.go:197-226). Nodes are visited in sorted-name order (Go's map order is random)."""
        res = []
        reqs = [n.get(NVIDIA_GPU_DEVICE, ContainerDeviceRequest()) for n in nums]
        for node_id in sorted(nodes):
            fit, sc, devices = score_node(nodes[node_id], reqs, annos, 1 if self.cfg.MultiContainer else 0)
            if fit:
                res.append(NodeScore(node_id, devices, sc))
        return res

    # -- Bind (scheduler.go:315-359)
    def bind(self, args):
        ns, name, uid, node = args.get("PodNamespace", ""), args.get("PodName", ""), args.get("PodUID", ""), args.get("Node", "")
        err = None
        try:
            lock_node(self.kube, node)
        except Exception:        # logged, not fatal in the reference
            pass
        try:
            self.kube.patch_pod_annotations(ns, name, {BIND_PHASE: BIND_ALLOCATING, BIND_TIME: str(int(time.time()))})
        except Exception:
            pass
        try:
            self.kube.bind_pod(ns, name, uid, node)
        except Exception as e:
            err = e
        return {"Error": "" if err is None else str(err)}


def filter_result(node_names=None, failed_nodes=None, error=""):
    """extenderv1.ExtenderFilterResult as Go marshals it (no json tags: field names, nil -> null)."""
    return {"Nodes": None, "NodeNames": node_names, "FailedNodes": failed_nodes, "FailedAndUnresolvableNodes": None, "Error": error}


def _handshake_time(t):
    return time.strftime(HANDSHAKE_TIME_FMT, time.localtime(t))


def _parse_handshake_time(s):
    try:
        return time.mktime(time.strptime(s, HANDSHAKE_TIME_FMT))
    except ValueError:
        return 0.0               # the reference ignores the parse error: zero time, i.e. long expired


def encode_pod_devices(checklist, pd):
    """util.EncodePodDevices (util.go:152-160)."""
    return {checklist[t]: core.encode_pod_single_device(single) for t, single in pd.items()}


def decode_pod_devices(checklist, annos):
    """util.DecodePodDevices (util.go:193-214): every ';'-separated piece is one container (so an encoded pod decodes with
    one trailing empty container — SURVEY.md Appendix E)."""
    if not annos:
        return {}
    pd = {}
    for dev_id, key in checklist.items():
        if key not in annos:
            continue
        try:
            pd[dev_id] = core.decode_pod_single_device(annos[key])
        except core.CodecError:
            return {}
    return pd


# ---------------------------------------------------------------- admission webhook (webhook.go:46-83)
def webhook_handle(review, cfg):
    """AdmissionReview (dict) -> AdmissionReview response. The reference marshals the mutated pod and lets
    controller-runtime diff it into a JSONPatch; here the patch operations are written directly."""
    req = review.get("request") or {}
    uid = req.get("uid", "")

    def resp(allowed, message=None, code=None, patch=None):
        r = {"uid": uid, "allowed": allowed}
        if message is not None or code is not None:
            r["status"] = {k: v for k, v in (("message", message), ("code", code)) if v is not None}
        if patch is not None:
            r["patchType"] = "JSONPatch"
            r["patch"] = base64.b64encode(json.dumps(patch).encode()).decode()
        return {"apiVersion": review.get("apiVersion", "admission.k8s.io/v1"), "kind": "AdmissionReview", "response": r}

    pod = req.get("object")
    if not isinstance(pod, dict):
        return resp(False, "there is no content to decode", 400)
    containers = (pod.get("spec") or {}).get("containers") or []
    if not containers:
        return resp(False, "pod has no containers", 403)
    patch, has_resource = [], False
    for idx, ctr in enumerate(containers):
        if ((ctr.get("securityContext") or {}).get("privileged")) is True:
            continue
        if has_resource:
            # webhook.go:68: `hasResource = hasResource || val.MutateAdmission(c)` — Go's || short-circuits, so once one
            # container asked for the resource no later container is mutated (no CUDA_TASK_PRIORITY for it). Kept.
            continue
        limits = (ctr.get("resources") or {}).get("limits") or {}
        if cfg.ResourcePriority in limits:                       # MutateAdmission (device.go:59-71)
            env = {"name": TASK_PRIORITY_ENV, "value": str(quantity_value(limits[cfg.ResourcePriority]))}
            if ctr.get("env"):
                patch.append({"op": "add", "path": f"/spec/containers/{idx}/env/-", "value": env})
            else:
                patch.append({"op": "add", "path": f"/spec/containers/{idx}/env", "value": [env]})
                ctr["env"] = [env]
        has_resource = has_resource or cfg.ResourceName in limits
    if not has_resource:
        return resp(True, "no resource found", 200)
    if cfg.SchedulerName:
        op = "replace" if "schedulerName" in (pod.get("spec") or {}) else "add"
        patch.append({"op": op, "path": "/spec/schedulerName", "value": cfg.SchedulerName})
    return resp(True, patch=patch)


# ---------------------------------------------------------------- Prometheus exposition (cmd/scheduler/metrics.go:55-200)
def _labels(**kv):
    return "{" + ",".join(f'{k}="{v}"' for k, v in kv.items()) + "}"


def collect_metrics(s, zone="vGPU"):
    MiB = 1024.0 * 1024.0
    out = []
    nu = s.inspect_all_nodes_usage()
    for node_id, val in nu.items():
        for d in val.Devices:
            base = dict(deviceidx=d.Index, deviceuuid=d.Id, nodeid=node_id, zone=zone)
            out.append(("GPUDeviceMemoryLimit", base, d.Totalmem * MiB))
            out.append(("GPUDeviceCoreLimit", base, float(d.Totalcore)))
            out.append(("GPUDeviceMemoryAllocated", dict(base, devicecores=d.Usedcores), d.Usedmem * MiB))
            out.append(("GPUDeviceSharedNum", base, float(d.Used)))
            out.append(("GPUDeviceCoreAllocated", base, float(d.Usedcores)))
            out.append(("nodeGPUOverview", dict(base, devicecores=d.Usedcores, sharedcontainers=d.Used, devicememorylimit=d.Totalmem,
                                                devicetype=d.Type), d.Usedmem * MiB))
            out.append(("nodeGPUMemoryPercentage", base, d.Usedmem / d.Totalmem if d.Totalmem else float("nan")))
    for p in s.get_scheduled_pods().values():
        # metrics.go:158: "ctridx" ranges over val.Devices, a map keyed by VENDOR — so the containeridx label carries the
        # vendor name ("NVIDIA"), not a container index. Kept.
        for vendor, single in p.Devices.items():
            for ctrdevs in single:
                for cd in ctrdevs:
                    base = dict(containeridx=vendor, deviceuuid=cd.UUID, nodename=p.NodeID, podname=p.Name, podnamespace=p.Namespace, zone=zone)
                    out.append(("vGPUPodsDeviceAllocated", dict(base, deviceusedcore=cd.Usedcores), cd.Usedmem * MiB))
                    total = next((d.Totalmem for u in nu.values() for d in u.Devices if d.Id == cd.UUID), 0)
                    if total > 0:
                        out.append(("vGPUMemoryPercentage", base, cd.Usedmem / total))
                    out.append(("vGPUCorePercentage", base, float(cd.Usedcores)))
    lines, seen = [], set()
    for name, labels, value in sorted(out, key=lambda x: x[0]):
        if name not in seen:
            seen.add(name)
            lines.append(f"# TYPE {name} gauge")
        lines.append(f"{name}{_labels(**dict(sorted(labels.items())))} {value:g}")
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------- HTTP routes (routes/route.go:41-134, main.go:66-84)
def make_handler(s):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, *a):
            pass

        def _send(self, code, body, ctype="application/json"):
            data = body if isinstance(body, bytes) else body.encode()
            self.send_response(code)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(data)))
            self.end_headers()
            self.wfile.write(data)

        def do_GET(self):
            if self.path == "/metrics":
                return self._send(200, collect_metrics(s), "text/plain; version=0.0.4")
            self._send(404, "404 page not found\n", "text/plain")

        def do_POST(self):
            n = int(self.headers.get("Content-Length") or 0)
            raw = self.rfile.read(n) if n else b""
            if self.path == "/filter":
                try:
                    args = json.loads(raw)
                    try:
                        result = s.filter(args)
                    except SchedulerPanic:
                        # net/http recovers the panic and drops the connection without a response
                        self.close_connection = True
                        self.connection.close()
                        return
                    except Exception as e:
                        result = filter_result(error=str(e))
                except ValueError as e:
                    result = filter_result(error=str(e))
                return self._send(200, json.dumps(result))
            if self.path == "/bind":
                try:
                    result = s.bind(json.loads(raw))
                except ValueError as e:
                    result = {"Error": str(e)}
                return self._send(200, json.dumps(result))
            if self.path == "/webhook":
                try:
                    return self._send(200, json.dumps(webhook_handle(json.loads(raw), s.cfg)))
                except ValueError as e:
                    return self._send(400, json.dumps({"response": {"allowed": False, "status": {"message": str(e), "code": 400}}}))
            self._send(404, "404 page not found\n", "text/plain")

    return Handler


def serve(s, bind=None, cert_file="", key_file=""):
    """Start the extender's HTTP(S) server in a thread; returns the server (server_address has the bound port)."""
    host, _, port = (bind or s.cfg.HttpBind).rpartition(":")
    srv = ThreadingHTTPServer((host or "0.0.0.0", int(port)), make_handler(s))
    if cert_file and key_file:
        import ssl
        ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
        ctx.load_cert_chain(cert_file, key_file)
        srv.socket = ctx.wrap_socket(srv.socket, server_side=True)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    return srv
