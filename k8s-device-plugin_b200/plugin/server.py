"""NvidiaDevicePlugin — the kubelet-facing server for the vGPU path, method for method after the reference's
pkg/device-plugin/nvidiadevice/nvinternal/plugin/server.go (Start:122, Serve:170, Register:220, ListAndWatch:253,
GetPreferredAllocation, Allocate:288-411, PreStartContainer) with the decisions delegated to the native core
(csrc/plugin_core.cc). Cluster access (pending pod lookup, annotation patch — util.GetPendingPod util.go:51-76,
util.PatchPodAnnotations) goes through a small PodSource interface so tests can run without an apiserver."""
import os
import queue
import threading
import uuid as uuidlib
from concurrent import futures
from dataclasses import dataclass, field

import grpc

from . import api, core

# annotation keys (pkg/util/types.go:25-30, pkg/device/nvidia/device.go:16-17,38-39)
ASSIGNED_NODE = "4pd.io/vgpu-node"
BIND_TIME = "4pd.io/bind-time"
BIND_PHASE = "4pd.io/bind-phase"
BIND_ALLOCATING, BIND_FAILED, BIND_SUCCESS = "allocating", "failed", "success"
HANDSHAKE = "4pd.io/node-handshake"
REGISTER = "4pd.io/node-nvidia-register"
TO_ALLOCATE = "hami.sh/vgpu-devices-to-allocate"
ALLOCATED = "hami.sh/vgpu-devices-allocated"
NVIDIA_GPU_DEVICE = "NVIDIA"
# resource names (charts/vgpu/values.yaml:9-13, pkg/device/nvidia/device.go:43-49)
RESOURCE_NAME = "nvidia.com/gpu"
RESOURCE_MEM = "nvidia.com/gpumem"
RESOURCE_MEM_PERCENTAGE = "nvidia.com/gpumem-percentage"
RESOURCE_CORES = "nvidia.com/gpucores"


@dataclass
class GpuDevice:                 # rm.Device: the physical GPU as the resource manager reports it
    ID: str
    Health: str = api.HEALTHY
    TotalMemory: int = 183359 << 20
    Model: str = "NVIDIA B200"
    Numa: int = 0


@dataclass
class Container:
    Name: str
    Env: dict = field(default_factory=dict)


@dataclass
class Pod:
    UID: str
    Name: str
    Annotations: dict
    Containers: list


class PodSource:
    """What Allocate needs from the cluster. InMemoryPodSource backs the kubelet-stub tests; a real deployment
    implements it with client-go (reference: util.GetPendingPod / PatchPodAnnotations / nodelock)."""

    def get_pending_pod(self, node):
        raise NotImplementedError

    def patch_pod_annotations(self, pod, annos):
        raise NotImplementedError

    def release_node_lock(self, node):
        pass


class InMemoryPodSource(PodSource):
    def __init__(self, pods=()):
        self.pods = list(pods)
        self.lock_released = 0

    def get_pending_pod(self, node):                          # util.GetPendingPod (util.go:51-76)
        for p in self.pods:
            a = p.Annotations
            if BIND_TIME not in a or a.get(BIND_PHASE) != BIND_ALLOCATING:
                continue
            if a.get(ASSIGNED_NODE) == node:
                return p
        raise LookupError(f"no binding pod found on node {node}")

    def patch_pod_annotations(self, pod, annos):
        pod.Annotations.update(annos)

    def release_node_lock(self, node):
        self.lock_released += 1


class NvidiaDevicePlugin:
    def __init__(self, devices, pods, node_name="node-0", resource_name=RESOURCE_NAME, socket_dir=api.DEVICE_PLUGIN_PATH,
                 host_hook_path="/usr/local", device_split_count=2, device_memory_scaling=1.0, device_cores_scaling=1.0,
                 disable_core_limit=False, kubelet_socket=None):
        self.devices = list(devices)
        self.pods = pods
        self.node_name = node_name
        self.resource_name = resource_name
        self.host_hook_path = host_hook_path
        self.device_split_count = device_split_count
        self.device_memory_scaling = device_memory_scaling
        self.device_cores_scaling = device_cores_scaling
        self.disable_core_limit = disable_core_limit
        # server.go:89: pluginapi.DevicePluginPath + "nvidia-<name>.sock"
        self.socket = os.path.join(socket_dir, "nvidia-" + resource_name.split("/")[-1] + ".sock")
        self.kubelet_socket = kubelet_socket or os.path.join(socket_dir, "kubelet.sock")
        self.server = None
        self._health = queue.Queue()
        self._stop = threading.Event()

    # ---- rm.Devices().GetPluginDevices() (rm/devices.go:144-167)
    def plugin_devices(self):
        out = []
        for d in self.devices:
            for i in range(self.device_split_count):
                out.append(api.Device(ID=core.device_id(d.ID, i), health=d.Health))
        return out

    # ---- register.go:96-162 getApiDevices + :164-183 RegistrInAnnotation
    def node_annotations(self, now="now"):
        devs = [core.NodeDevice(d.ID, self.device_split_count, core.registered_mem(d.TotalMemory, self.device_memory_scaling),
                                core.registered_cores(self.device_cores_scaling), f"NVIDIA-{d.Model}", d.Numa,
                                d.Health.lower() == "healthy") for d in self.devices]
        return {HANDSHAKE: "Reported " + now, REGISTER: core.encode_node_devices(devs)}

    # ---- gRPC service methods
    def GetDevicePluginOptions(self, request, context):
        return api.DevicePluginOptions(get_preferred_allocation_available=True)

    def ListAndWatch(self, request, context):                 # server.go:253-267
        yield api.ListAndWatchResponse(devices=self.plugin_devices())
        while not self._stop.is_set():
            try:
                d = self._health.get(timeout=0.2)
            except queue.Empty:
                if context is not None and not context.is_active():
                    return
                continue
            d.Health = api.UNHEALTHY
            yield api.ListAndWatchResponse(devices=self.plugin_devices())

    def mark_unhealthy(self, dev):
        self._health.put(dev)

    def GetPreferredAllocation(self, request, context):
        return api.PreferredAllocationResponse()

    def PreStartContainer(self, request, context):
        return api.PreStartContainerResponse()

    def Allocate(self, request, context):                     # server.go:288-411
        try:
            current = self.pods.get_pending_pod(self.node_name)
        except LookupError as e:
            self.pods.release_node_lock(self.node_name)
            return self._abort(context, str(e))
        resp = api.AllocateResponse()
        for req in request.container_requests:
            anno = current.Annotations.get(TO_ALLOCATE, "")
            try:
                ctr_idx, devreq = core.next_device_request(anno)
            except (LookupError, core.CodecError):
                self._failed(current)
                return self._abort(context, "device request not found")
            try:
                envs, mounts, cache_dir = core.allocate(
                    devreq, len(req.devices_ids), self.host_hook_path, current.UID, current.Containers[ctr_idx].Name,
                    cache_uuid=str(uuidlib.uuid4()), device_memory_scaling=self.device_memory_scaling,
                    disable_core_limit=self.disable_core_limit,
                    container_sets_disable_control="CUDA_DISABLE_CONTROL" in current.Containers[ctr_idx].Env,
                    license_present=os.path.exists(os.path.join(self.host_hook_path, "vgpu", "license")))
            except ValueError:
                self._failed(current)
                return self._abort(context, "device allocate number not matched")
            self.pods.patch_pod_annotations(current, {TO_ALLOCATE: core.erase_next_device_request(anno)})
            try:                                               # server.go:364-367
                # MkdirAll + Chmod(0777): the mode passed to makedirs is masked by the daemon's umask, and a
                # runAsNonRoot container must be able to create its region file in there
                for d in (cache_dir, "/tmp/vgpulock"):
                    os.makedirs(d, mode=0o777, exist_ok=True)
                    os.chmod(d, 0o777)
            except OSError:
                pass
            c = resp.container_responses.add()
            for k, v in envs.items():
                c.envs[k] = v
            for cp, hp, ro in mounts:
                c.mounts.add(container_path=cp, host_path=hp, read_only=ro)
        self._try_success(current)
        return resp

    # device.PodAllocationFailed / PodAllocationTrySuccess (pkg/device/devices.go:40-65)
    def _failed(self, pod):
        self.pods.patch_pod_annotations(pod, {BIND_PHASE: BIND_FAILED})
        self.pods.release_node_lock(self.node_name)

    def _try_success(self, pod):
        anno = pod.Annotations.get(TO_ALLOCATE, "")
        if NVIDIA_GPU_DEVICE in anno:                          # some container still waits for its devices
            return
        self.pods.patch_pod_annotations(pod, {BIND_PHASE: BIND_SUCCESS})
        self.pods.release_node_lock(self.node_name)

    @staticmethod
    def _abort(context, msg):
        if context is not None:
            context.abort(grpc.StatusCode.UNKNOWN, msg)
        raise RuntimeError(msg)

    # ---- lifecycle (server.go:122-243)
    def Serve(self):
        if os.path.exists(self.socket):
            os.remove(self.socket)
        os.makedirs(os.path.dirname(self.socket), exist_ok=True)
        s = grpc.server(futures.ThreadPoolExecutor(max_workers=4))
        ser = lambda m: m.SerializeToString()
        handlers = {
            "GetDevicePluginOptions": grpc.unary_unary_rpc_method_handler(self.GetDevicePluginOptions, api.Empty.FromString, ser),
            "ListAndWatch": grpc.unary_stream_rpc_method_handler(self.ListAndWatch, api.Empty.FromString, ser),
            "GetPreferredAllocation": grpc.unary_unary_rpc_method_handler(self.GetPreferredAllocation, api.PreferredAllocationRequest.FromString, ser),
            "Allocate": grpc.unary_unary_rpc_method_handler(self.Allocate, api.AllocateRequest.FromString, ser),
            "PreStartContainer": grpc.unary_unary_rpc_method_handler(self.PreStartContainer, api.PreStartContainerRequest.FromString, ser),
        }
        s.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("v1beta1.DevicePlugin", handlers),))
        s.add_insecure_port("unix://" + self.socket)
        s.start()
        self.server = s

    def Register(self):                                       # server.go:220-243
        with grpc.insecure_channel("unix://" + self.kubelet_socket) as ch:
            call = ch.unary_unary(api.M_REGISTER, request_serializer=lambda m: m.SerializeToString(), response_deserializer=api.Empty.FromString)
            call(api.RegisterRequest(version=api.VERSION, endpoint=os.path.basename(self.socket), resource_name=self.resource_name,
                                     options=api.DevicePluginOptions(get_preferred_allocation_available=True)), timeout=5)

    def Start(self):
        self.Serve()
        self.Register()

    def Stop(self):
        self._stop.set()
        if self.server:
            self.server.stop(0)
            self.server = None
        if os.path.exists(self.socket):
            os.remove(self.socket)
