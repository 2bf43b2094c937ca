"""A kubelet stand-in for BASELINE.json configs[0]: serves v1beta1.Registration on kubelet.sock, then dials the
registered plugin endpoint like the kubelet's device manager does (ListAndWatch stream, Allocate)."""
import os
import threading
from concurrent import futures

import grpc

from . import api


class KubeletStub:
    def __init__(self, socket_dir):
        self.socket_dir = socket_dir
        self.socket = os.path.join(socket_dir, "kubelet.sock")
        self.registered = threading.Event()
        self.request = None
        self.registrations = []           # every RegisterRequest seen, in order
        self.server = None

    def _register(self, request, context):
        self.request = request
        self.registrations.append(request)
        self.registered.set()
        return api.Empty()

    def start(self):
        os.makedirs(self.socket_dir, exist_ok=True)
        if os.path.exists(self.socket):
            os.remove(self.socket)
        s = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
        h = {"Register": grpc.unary_unary_rpc_method_handler(self._register, api.RegisterRequest.FromString, lambda m: m.SerializeToString())}
        s.add_generic_rpc_handlers((grpc.method_handlers_generic_handler("v1beta1.Registration", h),))
        s.add_insecure_port("unix://" + self.socket)
        s.start()
        self.server = s

    def stop(self):
        if self.server:
            self.server.stop(0)

    # ---- the kubelet's client side
    def _channel(self):
        return grpc.insecure_channel("unix://" + os.path.join(self.socket_dir, self.request.endpoint))

    def list_and_watch_once(self, n=1, timeout=5):
        out = []
        with self._channel() as ch:
            call = ch.unary_stream(api.M_LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                   response_deserializer=api.ListAndWatchResponse.FromString)
            stream = call(api.Empty(), timeout=timeout)
            for resp in stream:
                out.append(resp)
                if len(out) >= n:
                    stream.cancel()
                    break
        return out

    def allocate(self, device_ids_per_container, timeout=5):
        with self._channel() as ch:
            call = ch.unary_unary(api.M_ALLOCATE, request_serializer=lambda m: m.SerializeToString(), response_deserializer=api.AllocateResponse.FromString)
            req = api.AllocateRequest()
            for ids in device_ids_per_container:
                req.container_requests.add(devices_ids=list(ids))
            return call(req, timeout=timeout)

    def _unary(self, method, req, resp_cls, timeout=5):
        with self._channel() as ch:
            return ch.unary_unary(method, request_serializer=lambda m: m.SerializeToString(), response_deserializer=resp_cls.FromString)(req, timeout=timeout)

    def options(self):
        return self._unary(api.M_OPTIONS, api.Empty(), api.DevicePluginOptions)

    def preferred_allocation(self, available, must_include, size):
        req = api.PreferredAllocationRequest()
        req.container_requests.add(available_deviceIDs=list(available), must_include_deviceIDs=list(must_include), allocation_size=size)
        return self._unary(api.M_PREFERRED, req, api.PreferredAllocationResponse)

    def pre_start(self, ids):
        return self._unary(api.M_PRESTART, api.PreStartContainerRequest(devices_ids=list(ids)), api.PreStartContainerResponse)
