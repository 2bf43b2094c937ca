"""kubelet DevicePlugin API v1beta1 (k8s.io/kubelet v0.28.3, pluginapi in the reference's go.mod) as protobuf message
classes built at import time from a FileDescriptorProto — this image has grpcio + protobuf but no protoc. Only the
field NUMBERS and types matter on the wire; they follow k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

VERSION = "v1beta1"                                   # pluginapi.Version
DEVICE_PLUGIN_PATH = "/var/lib/kubelet/device-plugins/"
KUBELET_SOCKET = DEVICE_PLUGIN_PATH + "kubelet.sock"
HEALTHY, UNHEALTHY = "Healthy", "Unhealthy"

_T = descriptor_pb2.FieldDescriptorProto
_fd = descriptor_pb2.FileDescriptorProto(name="vgpu_b200/deviceplugin_v1beta1.proto", package="v1beta1", syntax="proto3")


def _msg(name, fields, maps=()):
    m = _fd.message_type.add(name=name)
    for fname, num, ftype, label, tname in fields:
        f = m.field.add(name=fname, number=num, type=ftype, label=label)
        if tname:
            f.type_name = tname
    for fname, num in maps:
        entry = m.nested_type.add(name="".join(p.capitalize() for p in fname.split("_")) + "Entry")
        entry.options.map_entry = True
        entry.field.add(name="key", number=1, type=_T.TYPE_STRING, label=_T.LABEL_OPTIONAL)
        entry.field.add(name="value", number=2, type=_T.TYPE_STRING, label=_T.LABEL_OPTIONAL)
        m.field.add(name=fname, number=num, type=_T.TYPE_MESSAGE, label=_T.LABEL_REPEATED, type_name=f".v1beta1.{name}.{entry.name}")
    return m


O, R = _T.LABEL_OPTIONAL, _T.LABEL_REPEATED
_msg("Empty", [])
_msg("DevicePluginOptions", [("pre_start_required", 1, _T.TYPE_BOOL, O, None), ("get_preferred_allocation_available", 2, _T.TYPE_BOOL, O, None)])
_msg("RegisterRequest", [("version", 1, _T.TYPE_STRING, O, None), ("endpoint", 2, _T.TYPE_STRING, O, None),
                         ("resource_name", 3, _T.TYPE_STRING, O, None), ("options", 4, _T.TYPE_MESSAGE, O, ".v1beta1.DevicePluginOptions")])
_msg("NUMANode", [("ID", 1, _T.TYPE_INT64, O, None)])
_msg("TopologyInfo", [("nodes", 1, _T.TYPE_MESSAGE, R, ".v1beta1.NUMANode")])
_msg("Device", [("ID", 1, _T.TYPE_STRING, O, None), ("health", 2, _T.TYPE_STRING, O, None), ("topology", 3, _T.TYPE_MESSAGE, O, ".v1beta1.TopologyInfo")])
_msg("ListAndWatchResponse", [("devices", 1, _T.TYPE_MESSAGE, R, ".v1beta1.Device")])
_msg("PreStartContainerRequest", [("devices_ids", 1, _T.TYPE_STRING, R, None)])
_msg("PreStartContainerResponse", [])
_msg("ContainerPreferredAllocationRequest", [("available_deviceIDs", 1, _T.TYPE_STRING, R, None), ("must_include_deviceIDs", 2, _T.TYPE_STRING, R, None),
                                             ("allocation_size", 3, _T.TYPE_INT32, O, None)])
_msg("PreferredAllocationRequest", [("container_requests", 1, _T.TYPE_MESSAGE, R, ".v1beta1.ContainerPreferredAllocationRequest")])
_msg("ContainerPreferredAllocationResponse", [("deviceIDs", 1, _T.TYPE_STRING, R, None)])
_msg("PreferredAllocationResponse", [("container_responses", 1, _T.TYPE_MESSAGE, R, ".v1beta1.ContainerPreferredAllocationResponse")])
_msg("ContainerAllocateRequest", [("devices_ids", 1, _T.TYPE_STRING, R, None)])
_msg("AllocateRequest", [("container_requests", 1, _T.TYPE_MESSAGE, R, ".v1beta1.ContainerAllocateRequest")])
_msg("Mount", [("container_path", 1, _T.TYPE_STRING, O, None), ("host_path", 2, _T.TYPE_STRING, O, None), ("read_only", 3, _T.TYPE_BOOL, O, None)])
_msg("DeviceSpec", [("container_path", 1, _T.TYPE_STRING, O, None), ("host_path", 2, _T.TYPE_STRING, O, None), ("permissions", 3, _T.TYPE_STRING, O, None)])
_msg("CDIDevice", [("name", 1, _T.TYPE_STRING, O, None)])
_msg("ContainerAllocateResponse", [("mounts", 2, _T.TYPE_MESSAGE, R, ".v1beta1.Mount"), ("devices", 3, _T.TYPE_MESSAGE, R, ".v1beta1.DeviceSpec"),
                                   ("cdi_devices", 5, _T.TYPE_MESSAGE, R, ".v1beta1.CDIDevice")], maps=[("envs", 1), ("annotations", 4)])
_msg("AllocateResponse", [("container_responses", 1, _T.TYPE_MESSAGE, R, ".v1beta1.ContainerAllocateResponse")])

_pool = descriptor_pool.DescriptorPool()
_pool.Add(_fd)


def _cls(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName("v1beta1." + name))


Empty = _cls("Empty")
DevicePluginOptions = _cls("DevicePluginOptions")
RegisterRequest = _cls("RegisterRequest")
Device = _cls("Device")
ListAndWatchResponse = _cls("ListAndWatchResponse")
PreStartContainerRequest = _cls("PreStartContainerRequest")
PreStartContainerResponse = _cls("PreStartContainerResponse")
PreferredAllocationRequest = _cls("PreferredAllocationRequest")
PreferredAllocationResponse = _cls("PreferredAllocationResponse")
AllocateRequest = _cls("AllocateRequest")
ContainerAllocateRequest = _cls("ContainerAllocateRequest")
AllocateResponse = _cls("AllocateResponse")
ContainerAllocateResponse = _cls("ContainerAllocateResponse")
Mount = _cls("Mount")

# full gRPC method names of the two services
M_REGISTER = "/v1beta1.Registration/Register"
M_OPTIONS = "/v1beta1.DevicePlugin/GetDevicePluginOptions"
M_LIST_AND_WATCH = "/v1beta1.DevicePlugin/ListAndWatch"
M_PREFERRED = "/v1beta1.DevicePlugin/GetPreferredAllocation"
M_ALLOCATE = "/v1beta1.DevicePlugin/Allocate"
M_PRESTART = "/v1beta1.DevicePlugin/PreStartContainer"
