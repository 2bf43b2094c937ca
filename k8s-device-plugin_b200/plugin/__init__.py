"""Host-side (kubelet-facing) half of the vGPU path: the device plugin's ListAndWatch/Allocate surface.

The decision logic is native (csrc/plugin_core.cc behind include/vgpu_plugin.h, mirroring the reference's Go:
plugin/server.go, rm/devices.go, pkg/util/util.go); this package is only the transport a kubelet talks to — the
v1beta1 DevicePlugin gRPC API — plus a kubelet stub for tests (BASELINE.json configs[0]). The image has no Go
toolchain; INTEGRATION.md shows the cgo binding a Go plugin would use instead of this shell."""
from .core import (ContainerDevice, NodeDevice, allocate, decode_container_devices, decode_node_devices,  # noqa: F401
                   decode_pod_single_device, device_id, encode_container_devices, encode_node_devices,
                   encode_pod_single_device, erase_next_device_request, next_device_request, registered_cores,
                   registered_mem)
