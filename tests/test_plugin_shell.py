"""The device-plugin shell around Allocate: Xid health policy, registration loop, kubelet-restart
loop — scripted events, no GPU, no cluster."""
import json
import os
import sys
import threading
import time

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import k8s_device_plugin_b200  # noqa: F401,E402
from k8s_device_plugin_b200.plugin import api, rm, scheduler as S  # noqa: E402
from k8s_device_plugin_b200.plugin import server as P  # noqa: E402
from k8s_device_plugin_b200.plugin.kubelet_stub import KubeletStub  # noqa: E402


def test_xid_policy():
    assert rm.skipped_xids("") == {13, 31, 43, 45, 68}                       # health.go:66-72
    assert rm.skipped_xids("79,abc, 94") == {13, 31, 43, 45, 68, 79, 94}
    assert rm.skipped_xids("all") is None and rm.skipped_xids("XIDS") is None and rm.skipped_xids("1,xids") is None


class Script(rm.EventSource):
    def __init__(self, events, stop, register_errors=()):
        self.events, self.stop, self.register_errors, self.closed = list(events), stop, set(register_errors), False

    def register(self, uuid):
        return "Not Supported" if uuid in self.register_errors else None

    def wait(self, timeout_ms):
        if not self.events:
            self.stop.set()
            return rm.Event(error="timeout")
        return self.events.pop(0)

    def close(self):
        self.closed = True


def test_health_check_marks_only_the_right_devices():
    devs = [P.GpuDevice("GPU-a"), P.GpuDevice("GPU-b"), P.GpuDevice("GPU-c")]
    stop, hit = threading.Event(), []
    src = Script([rm.Event(error="timeout"),
                  rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=31, uuid="GPU-a"),          # application error: ignored
                  rm.Event(etype=rm.EVENT_SINGLE_BIT_ECC, xid=79, uuid="GPU-a"),        # not an Xid event: ignored
                  rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=94, uuid="GPU-a"),          # operator-skipped
                  rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=79, uuid="GPU-zz"),         # unknown device: ignored
                  rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=79, uuid="GPU-b")], stop, register_errors={"GPU-c"})
    rm.check_health(stop, devs, lambda d: hit.append(d.ID), src, disable="94")
    assert hit == ["GPU-c", "GPU-b"] and src.closed
    # an unidentifiable device or a failing wait condemns every device
    stop, hit = threading.Event(), []
    rm.check_health(stop, devs, lambda d: hit.append(d.ID), Script([rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=79, uuid=None)], stop), disable="")
    assert hit == ["GPU-a", "GPU-b", "GPU-c"]
    stop, hit = threading.Event(), []
    rm.check_health(stop, devs, lambda d: hit.append(d.ID), Script([rm.Event(error="GPU is lost")], stop), disable="")
    assert hit == ["GPU-a", "GPU-b", "GPU-c"]
    stop, hit = threading.Event(), []
    rm.check_health(stop, devs, lambda d: hit.append(d.ID), Script([rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=79, uuid="GPU-a")], stop), disable="all")
    assert hit == []


def test_unhealthy_device_reaches_list_and_watch(tmp_path):
    devs = [P.GpuDevice("GPU-a"), P.GpuDevice("GPU-b")]
    plugin = P.NvidiaDevicePlugin(devs, P.InMemoryPodSource(), socket_dir=str(tmp_path), device_split_count=2)
    stream = plugin.ListAndWatch(None, None)
    first = next(stream)
    assert [d.health for d in first.devices] == [api.HEALTHY] * 4
    stop = threading.Event()
    rm.check_health(stop, devs, plugin.mark_unhealthy, Script([rm.Event(etype=rm.EVENT_XID_CRITICAL, xid=79, uuid="GPU-b")], stop), disable="")
    second = next(stream)
    assert [(d.ID, d.health) for d in second.devices] == [("GPU-a-0", api.HEALTHY), ("GPU-a-1", api.HEALTHY),
                                                          ("GPU-b-0", api.UNHEALTHY), ("GPU-b-1", api.UNHEALTHY)]
    assert plugin.node_annotations()[P.REGISTER].split(":")[1].endswith(",false")   # and the scheduler is told (register.go:131-142)
    plugin._stop.set()


def test_node_config_overrides(tmp_path):
    p = tmp_path / "config.json"
    p.write_text(json.dumps({"nodeconfig": [{"name": "n1", "devicememoryscaling": 1.8, "devicesplitcount": 10},
                                            {"name": "n2", "devicecorescaling": 2.0, "devicememoryscaling": 0}]}))
    assert rm.read_node_config(str(p), "n1") == (10, 1.8, 1.0)
    assert rm.read_node_config(str(p), "n2") == (2, 1.0, 2.0)
    assert rm.read_node_config(str(p), "other", 4, 1.5, 1.0) == (4, 1.5, 1.0)
    assert rm.read_node_config(str(tmp_path / "missing.json"), "n1") == (2, 1.0, 1.0)


def test_watch_and_register_drives_the_scheduler_handshake(tmp_path):
    kube = S.InMemoryKube([{"metadata": {"name": "node-0", "annotations": {}}}])
    plugin = P.NvidiaDevicePlugin([P.GpuDevice("GPU-a")], P.InMemoryPodSource(), node_name="node-0", socket_dir=str(tmp_path), device_split_count=4,
                                  device_memory_scaling=2.0)
    stop, logs = threading.Event(), []
    t = threading.Thread(target=rm.watch_and_register, args=(plugin, kube, stop, 0.05, 0.02, logs.append))
    t.start()
    time.sleep(0.2)
    annos = kube.nodes["node-0"]["metadata"]["annotations"]
    assert annos[P.HANDSHAKE].startswith("Reported ") and annos[P.REGISTER] == "GPU-a,4,366718,100,NVIDIA-NVIDIA B200,0,true:"
    # the scheduler answers "Requesting_<t>", the next pass overwrites it with "Reported" again: the node stays registered
    sch = S.Scheduler(kube)
    sch.register_from_node_annotations_once()
    assert kube.nodes["node-0"]["metadata"]["annotations"][P.HANDSHAKE].startswith("Requesting_") or \
        kube.nodes["node-0"]["metadata"]["annotations"][P.HANDSHAKE].startswith("Reported ")
    time.sleep(0.15)
    assert kube.nodes["node-0"]["metadata"]["annotations"][P.HANDSHAKE].startswith("Reported ")
    assert [d.Devmem for d in sch.nodes["node-0"].Devices] == [366718]
    # node gone: the loop keeps retrying on the short interval
    del kube.nodes["node-0"]
    time.sleep(0.1)
    stop.set(); t.join(2)
    assert any("Failed to register annotation" in l for l in logs)


def test_plugin_manager_restarts_when_the_kubelet_socket_is_recreated(tmp_path):
    made = []

    def make():
        p = P.NvidiaDevicePlugin([P.GpuDevice("GPU-a")], P.InMemoryPodSource(), socket_dir=str(tmp_path))
        made.append(p)
        return [p]

    stub = KubeletStub(str(tmp_path)); stub.start()
    mgr = rm.PluginManager(make, os.path.join(str(tmp_path), "kubelet.sock"), retry_s=0.3, poll_s=0.05)
    t = threading.Thread(target=mgr.run)
    t.start()
    try:
        deadline = time.time() + 5
        while len(stub.registrations) < 1 and time.time() < deadline:
            time.sleep(0.02)
        assert len(stub.registrations) == 1 and stub.registrations[0].resource_name == "nvidia.com/gpu"
        stub.stop(); stub = KubeletStub(str(tmp_path)); stub.start()              # kubelet restart: new socket inode
        deadline = time.time() + 5
        while len(stub.registrations) < 1 and time.time() < deadline:
            time.sleep(0.02)
        assert len(stub.registrations) == 1 and len(made) == 2 and made[0].server is None
        mgr.notify("restart")                                                       # SIGHUP
        deadline = time.time() + 5
        while len(stub.registrations) < 2 and time.time() < deadline:
            time.sleep(0.02)
        assert len(stub.registrations) == 2 and len(made) == 3
    finally:
        mgr.notify("exit"); t.join(5); stub.stop()
    assert not t.is_alive() and made[-1].server is None


def test_plugin_manager_retries_while_the_kubelet_is_away(tmp_path):
    made = []

    def make():
        p = P.NvidiaDevicePlugin([P.GpuDevice("GPU-a")], P.InMemoryPodSource(), socket_dir=str(tmp_path))
        made.append(p)
        return [p]

    mgr = rm.PluginManager(make, os.path.join(str(tmp_path), "kubelet.sock"), retry_s=0.2, poll_s=0.05)
    t = threading.Thread(target=mgr.run)
    t.start()
    time.sleep(0.7)                      # no kubelet: Start() fails, retried every retry_s
    assert len(made) >= 2
    stub = KubeletStub(str(tmp_path)); stub.start()
    try:
        deadline = time.time() + 5
        while len(stub.registrations) < 1 and time.time() < deadline:
            time.sleep(0.02)
        assert len(stub.registrations) >= 1
    finally:
        mgr.notify("exit"); t.join(5); stub.stop()


def test_additional_xids_vectors_of_the_reference_test():
    """rm/health_test.go:25-88 TestGetAdditionalXids, case for case."""
    cases = [("", []), (",", []), ("not-an-int", []), ("68", [68]), ("-68", []), ("68  ", [68]), ("68,", [68]), (",68", [68]),
             ("68,67", [68, 67]), ("68,not-an-int,67", [68, 67])]
    for text, want in cases:
        assert rm.additional_xids(text) == want, text


def test_numa_from_nvidia_smi_topo_like_the_reference():
    """plugin/register_test.go:21-69 Test_parseNvidiaNumaInfo (its three elided topologies all want 0), then real layouts:
    the header's NUMA Affinity column applies to the rows after double tabs are collapsed (register.go:45-93)."""
    elided = "GPU0    CPU Affinity    NUMA Affinity ...\n                            ..."
    assert rm.parse_nvidia_numa_info(0, elided) == 0
    two = "GPU0    GPU1    CPU Affinity    NUMA Affinity ...\n                            ..."
    assert rm.parse_nvidia_numa_info(0, two) == 0 and rm.parse_nvidia_numa_info(1, two) == 0
    single = "\tGPU0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\nGPU0\t X \t0-7\t\tN/A\t\tN/A\nLegend:\n  X = Self\n"
    assert rm.parse_nvidia_numa_info(0, single) == 0                                    # N/A: no NUMA topology established
    multi = ("\tGPU0\tGPU1\tGPU2\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\n"
             "GPU0\t X \tNV18\tNV18\t0-55\t\t0\t\tN/A\n"
             "GPU1\tNV18\t X \tNV18\t0-55\t\t0\t\tN/A\n"
             "GPU2\tNV18\tNV18\t X \t56-111\t\t1\t\tN/A\n"
             "\nLegend:\n\n  X    = Self\n  NV#  = Connection traversing a bonded set of # NVLinks\n")
    assert [rm.parse_nvidia_numa_info(i, multi) for i in range(3)] == [0, 0, 1]
    with pytest.raises(ValueError):
        rm.parse_nvidia_numa_info(2, multi.replace("\t1\t\tN/A", "\t0-1\t\tN/A"))       # strconv.Atoi error is returned, not swallowed
    # sysfs missing -> the nvidia-smi route; a failing or absent nvidia-smi -> 0
    assert rm.numa_node_of(2, "0000:FF:1F.7", run=lambda: multi) == 1
    assert rm.numa_node_of(2, "0000:FF:1F.7", run=lambda: (_ for _ in ()).throw(OSError("no nvidia-smi"))) == 0
