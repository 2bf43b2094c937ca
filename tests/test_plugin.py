"""BASELINE.json configs[0]: kubelet-stub Register / ListAndWatch / Allocate round trip on 2 fake vGPU-capable GPUs,
CPU only — plus the reference's own goldens for this boundary (pkg/util/util_test.go:33-64 codec,
plugin/server_test.go:174-184 container dir format) and the wire formats of SURVEY.md §8b."""
import time

import pytest

import k8s_device_plugin_b200.plugin as P
from k8s_device_plugin_b200.plugin import ContainerDevice as CD
from k8s_device_plugin_b200.plugin import server as S
from k8s_device_plugin_b200.plugin.kubelet_stub import KubeletStub


def test_codec_empty_cases_from_reference_util_test():
    # util_test.go:33-47 TestEmptyContainerDevicesCoding: encode of nothing is "", decode of "" is []
    assert P.encode_container_devices([]) == ""
    assert P.decode_container_devices("") == []


def test_codec_single_container_roundtrip_and_reference_quirk():
    # util_test.go:49-64 builds ContainerDevice{0,"UUID1","Type1",1000,30} twice
    d = CD("UUID1", "Type1", 1000, 30)
    assert P.encode_container_devices([d]) == "UUID1,Type1,1000,30:"
    assert P.decode_container_devices("UUID1,Type1,1000,30:") == [d]
    one = P.encode_pod_single_device([[d]])
    assert one == "UUID1,Type1,1000,30:;"
    assert P.decode_pod_single_device(one) == [[d], []]           # trailing empty container
    # two containers x one GPU: EncodePodSingleDevice (util.go:142-150) emits ONE ';' -> both devices land in
    # container 0 when decoded. Reproduced byte for byte; the reference's own test expects [[A],[B]] and cannot pass.
    two = P.encode_pod_single_device([[d], [d]])
    assert two == "UUID1,Type1,1000,30:UUID1,Type1,1000,30:;"
    assert P.decode_pod_single_device(two) == [[d, d], []]


def test_codec_errors_and_node_devices():
    with pytest.raises(ValueError):
        P.decode_container_devices("UUID1,Type1,1000:")           # < 4 fields: "information missing"
    nd = P.NodeDevice("GPU-aaaa", 10, 183359, 100, "NVIDIA-NVIDIA B200", 1, True)
    s = P.encode_node_devices([nd, nd])
    assert s == "GPU-aaaa,10,183359,100,NVIDIA-NVIDIA B200,1,true:" * 2
    assert P.decode_node_devices(s) == [nd, nd]
    with pytest.raises(ValueError):
        P.decode_node_devices("garbage")
    with pytest.raises(ValueError):
        P.decode_node_devices("a,b,c:")                            # not 7 fields
    assert P.registered_mem(183359 << 20, 1.0) == 183359 and P.registered_mem(183359 << 20, 2.0) == 366718
    assert P.registered_cores(1.0) == 100


def test_next_request_and_erase_walk_through_containers():
    a = "GPU-0,NVIDIA,4096,30:;"
    assert P.next_device_request(a) == (0, [CD("GPU-0", "NVIDIA", 4096, 30)])
    assert P.erase_next_device_request(a) == ";"
    with pytest.raises(LookupError):
        P.next_device_request(";")
    multi = ";GPU-1,NVIDIA,1,2:GPU-2,NVIDIA,3,4:;"                  # container 0 needs nothing, container 1 two GPUs
    assert P.next_device_request(multi)[0] == 1 and len(P.next_device_request(multi)[1]) == 2


def test_allocate_contract_matches_server_go(tmp_path):
    devs = [CD("GPU-fake-0", "NVIDIA", 8192, 30), CD("GPU-fake-1", "NVIDIA", 4096, 50)]
    envs, mounts, cdir = P.allocate(devs, 2, "/usr/local", "uid-1", "main", cache_uuid="c0ffee", device_memory_scaling=1.0)
    assert list(envs.items()) == [
        ("NVIDIA_VISIBLE_DEVICES", "GPU-fake-0,GPU-fake-1"),
        ("CUDA_DEVICE_MEMORY_LIMIT_0", "8192m"), ("CUDA_DEVICE_MEMORY_LIMIT_1", "4096m"),
        ("CUDA_DEVICE_SM_LIMIT", "30"),                               # cores of device 0 only (server.go:354)
        ("CUDA_DEVICE_MEMORY_SHARED_CACHE", "/usr/local/vgpu/c0ffee.cache")]
    # server_test.go:174-184: /usr/local/vgpu/containers/<uid>_<ctr>
    assert cdir == "/usr/local/vgpu/containers/uid-1_main"
    assert mounts == [("/usr/local/vgpu/libvgpu.so", "/usr/local/vgpu/libvgpu.so", True),
                      ("/usr/local/vgpu", "/usr/local/vgpu/containers/uid-1_main", False),
                      ("/tmp/vgpulock", "/tmp/vgpulock", False),
                      ("/etc/ld.so.preload", "/usr/local/vgpu/ld.so.preload", True)]
    envs, mounts, _ = P.allocate(devs[:1], 1, "/usr/local", "u", "c", device_memory_scaling=1.5, disable_core_limit=True,
                                 container_sets_disable_control=True, license_present=True)
    assert envs["CUDA_OVERSUBSCRIBE"] == "true" and envs["GPU_CORE_UTILIZATION_POLICY"] == "disable"
    assert ("/etc/ld.so.preload", "/usr/local/vgpu/ld.so.preload", True) not in mounts
    assert mounts[-2:] == [("/vgpu/", "/usr/local/vgpu/license", True), ("/usr/bin/vgpuvalidator", "/usr/local/vgpu/vgpuvalidator", True)]
    with pytest.raises(ValueError):
        P.allocate(devs, 1, "/usr/local", "u", "c")                  # device allocate number not matched


def test_the_env_contract_is_what_the_hook_parses():
    import k8s_device_plugin_b200 as v
    envs, _, _ = P.allocate([CD("GPU-x", "NVIDIA", 8192, 30)], 1, "/usr/local", "u", "c")
    assert v.parse_limit(envs["CUDA_DEVICE_MEMORY_LIMIT_0"]) == 8192 << 20


@pytest.fixture
def cluster(tmp_path):
    sock_dir = str(tmp_path / "device-plugins")
    kubelet = KubeletStub(sock_dir)
    kubelet.start()
    pod = S.Pod(UID="pod-uid-1", Name="p", Containers=[S.Container("main")], Annotations={
        S.ASSIGNED_NODE: "node-0", S.BIND_TIME: "1", S.BIND_PHASE: S.BIND_ALLOCATING,
        S.TO_ALLOCATE: "GPU-fake-0,NVIDIA,8192,30:;"})
    pods = S.InMemoryPodSource([pod])
    plugin = S.NvidiaDevicePlugin([S.GpuDevice("GPU-fake-0"), S.GpuDevice("GPU-fake-1")], pods, node_name="node-0",
                                  socket_dir=sock_dir, host_hook_path=str(tmp_path / "hook"), device_split_count=10)
    plugin.Start()
    assert kubelet.registered.wait(5)
    yield kubelet, plugin, pods, pod
    plugin.Stop()
    kubelet.stop()


def test_kubelet_stub_round_trip(cluster, tmp_path):
    kubelet, plugin, pods, pod = cluster
    t0 = time.perf_counter()
    r = kubelet.request
    assert (r.version, r.resource_name, r.endpoint) == ("v1beta1", "nvidia.com/gpu", "nvidia-gpu.sock")
    assert r.options.get_preferred_allocation_available is True
    lw = kubelet.list_and_watch_once()[0]
    ids = [d.ID for d in lw.devices]
    assert ids == [f"GPU-fake-{g}-{i}" for g in range(2) for i in range(10)]      # 2 x split count
    assert {d.health for d in lw.devices} == {"Healthy"}
    resp = kubelet.allocate([["GPU-fake-0-3"]])
    c = resp.container_responses[0]
    hook = str(tmp_path / "hook")
    assert c.envs["NVIDIA_VISIBLE_DEVICES"] == "GPU-fake-0"
    assert c.envs["CUDA_DEVICE_MEMORY_LIMIT_0"] == "8192m" and c.envs["CUDA_DEVICE_SM_LIMIT"] == "30"
    assert c.envs["CUDA_DEVICE_MEMORY_SHARED_CACHE"].startswith(hook + "/vgpu/") and c.envs["CUDA_DEVICE_MEMORY_SHARED_CACHE"].endswith(".cache")
    assert "CUDA_OVERSUBSCRIBE" not in c.envs
    assert [(m.container_path, m.host_path, m.read_only) for m in c.mounts] == [
        (hook + "/vgpu/libvgpu.so", hook + "/vgpu/libvgpu.so", True),
        (hook + "/vgpu", hook + "/vgpu/containers/pod-uid-1_main", False),
        ("/tmp/vgpulock", "/tmp/vgpulock", False),
        ("/etc/ld.so.preload", hook + "/vgpu/ld.so.preload", True)]
    assert pod.Annotations[S.TO_ALLOCATE] == ";"                    # EraseNextDeviceTypeFromAnnotation
    assert pod.Annotations[S.BIND_PHASE] == S.BIND_SUCCESS and pods.lock_released == 1
    print("round trip ms:", (time.perf_counter() - t0) * 1e3)


def test_the_other_kubelet_calls_answer_like_the_reference(cluster):
    """server.go:245-250 (options: preferred allocation advertised), :270-288 (the answer is nonetheless empty: the body is
    commented out upstream), :501-503 (PreStartContainer: empty)."""
    kubelet, plugin, pods, pod = cluster
    o = kubelet.options()
    assert o.get_preferred_allocation_available is True and o.pre_start_required is False
    r = kubelet.preferred_allocation([f"GPU-fake-0-{i}" for i in range(4)], ["GPU-fake-0-1"], 2)
    assert len(r.container_responses) == 0
    assert kubelet.pre_start(["GPU-fake-0-1"]).SerializeToString() == b""


def test_allocate_failure_paths(cluster):
    kubelet, plugin, pods, pod = cluster
    import grpc
    with pytest.raises(grpc.RpcError) as ei:
        kubelet.allocate([["GPU-fake-0-0", "GPU-fake-0-1"]])       # kubelet asks for 2, scheduler granted 1
    assert "device allocate number not matched" in ei.value.details()
    assert pod.Annotations[S.BIND_PHASE] == S.BIND_FAILED
    with pytest.raises(grpc.RpcError) as ei:
        kubelet.allocate([["GPU-fake-0-0"]])                        # no pod is in bind-phase allocating any more
    assert "no binding pod found" in ei.value.details()


def test_health_event_resends_device_list(cluster):
    kubelet, plugin, pods, pod = cluster
    plugin.mark_unhealthy(plugin.devices[1])
    lw = kubelet.list_and_watch_once(n=2)
    assert [d.health for d in lw[1].devices] == ["Healthy"] * 10 + ["Unhealthy"] * 10


def test_node_registration_annotation(cluster):
    _, plugin, _, _ = cluster
    a = plugin.node_annotations(now="T")
    assert a[S.HANDSHAKE] == "Reported T"
    assert a[S.REGISTER] == "GPU-fake-0,10,183359,100,NVIDIA-NVIDIA B200,0,true:GPU-fake-1,10,183359,100,NVIDIA-NVIDIA B200,0,true:"


def test_native_codec_equals_the_python_restatement_on_random_annotations():
    """Differential test of csrc/plugin_core.cc against oracle/codec_oracle.py (util.go:78-271 restated): well-formed,
    ragged and malformed annotations — same values, same errors, same re-encoded bytes."""
    import os
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import codec_oracle as O
    from k8s_device_plugin_b200.plugin import core
    rng = random.Random(77)
    word = lambda: "".join(rng.choice("abcXYZ-019 _") for _ in range(rng.randint(0, 8)))
    num = lambda: rng.choice(["0", "1", "10", "8192", "183359", "-5", "2147483647", "2147483648", "99999999999", "", "x1", "1x", "+7", " 3"])

    def node_text():
        parts = []
        for _ in range(rng.randint(0, 4)):
            n = rng.choice([7, 7, 7, 6, 8, 1])
            fields = [word(), num(), num(), num(), word(), num(), rng.choice(["true", "false", "1", "T", "yes", ""])][:n] + ["z"] * max(0, n - 7)
            parts.append(",".join(fields))
        return ":".join(parts) + rng.choice([":", "", "::"])

    def pod_text():
        ctrs = []
        for _ in range(rng.randint(0, 4)):
            devs = []
            for _ in range(rng.randint(0, 3)):
                n = rng.choice([4, 4, 4, 3, 5, 1])
                devs.append(",".join([word(), word(), num(), num(), "extra"][:n]))
            ctrs.append(":".join(devs) + rng.choice([":", ""]))
        return ";".join(ctrs) + rng.choice([";", "", ";;"])

    checked = errors = 0
    for _ in range(1500):
        t = node_text()
        try:
            want = O.decode_node_devices(t)
        except O.CodecError:
            with pytest.raises(core.CodecError):
                core.decode_node_devices(t)
            errors += 1
        else:
            got = core.decode_node_devices(t)
            assert [(d.Id, d.Count, d.Devmem, d.Devcore, d.Type, d.Numa, d.Health) for d in got] == \
                [(d["Id"], d["Count"], d["Devmem"], d["Devcore"], d["Type"], d["Numa"], d["Health"]) for d in want], repr(t)
            assert core.encode_node_devices(got) == O.encode_node_devices(want)
        t = pod_text()
        try:
            want = O.decode_pod_single_device(t)
        except O.CodecError:
            with pytest.raises(core.CodecError):
                core.decode_pod_single_device(t)
            errors += 1
        else:
            got = core.decode_pod_single_device(t)
            assert [[(d.UUID, d.Type, d.Usedmem, d.Usedcores) for d in c] for c in got] == \
                [[(d["UUID"], d["Type"], d["Usedmem"], d["Usedcores"]) for d in c] for c in want], repr(t)
            assert core.encode_pod_single_device(got) == O.encode_pod_single_device(want)
            assert core.erase_next_device_request(t) == O.erase_next_device_request(t), repr(t)
            try:
                wi, wd = O.next_device_request(t)
            except LookupError:
                with pytest.raises(LookupError):
                    core.next_device_request(t)
            else:
                gi, gd = core.next_device_request(t)
                assert gi == wi and [(d.UUID, d.Usedmem) for d in gd] == [(d["UUID"], d["Usedmem"]) for d in wd]
        checked += 1
    assert checked == 1500 and errors > 100
