"""Scheduler extender + webhook against the reference's own test (pkg/scheduler/scheduler_test.go:27-99)
and hand-derived cases of score.go / device.go. CPU only."""
import base64
import http.client
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import k8s_device_plugin_b200 as v  # noqa: E402
from k8s_device_plugin_b200.plugin import core, scheduler as S  # noqa: E402
from k8s_device_plugin_b200.plugin import server as P  # noqa: E402


def _node(name, devices, handshake=None):
    annos = {P.REGISTER: core.encode_node_devices(devices)}
    if handshake:
        annos[P.HANDSHAKE] = handshake
    return {"metadata": {"name": name, "annotations": annos}}


def _pod(name, containers, annos=None, uid=None, ns="default"):
    return {"metadata": {"name": name, "namespace": ns, "uid": uid or f"uid-{name}", "annotations": dict(annos or {})},
            "spec": {"containers": containers}, "status": {"phase": "Pending"}}


def _ctr(name="c", **limits):
    return {"name": name, "resources": {"limits": {k.replace("_", "/", 1).replace("__", "-"): val for k, val in limits.items()}}}


def _gpu_ctr(name="c", gpu=1, mem=None, pct=None, cores=None, priority=None):
    lim = {P.RESOURCE_NAME: str(gpu)}
    if mem is not None:
        lim[P.RESOURCE_MEM] = str(mem)
    if pct is not None:
        lim[P.RESOURCE_MEM_PERCENTAGE] = str(pct)
    if cores is not None:
        lim[P.RESOURCE_CORES] = str(cores)
    if priority is not None:
        lim["vgputaskpriority"] = str(priority)
    return {"name": name, "resources": {"limits": lim}}


def _b200(i, numa=0, count=10, mem=183359, typ="NVIDIA-NVIDIA B200"):
    return core.NodeDevice(f"GPU-{i:04d}", count, mem, 100, typ, numa, True)


def test_get_nodes_usage_matches_reference_golden():
    """scheduler_test.go:27-99: two pods each holding GPU0 (100 MiB, 10 cores) -> Used 2, Usedmem 200, Usedcores 20."""
    s = S.Scheduler()
    s.add_node("node1", S.NodeInfo("node1", [S.DeviceInfo("GPU0", 0, 10, 1024, 100, "", 1, True), S.DeviceInfo("GPU1", 1, 10, 1024, 100, "", 1, True)]))
    devs = {"NVIDIA": [[core.ContainerDevice("GPU0", "", 100, 10)]]}
    s.add_pod({"metadata": {"uid": "1111", "name": "test1", "namespace": "default"}}, "node1", devs)
    s.add_pod({"metadata": {"uid": "2222", "name": "test2", "namespace": "default"}}, "node1", devs)
    cache, failed = s.get_nodes_usage(["node1"], None)
    assert len(cache) == 1 and "node1" in cache and failed == {}
    d = cache["node1"].Devices
    assert len(d) == 2
    assert (d[0].Used, d[0].Usedmem, d[0].Usedcores) == (2, 200, 20)
    assert (d[1].Used, d[1].Usedmem, d[1].Usedcores) == (0, 0, 0)
    assert s.get_nodes_usage(["ghost"], None)[1] == {"ghost": "node unregisterd"}


def test_quantity_accessors():
    assert S.quantity_as_int64("8192") == (8192, True)
    assert S.quantity_as_int64("1k") == (1000, True) and S.quantity_as_int64("1Ki") == (1024, True)
    assert S.quantity_as_int64("2Gi") == (2 << 30, True) and S.quantity_as_int64("1e3") == (1000, True)
    assert S.quantity_as_int64("100m") == (0, False) and S.quantity_as_int64("1500m") == (0, False)
    assert S.quantity_as_int64("2000m") == (2, True)
    assert S.quantity_as_int64("abc") == (0, False)
    assert S.quantity_value("100m") == 1 and S.quantity_value("1") == 1 and S.quantity_value("0") == 0


def test_generate_resource_requests_defaults():
    cfg = S.Config()
    r = S.generate_resource_requests(_gpu_ctr(gpu=2, mem=8192, cores=30), cfg)
    assert (r.Nums, r.Type, r.Memreq, r.MemPercentagereq, r.Coresreq) == (2, "NVIDIA", 8192, 101, 30)
    r = S.generate_resource_requests(_gpu_ctr(gpu=1), cfg)               # nothing given: 100 % of the card
    assert (r.Memreq, r.MemPercentagereq, r.Coresreq) == (0, 100, 0)
    r = S.generate_resource_requests(_gpu_ctr(gpu=1), S.Config(DefaultMem=5000, DefaultCores=10))
    assert (r.Memreq, r.MemPercentagereq, r.Coresreq) == (5000, 101, 10)
    r = S.generate_resource_requests(_gpu_ctr(gpu=1, pct=50), cfg)
    assert (r.Memreq, r.MemPercentagereq) == (0, 50)
    assert S.generate_resource_requests({"name": "x"}, cfg).Nums == 0
    only_requests = {"name": "x", "resources": {"requests": {P.RESOURCE_NAME: "1", P.RESOURCE_MEM: "100"}}}
    assert S.generate_resource_requests(only_requests, cfg).Memreq == 100
    reqs = S.resource_reqs(_pod("p", [_gpu_ctr(), {"name": "sidecar"}]), cfg)
    assert len(reqs) == 2 and list(reqs[0]) == ["NVIDIA"] and reqs[1] == {}


def _usage(n=4, **kw):
    return S.NodeUsage([S.DeviceUsage(f"GPU-{i}", i, 0, 10, 0, 1000, 100, 0, kw.get("numa", [0] * n)[i], kw.get("type", "NVIDIA-B200")) for i in range(n)])


def _req(nums=1, mem=0, pct=101, cores=0, typ="NVIDIA"):
    return S.ContainerDeviceRequest(nums, typ, mem, pct, cores)


def test_fit_walks_from_the_most_free_device_and_charges_it():
    u = _usage(3)
    u.Devices[0].Used, u.Devices[1].Used = 3, 1                  # free slots: 7, 9, 10 -> sorted ascending, walked from the end
    fit, score, devs = S.score_node(u, [_req(1, mem=100, cores=30)], {})
    assert fit and [d.UUID for d in devs["NVIDIA"][0]] == ["GPU-2"]
    assert devs["NVIDIA"][0][0] == core.ContainerDevice("GPU-2", "NVIDIA", 100, 30)
    assert score == pytest.approx(10 / 10 + (3 - 1))             # total/free + (len(devices) - sums), float32
    charged = next(d for d in u.Devices if d.Id == "GPU-2")
    assert (charged.Used, charged.Usedmem, charged.Usedcores) == (1, 100, 30)


def test_fit_rules_of_fit_in_certain_device():
    # memory percentage is taken of EACH card's total (score.go:118-121)
    u = _usage(1)
    fit, _, devs = S.score_node(u, [_req(1, pct=50)], {})
    assert fit and devs["NVIDIA"][0][0].Usedmem == 500
    # insufficient memory / cores
    u = _usage(1); u.Devices[0].Usedmem = 950
    assert not S.score_node(u, [_req(1, mem=100)], {})[0]
    u = _usage(1); u.Devices[0].Usedcores = 80
    assert not S.score_node(u, [_req(1, mem=1, cores=30)], {})[0]
    # cores=100 wants the card alone
    u = _usage(1); u.Devices[0].Used = 1
    assert not S.score_node(u, [_req(1, mem=1, cores=100)], {})[0]
    # cores=0 job cannot land on a card whose cores are all given away
    u = _usage(1); u.Devices[0].Used, u.Devices[0].Usedcores = 1, 100
    assert not S.score_node(u, [_req(1, mem=1, cores=0)], {})[0]
    # no free share
    u = _usage(1); u.Devices[0].Used = 10
    assert not S.score_node(u, [_req(1, mem=1)], {})[0]
    # cores > 100 is refused outright; more cards than the node has too
    assert not S.score_node(_usage(2), [_req(1, mem=1, cores=101)], {})[0]
    assert not S.score_node(_usage(2), [_req(3, mem=1)], {})[0]
    # equality is enough (strict "<" comparisons)
    u = _usage(1); u.Devices[0].Usedmem, u.Devices[0].Usedcores = 900, 70
    assert S.score_node(u, [_req(1, mem=100, cores=30)], {})[0]
    # unknown vendor / type mismatch
    assert not S.score_node(_usage(1), [_req(1, mem=1, typ="MLU")], {})[0]
    assert not S.score_node(_usage(1, type="DCU-Z100"), [_req(1, mem=1)], {})[0]


def test_gpu_type_annotations():
    u = lambda: _usage(2, type="NVIDIA-NVIDIA B200")
    assert S.score_node(u(), [_req(1, mem=1)], {S.GPU_IN_USE: "b200"})[0]          # case-insensitive substring
    assert not S.score_node(u(), [_req(1, mem=1)], {S.GPU_IN_USE: "A100"})[0]
    assert S.score_node(u(), [_req(1, mem=1)], {S.GPU_IN_USE: "A100,B200"})[0]
    assert not S.score_node(u(), [_req(1, mem=1)], {S.GPU_NO_USE: "B200"})[0]
    assert S.score_node(u(), [_req(1, mem=1)], {S.GPU_NO_USE: "A100,H100"})[0]
    assert S.score_node(u(), [_req(1, mem=1)], {S.GPU_IN_USE: "B200", S.GPU_NO_USE: "B200"})[0]   # use-list wins


def test_numa_bind_restarts_on_every_numa_boundary():
    # 4 cards: numa 0,0,1,1 ; cards on numa 1 are nearly full so two cards only fit on numa 0
    u = _usage(4, numa=[0, 0, 1, 1])
    u.Devices[3].Used = 10
    fit, _, devs = S.score_node(u, [_req(2, mem=1)], {S.NUMA_BIND: "true"})
    assert fit and sorted(d.UUID for d in devs["NVIDIA"][0]) == ["GPU-0", "GPU-1"]
    # without the binding the request is allowed to straddle
    u = _usage(4, numa=[0, 0, 1, 1]); u.Devices[3].Used = 10; u.Devices[0].Used = 10
    fit, _, devs = S.score_node(u, [_req(2, mem=1)], {})
    assert fit and sorted(d.UUID for d in devs["NVIDIA"][0]) == ["GPU-1", "GPU-2"]
    u = _usage(4, numa=[0, 0, 1, 1]); u.Devices[3].Used = 10; u.Devices[0].Used = 10
    assert not S.score_node(u, [_req(2, mem=1)], {S.NUMA_BIND: "true"})[0]
    assert S.score_node(_usage(4, numa=[0, 0, 1, 1]), [_req(2, mem=1)], {S.NUMA_BIND: "notabool"})[0]


def test_container_bookkeeping_reference_mode_and_fixed_mode():
    """score.go:222 compares the number of vendor lists with the number of containers; score.go:211 indexes past the slice."""
    two = [_req(1, mem=1), _req(1, mem=1)]
    assert not S.score_node(_usage(2), two, {}, mode=0)[0]                       # two GPU containers never fit in the reference
    fit, _, devs = S.score_node(_usage(2), two, {}, mode=1)
    assert fit and [len(c) for c in devs["NVIDIA"]] == [1, 1]
    with pytest.raises(S.SchedulerPanic):
        S.score_node(_usage(2), [_req(1, mem=1), _req(0)], {}, mode=0)           # GPU container + sidecar: Go panics
    fit, _, devs = S.score_node(_usage(2), [_req(1, mem=1), _req(0)], {}, mode=1)
    assert fit and [len(c) for c in devs["NVIDIA"]] == [1, 0]
    assert not S.score_node(_usage(2), [_req(0), _req(1, mem=1)], {}, mode=0)[0]
    fit, _, devs = S.score_node(_usage(2), [_req(0), _req(1, mem=1)], {}, mode=1)
    assert fit and [len(c) for c in devs["NVIDIA"]] == [0, 1]


def _cluster():
    nodes = [_node("node-a", [_b200(0), _b200(1)]), _node("node-b", [_b200(10), _b200(11), _b200(12), _b200(13)])]
    pod = _pod("train", [_gpu_ctr("main", gpu=1, mem=8192, cores=30)])
    kube = S.InMemoryKube(nodes, [pod])
    s = S.Scheduler(kube, S.Config(SchedulerName="4pd-scheduler"))
    s.register_from_node_annotations_once(now=1_700_000_000)
    return kube, s, pod


def test_registration_handshake_and_node_leave():
    kube, s, _ = _cluster()
    assert sorted(s.nodes) == ["node-a", "node-b"] and [d.ID for d in s.nodes["node-b"].Devices] == ["GPU-0010", "GPU-0011", "GPU-0012", "GPU-0013"]
    hs = kube.nodes["node-a"]["metadata"]["annotations"][P.HANDSHAKE]
    assert hs.startswith("Requesting_")
    # the plugin never answers ("Reported ..."): after 60 s the node's devices are dropped and the handshake says Deleted
    s.register_from_node_annotations_once(now=1_700_000_000 + 30)
    assert len(s.nodes["node-a"].Devices) == 2
    s.register_from_node_annotations_once(now=1_700_000_000 + 61)
    assert kube.nodes["node-b"]["metadata"]["annotations"][P.HANDSHAKE].startswith("Deleted_")
    # plugin reports again -> scheduler re-requests and refreshes (devices are already known, mem/cores updated in place)
    kube.nodes["node-a"]["metadata"]["annotations"][P.HANDSHAKE] = "Reported 2023"
    kube.nodes["node-a"]["metadata"]["annotations"][P.REGISTER] = core.encode_node_devices([_b200(0, mem=1000), _b200(1)])
    s.nodes.pop("node-a", None)
    s.register_from_node_annotations_once(now=1_700_000_000 + 100)
    assert [d.Devmem for d in s.nodes["node-a"].Devices] == [1000, 183359]


def test_filter_picks_the_highest_score_and_writes_the_annotations_allocate_consumes(tmp_path):
    kube, s, pod = _cluster()
    res = s.filter({"Pod": pod, "NodeNames": ["node-a", "node-b", "ghost"]})
    # score = total/free + (len(devices) - nums): node-b has more idle cards -> higher
    assert res == S.filter_result(node_names=["node-b"])
    annos = kube.pods[("default", "train")]["metadata"]["annotations"]
    assert annos[S.ASSIGNED_NODE] == "node-b" and annos[S.ASSIGNED_TIME].isdigit()
    assert annos[P.TO_ALLOCATE] == annos[P.ALLOCATED] == "GPU-0013,NVIDIA,8192,30:;"
    assert s.pods["uid-train"].NodeID == "node-b"
    # the next pod sees the first one's share
    pod2 = _pod("train2", [_gpu_ctr("main", gpu=1, mem=8192, cores=80)])
    kube.pods[("default", "train2")] = pod2
    s.filter({"Pod": pod2, "NodeNames": ["node-b"]})
    assert kube.pods[("default", "train2")]["metadata"]["annotations"][P.TO_ALLOCATE].split(",")[0] != "GPU-0013"
    cache, _ = s.get_nodes_usage(["node-b"], None)
    assert sorted((d.Id, d.Used, d.Usedmem, d.Usedcores) for d in cache["node-b"].Devices if d.Used) == \
        [("GPU-0012", 1, 8192, 80), ("GPU-0013", 1, 8192, 30)]

    # bind: node lock + allocating phase, then the device plugin on that node finds the pod and answers Allocate with the
    # envs/mounts the hook consumes — scheduler -> plugin -> hook contract end to end
    assert s.bind({"PodName": "train", "PodNamespace": "default", "PodUID": "uid-train", "Node": "node-b"}) == {"Error": ""}
    annos = kube.pods[("default", "train")]["metadata"]["annotations"]
    assert annos[P.BIND_PHASE] == "allocating" and annos[P.BIND_TIME].isdigit()
    assert S.NODE_LOCK_TIME in kube.nodes["node-b"]["metadata"]["annotations"]
    assert kube.bindings == [("default", "train", "uid-train", "node-b")]
    idx, devs = core.next_device_request(annos[P.TO_ALLOCATE])
    assert idx == 0 and devs == [core.ContainerDevice("GPU-0013", "NVIDIA", 8192, 30)]
    envs, _mounts, _dir = core.allocate(devs, 1, str(tmp_path), "uid-train", "main")
    assert envs["CUDA_DEVICE_MEMORY_LIMIT_0"] == "8192m" and envs["CUDA_DEVICE_SM_LIMIT"] == "30"
    assert envs["NVIDIA_VISIBLE_DEVICES"] == "GPU-0013"
    # a second bind inside five minutes finds the node locked, logs it, and binds anyway (scheduler.go:329-332)
    assert s.bind({"PodName": "train2", "PodNamespace": "default", "PodUID": "uid-train2", "Node": "node-b"}) == {"Error": ""}
    assert s.bind({"PodName": "train2", "PodNamespace": "default", "PodUID": "uid-train2", "Node": "nowhere"})["Error"] != ""


def test_filter_passthrough_failure_and_patch_error():
    kube, s, pod = _cluster()
    plain = _pod("web", [{"name": "nginx"}])
    assert s.filter({"Pod": plain, "NodeNames": ["node-a"]}) == S.filter_result(node_names=["node-a"])
    big = _pod("big", [_gpu_ctr(gpu=1, mem=999999)])
    kube.pods[("default", "big")] = big
    assert s.filter({"Pod": big, "NodeNames": ["node-a", "ghost"]}) == S.filter_result(failed_nodes={"ghost": "node unregisterd"})
    kube.fail_patch = True
    with pytest.raises(RuntimeError):
        s.filter({"Pod": pod, "NodeNames": ["node-a"]})
    assert "uid-train" not in s.pods                                              # rolled back (scheduler.go:401-404)


def test_informer_callbacks_track_scheduled_pods():
    s = S.Scheduler()
    annos = {S.ASSIGNED_NODE: "node-a", P.ALLOCATED: "GPU-0000,NVIDIA,100,10:;"}
    p = _pod("x", [_gpu_ctr()], annos)
    s.on_add_pod(p)
    # decode keeps the reference's trailing empty container
    assert s.pods["uid-x"].Devices == {"NVIDIA": [[core.ContainerDevice("GPU-0000", "NVIDIA", 100, 10)], []]}
    s.on_add_pod(_pod("nope", [_gpu_ctr()]))
    assert "uid-nope" not in s.pods
    done = _pod("x", [_gpu_ctr()], annos); done["status"]["phase"] = "Succeeded"
    s.on_update_pod(p, done)
    assert s.pods == {}
    s.on_add_pod(p); s.on_del_pod(p)
    assert s.pods == {}


def test_node_lock_protocol():
    kube = S.InMemoryKube([{"metadata": {"name": "n", "annotations": {}}}])
    S.lock_node(kube, "n", now=1000)
    assert kube.nodes["n"]["metadata"]["annotations"][S.NODE_LOCK_TIME] == "1970-01-01T00:16:40Z"
    with pytest.raises(RuntimeError, match="locked within 5 minutes"):
        S.lock_node(kube, "n", now=1000 + 299)
    S.lock_node(kube, "n", now=1000 + 301)                                        # expired: stolen
    assert kube.nodes["n"]["metadata"]["annotations"][S.NODE_LOCK_TIME] == "1970-01-01T00:21:41Z"
    with pytest.raises(RuntimeError, match="is locked"):
        S.set_node_lock(kube, "n")
    S.release_node_lock(kube, "n"); S.release_node_lock(kube, "n")
    assert S.NODE_LOCK_TIME not in kube.nodes["n"]["metadata"]["annotations"]


def _review(pod):
    return {"apiVersion": "admission.k8s.io/v1", "kind": "AdmissionReview", "request": {"uid": "r-1", "namespace": "default", "name": pod["metadata"]["name"], "object": pod}}


def _patch(resp):
    return json.loads(base64.b64decode(resp["response"]["patch"]))


def test_webhook_mutation():
    cfg = S.Config(SchedulerName="4pd-scheduler")
    r = S.webhook_handle(_review(_pod("p", [_gpu_ctr("main", priority=1), {"name": "side"}])), cfg)
    assert r["response"]["allowed"] and r["response"]["uid"] == "r-1" and r["response"]["patchType"] == "JSONPatch"
    assert _patch(r) == [{"op": "add", "path": "/spec/containers/0/env", "value": [{"name": "CUDA_TASK_PRIORITY", "value": "1"}]},
                         {"op": "add", "path": "/spec/schedulerName", "value": "4pd-scheduler"}]
    with_env = _gpu_ctr("main", priority=0); with_env["env"] = [{"name": "A", "value": "b"}]
    p = _pod("p", [with_env]); p["spec"]["schedulerName"] = "default-scheduler"
    assert _patch(S.webhook_handle(_review(p), cfg)) == [
        {"op": "add", "path": "/spec/containers/0/env/-", "value": {"name": "CUDA_TASK_PRIORITY", "value": "0"}},
        {"op": "replace", "path": "/spec/schedulerName", "value": "4pd-scheduler"}]
    # Go's short-circuit: after the first container that asks for the resource, later containers are not mutated
    two = _pod("p", [_gpu_ctr("a", priority=1), _gpu_ctr("b", priority=0)])
    assert _patch(S.webhook_handle(_review(two), cfg))[:-1] == [{"op": "add", "path": "/spec/containers/0/env", "value": [{"name": "CUDA_TASK_PRIORITY", "value": "1"}]}]
    # ... but a container WITHOUT the resource in front of it is still visited (and gets its priority env)
    mixed = _pod("p", [{"name": "side", "resources": {"limits": {"vgputaskpriority": "1"}}}, _gpu_ctr("b", priority=0)])
    assert [op["path"] for op in _patch(S.webhook_handle(_review(mixed), cfg))] == ["/spec/containers/0/env", "/spec/containers/1/env", "/spec/schedulerName"]
    r = S.webhook_handle(_review(_pod("p", [{"name": "nginx"}])), cfg)
    assert r["response"]["allowed"] and "patch" not in r["response"] and r["response"]["status"]["message"] == "no resource found"
    priv = _gpu_ctr("main"); priv["securityContext"] = {"privileged": True}
    assert "patch" not in S.webhook_handle(_review(_pod("p", [priv])), cfg)["response"]       # privileged containers are skipped
    r = S.webhook_handle(_review(_pod("p", [])), cfg)
    assert not r["response"]["allowed"] and r["response"]["status"]["message"] == "pod has no containers"
    assert _patch(S.webhook_handle(_review(_pod("p", [_gpu_ctr()])), S.Config())) == []        # no scheduler name configured


def test_http_routes_and_metrics():
    kube, s, pod = _cluster()
    srv = S.serve(s, "127.0.0.1:0")
    try:
        port = srv.server_address[1]

        def post(path, body):
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=10)
            c.request("POST", path, body if isinstance(body, (bytes, str)) else json.dumps(body), {"Content-Type": "application/json"})
            r = c.getresponse()
            return r.status, r.getheader("Content-Type"), r.read()

        st, ct, body = post("/filter", {"Pod": pod, "NodeNames": ["node-a", "node-b"]})
        assert st == 200 and ct == "application/json"
        assert json.loads(body) == {"Nodes": None, "NodeNames": ["node-b"], "FailedNodes": None, "FailedAndUnresolvableNodes": None, "Error": ""}
        st, _, body = post("/filter", "{not json")
        assert st == 200 and json.loads(body)["Error"] != ""
        st, _, body = post("/bind", {"PodName": "train", "PodNamespace": "default", "PodUID": "uid-train", "Node": "node-b"})
        assert st == 200 and json.loads(body) == {"Error": ""}
        st, _, body = post("/webhook", _review(_pod("q", [_gpu_ctr()])))
        assert st == 200 and json.loads(body)["response"]["allowed"] is True
        # reference mode: GPU container + sidecar panics inside the handler -> connection dropped, no response
        side = _pod("side", [_gpu_ctr(mem=10), {"name": "sidecar"}]); kube.pods[("default", "side")] = side
        with pytest.raises((http.client.RemoteDisconnected, ConnectionError)):
            post("/filter", {"Pod": side, "NodeNames": ["node-a"]})
        s.cfg.MultiContainer = True
        st, _, body = post("/filter", {"Pod": side, "NodeNames": ["node-a"]})
        assert json.loads(body)["NodeNames"] == ["node-a"]
        assert kube.pods[("default", "side")]["metadata"]["annotations"][P.TO_ALLOCATE].endswith(",NVIDIA,10,0:;")

        c = http.client.HTTPConnection("127.0.0.1", port, timeout=10)
        c.request("GET", "/metrics")
        text = c.getresponse().read().decode()
        assert 'GPUDeviceMemoryLimit{deviceidx="3",deviceuuid="GPU-0013",nodeid="node-b",zone="vGPU"} 1.92266e+11' in text
        assert 'GPUDeviceSharedNum{deviceidx="3",deviceuuid="GPU-0013",nodeid="node-b",zone="vGPU"} 1' in text
        assert 'vGPUCorePercentage{containeridx="NVIDIA",deviceuuid="GPU-0013",nodename="node-b",podname="train",podnamespace="default",zone="vGPU"} 30' in text
        for name in ("GPUDeviceCoreLimit", "GPUDeviceMemoryAllocated", "GPUDeviceCoreAllocated", "nodeGPUOverview", "nodeGPUMemoryPercentage",
                     "vGPUPodsDeviceAllocated", "vGPUMemoryPercentage"):
            assert f"# TYPE {name} gauge" in text
    finally:
        srv.shutdown()


def test_native_scoring_core_equals_the_python_restatement_on_random_clusters():
    """Differential test: csrc/sched_core.cc against oracle/sched_oracle.py (an independent statement-level restatement
    of score.go) on random nodes, usages, requests and annotations — same fit decision, same float32 score, same
    devices in the same order, same charged usage."""
    import random
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import sched_oracle as O
    rng = random.Random(0xB200)
    types = ["NVIDIA-NVIDIA B200", "NVIDIA-NVIDIA A100-SXM4-80GB", "NVIDIA-Tesla P4", "MLU-370"]
    outcomes = {"fit": 0, "nofit": 0, "panic": 0}
    for case in range(600):
        ndev = rng.randint(1, 9)
        devs = []
        for i in range(ndev):
            count = rng.choice([1, 2, 4, 10])
            used = rng.randint(0, count)
            totalmem = rng.choice([16384, 81920, 183359])
            devs.append(dict(Id=f"GPU-{case}-{i}", Index=i, Used=used, Count=count, Usedmem=rng.randint(0, totalmem) if used else 0, Totalmem=totalmem,
                             Totalcore=rng.choice([100, 100, 200, 0]), Usedcores=rng.choice([0, 10, 50, 100]) if used else 0, Numa=rng.randint(0, 1),
                             Type=rng.choice(types), Health=True))
        nctr = rng.choice([1, 1, 1, 2, 3])
        reqs = []
        for _ in range(nctr):
            if rng.random() < 0.15:
                reqs.append(None)
                continue
            mem = rng.choice([0, 0, 1024, 8192, 100000])
            reqs.append(dict(Nums=rng.randint(1, 3), Type="NVIDIA", Memreq=mem, MemPercentagereq=101 if mem else rng.choice([101, 10, 50, 100]),
                             Coresreq=rng.choice([0, 0, 10, 30, 100, 101])))
        annos = {}
        r = rng.random()
        if r < 0.2:
            annos[O.GPU_IN_USE] = rng.choice(["B200", "a100,b200", "P4", ""])
        elif r < 0.4:
            annos[O.GPU_NO_USE] = rng.choice(["B200", "a100,p4", "V100"])
        if rng.random() < 0.3:
            annos[O.NUMA_BIND] = rng.choice(["true", "false", "1", "yes"])

        odevs = [dict(d) for d in devs]
        want = O.score_node(odevs, reqs, annos)
        usage = S.NodeUsage([S.DeviceUsage(d["Id"], d["Index"], d["Used"], d["Count"], d["Usedmem"], d["Totalmem"], d["Totalcore"], d["Usedcores"],
                                           d["Numa"], d["Type"], d["Health"]) for d in devs])
        creqs = [S.ContainerDeviceRequest() if q is None else S.ContainerDeviceRequest(q["Nums"], q["Type"], q["Memreq"], q["MemPercentagereq"], q["Coresreq"])
                 for q in reqs]
        outcomes[want[0]] += 1
        if want[0] == "panic":
            with pytest.raises(S.SchedulerPanic):
                S.score_node(usage, creqs, annos, mode=0)
            continue
        fit, score, devices = S.score_node(usage, creqs, annos, mode=0)
        assert fit == (want[0] == "fit"), (case, want)
        if fit:
            assert score == want[1], (case, score, want[1])
            got = [[(d.UUID, d.Type, d.Usedmem, d.Usedcores) for d in ctr] for ctr in devices["NVIDIA"]]
            exp = [[(d["UUID"], d["Type"], d["Usedmem"], d["Usedcores"]) for d in ctr if d] for ctr in want[2]]
            assert got == exp, (case, got, exp)
            assert [(d.Id, d.Used, d.Usedmem, d.Usedcores) for d in usage.Devices] == [(d["Id"], d["Used"], d["Usedmem"], d["Usedcores"]) for d in odevs], case
    assert outcomes["fit"] > 50 and outcomes["nofit"] > 50 and outcomes["panic"] > 5, outcomes
