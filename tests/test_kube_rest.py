"""RestKube against a scripted apiserver: the verbs, paths, content types and bodies the reference's client-go calls
produce (pkg/util/util.go:273-318 strategic-merge patches, scheduler.go:340 binding sub-resource, nodelock Update), the
list+watch pod informer, and the scheduler CLI end to end over real sockets. CPU only."""
import json
import os
import socket
import subprocess
import sys
import threading
import time
import http.client
import pytest
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import k8s_device_plugin_b200  # noqa: F401,E402
from k8s_device_plugin_b200.plugin import core, kube as K, scheduler as S  # noqa: E402
from k8s_device_plugin_b200.plugin import server as P  # noqa: E402
from conftest import ROOT  # noqa: E402


class FakeApiServer:
    def __init__(self, nodes=(), pods=()):
        self.nodes = {n["metadata"]["name"]: n for n in nodes}
        self.pods = {(p["metadata"]["namespace"], p["metadata"]["name"]): p for p in pods}
        self.log, self.bindings, self.rv = [], [], 100
        self.watchers, self.lock = [], threading.Lock()
        api = self

        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                pass

            def _json(self, code, obj):
                data = json.dumps(obj).encode()
                self.send_response(code); self.send_header("Content-Type", "application/json"); self.send_header("Content-Length", str(len(data)))
                self.end_headers(); self.wfile.write(data)

            def _body(self):
                n = int(self.headers.get("Content-Length") or 0)
                return json.loads(self.rfile.read(n)) if n else None

            def _route(self, method):
                path, _, query = self.path.partition("?")
                body = self._body() if method in ("POST", "PATCH", "PUT") else None
                api.log.append((method, self.path, self.headers.get("Content-Type"), self.headers.get("Authorization"), body))
                parts = path.strip("/").split("/")
                if method == "GET" and path == "/api/v1/nodes":
                    return self._json(200, {"items": list(api.nodes.values())})
                if path == "/api/v1/pods" and "watch=true" in query:
                    self.send_response(200); self.send_header("Content-Type", "application/json"); self.send_header("Transfer-Encoding", "chunked"); self.end_headers()
                    q = []
                    with api.lock:
                        api.watchers.append(q)
                    t0 = time.time()
                    try:
                        while time.time() - t0 < 5:
                            while q:
                                line = (json.dumps(q.pop(0)) + "\n").encode()
                                self.wfile.write(b"%x\r\n%s\r\n" % (len(line), line)); self.wfile.flush()
                            time.sleep(0.02)
                        self.wfile.write(b"0\r\n\r\n")
                    except OSError:
                        pass
                    self.close_connection = True
                    return
                if method == "GET" and path == "/api/v1/pods":
                    return self._json(200, {"metadata": {"resourceVersion": str(api.rv)}, "items": list(api.pods.values())})
                if parts[:3] == ["api", "v1", "nodes"] and len(parts) == 4:
                    name = parts[3]
                    if name not in api.nodes:
                        return self._json(404, {"kind": "Status", "message": f'nodes "{name}" not found', "code": 404})
                    if method == "GET":
                        return self._json(200, api.nodes[name])
                    if method == "PATCH":
                        api.nodes[name]["metadata"].setdefault("annotations", {}).update(body["metadata"]["annotations"])
                        return self._json(200, api.nodes[name])
                    if method == "PUT":
                        api.nodes[name] = body
                        return self._json(200, body)
                if parts[:3] == ["api", "v1", "namespaces"] and len(parts) >= 6 and parts[4] == "pods":
                    key = (parts[3], parts[5])
                    if key not in api.pods:
                        return self._json(404, {"kind": "Status", "message": f'pods "{parts[5]}" not found', "code": 404})
                    if len(parts) == 7 and parts[6] == "binding" and method == "POST":
                        api.bindings.append(body)
                        api.pods[key].setdefault("spec", {})["nodeName"] = body["target"]["name"]
                        return self._json(201, {"kind": "Status", "status": "Success"})
                    if method == "GET":
                        return self._json(200, api.pods[key])
                    if method == "PATCH":
                        api.pods[key]["metadata"].setdefault("annotations", {}).update(body["metadata"]["annotations"])
                        api.emit("MODIFIED", api.pods[key])
                        return self._json(200, api.pods[key])
                self._json(404, {"kind": "Status", "message": "not found", "code": 404})

            def do_GET(self): self._route("GET")
            def do_POST(self): self._route("POST")
            def do_PATCH(self): self._route("PATCH")
            def do_PUT(self): self._route("PUT")

        self.srv = ThreadingHTTPServer(("127.0.0.1", 0), H)
        self.srv.daemon_threads = True
        self.url = f"http://127.0.0.1:{self.srv.server_address[1]}"
        threading.Thread(target=self.srv.serve_forever, daemon=True).start()

    def emit(self, typ, obj):
        with self.lock:
            for q in self.watchers:
                q.append({"type": typ, "object": json.loads(json.dumps(obj))})

    def stop(self):
        self.srv.shutdown()


def _node(name, devs):
    return {"metadata": {"name": name, "annotations": {P.REGISTER: core.encode_node_devices(devs), P.HANDSHAKE: "Reported now"}}}


def _pod(name, limits, ns="default"):
    return {"metadata": {"name": name, "namespace": ns, "uid": "uid-" + name, "annotations": {}},
            "spec": {"containers": [{"name": "main", "resources": {"limits": limits}}]}, "status": {"phase": "Pending"}}


def _cluster():
    devs = [core.NodeDevice(f"GPU-{i}", 10, 183359, 100, "NVIDIA-NVIDIA B200", 0, True) for i in range(2)]
    return FakeApiServer([_node("node-a", devs)], [_pod("train", {P.RESOURCE_NAME: "1", P.RESOURCE_MEM: "8192", P.RESOURCE_CORES: "30"})])


def test_rest_verbs_match_the_references_client_calls():
    api = _cluster()
    try:
        k = K.RestKube(api.url, token="sekret")
        assert [n["metadata"]["name"] for n in k.list_nodes()] == ["node-a"]
        k.patch_pod_annotations("default", "train", {"a": "b"})
        m, path, ctype, auth, body = api.log[-1]
        assert (m, path, ctype, auth) == ("PATCH", "/api/v1/namespaces/default/pods/train", "application/strategic-merge-patch+json", "Bearer sekret")
        assert body == {"metadata": {"annotations": {"a": "b"}}}                      # patchPod{Metadata{Annotations}} util.go:300-308
        k.patch_node_annotations("node-a", {P.HANDSHAKE: "Requesting_x"})
        assert api.log[-1][:3] == ("PATCH", "/api/v1/nodes/node-a", "application/strategic-merge-patch+json")
        k.bind_pod("default", "train", "uid-train", "node-a")
        assert api.log[-1][1] == "/api/v1/namespaces/default/pods/train/binding"
        assert api.bindings[-1] == {"apiVersion": "v1", "kind": "Binding", "metadata": {"name": "train", "uid": "uid-train"}, "target": {"kind": "Node", "name": "node-a"}}
        S.lock_node(k, "node-a")                                                       # nodelock: GET + full PUT
        assert api.log[-1][0] == "PUT" and S.NODE_LOCK_TIME in api.nodes["node-a"]["metadata"]["annotations"]
        try:
            k.get_node("ghost")
            assert False
        except K.ApiError as e:
            assert e.code == 404
    finally:
        api.stop()


def test_scheduler_over_rest_with_the_pod_informer_and_plugin_pod_source():
    api = _cluster()
    try:
        k = K.RestKube(api.url)
        sch = S.Scheduler(k)
        stop, t = K.start_pod_informer(k, sch)
        sch.register_from_node_annotations_once()
        assert api.nodes["node-a"]["metadata"]["annotations"][P.HANDSHAKE].startswith("Requesting_")
        res = sch.filter({"Pod": api.pods[("default", "train")], "NodeNames": ["node-a"]})
        assert res["NodeNames"] == ["node-a"]
        annos = api.pods[("default", "train")]["metadata"]["annotations"]
        assert annos[P.TO_ALLOCATE] == "GPU-1,NVIDIA,8192,30:;"
        assert sch.bind({"PodName": "train", "PodNamespace": "default", "PodUID": "uid-train", "Node": "node-a"}) == {"Error": ""}
        # the device plugin on node-a finds the pod through the same apiserver
        src = K.KubePodSource(k)
        pod = src.get_pending_pod("node-a")
        assert pod.UID == "uid-train" and pod.Containers[0].Name == "main" and pod.Annotations[P.BIND_PHASE] == "allocating"
        src.patch_pod_annotations(pod, {P.BIND_PHASE: "success"})
        src.release_node_lock("node-a")
        assert S.NODE_LOCK_TIME not in api.nodes["node-a"]["metadata"]["annotations"]
        try:
            src.get_pending_pod("node-a")
            assert False
        except LookupError:
            pass
        # informer: a pod scheduled by ANOTHER extender replica shows up through the watch, a deleted one disappears
        other = _pod("other", {P.RESOURCE_NAME: "1"})
        other["metadata"]["annotations"] = {S.ASSIGNED_NODE: "node-a", P.ALLOCATED: "GPU-0,NVIDIA,100,10:;"}
        deadline = time.time() + 3
        while not api.watchers and time.time() < deadline:
            time.sleep(0.02)
        api.emit("ADDED", other)
        deadline = time.time() + 3
        while "uid-other" not in sch.pods and time.time() < deadline:
            time.sleep(0.02)
        assert sch.pods["uid-other"].NodeID == "node-a"
        api.emit("DELETED", other)
        deadline = time.time() + 3
        while "uid-other" in sch.pods and time.time() < deadline:
            time.sleep(0.02)
        assert "uid-other" not in sch.pods
        stop.set()
    finally:
        api.stop()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_scheduler_cli_serves_filter_and_metrics():
    api = _cluster()
    port, mport = _free_port(), _free_port()
    env = dict(os.environ, PYTHONPATH=ROOT)
    # the argument list of charts/vgpu/templates/scheduler/deployment.yaml:55-71 (values.yaml:61-63 adds --debug -v=4), ports changed
    proc = subprocess.Popen([sys.executable, "-m", "k8s_device_plugin_b200.plugin", "scheduler", "--apiserver", api.url,
                             "--resource-name=nvidia.com/gpu", "--resource-mem=nvidia.com/gpumem", "--resource-cores=nvidia.com/gpucores",
                             "--resource-mem-percentage=nvidia.com/gpumem-percentage", "--resource-priority=nvidia.com/priority",
                             f"--http_bind=127.0.0.1:{port}", "--cert_file=", "--key_file=", "--scheduler-name=4pd-scheduler",
                             f"--metrics-bind-address=127.0.0.1:{mport}", "--default-mem=0", "--default-cores=0", "--debug", "-v=4"], env=env, cwd=ROOT,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        body = None
        deadline = time.time() + 30
        while time.time() < deadline:
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
                c.request("POST", "/filter", json.dumps({"Pod": api.pods[("default", "train")], "NodeNames": ["node-a"]}), {"Content-Type": "application/json"})
                body = json.loads(c.getresponse().read())
                if body.get("NodeNames"):
                    break
            except OSError:
                pass
            time.sleep(0.3)
        assert body and body["NodeNames"] == ["node-a"], (body, proc.poll())
        assert api.pods[("default", "train")]["metadata"]["annotations"][S.ASSIGNED_NODE] == "node-a"
        c = http.client.HTTPConnection("127.0.0.1", mport, timeout=5)
        c.request("GET", "/metrics")
        assert "GPUDeviceSharedNum" in c.getresponse().read().decode()
    finally:
        proc.terminate()
        try:
            proc.wait(5)
        except subprocess.TimeoutExpired:
            proc.kill()
        api.stop()


def test_device_plugin_cli_end_to_end_on_fake_nvml(tmp_path):
    """BASELINE.json configs[0] through the real entry point: `python -m k8s_device_plugin_b200.plugin device-plugin`
    enumerates two (fake) GPUs through NVML, registers with a kubelet stub, answers ListAndWatch with 2 x split ids,
    keeps the node annotation registered on the (scripted) apiserver, answers Allocate for a pod the scheduler bound,
    and reports a GPU unhealthy after a critical Xid event."""
    from conftest import FAKE
    from k8s_device_plugin_b200.plugin import api as A
    from k8s_device_plugin_b200.plugin.kubelet_stub import KubeletStub
    sock = tmp_path / "dp"
    sock.mkdir()
    uuid0 = "GPU-b2a39081-f6e7-d4c5-3a2b-18097e6f5c00"            # what the fake NVML reports for device 0
    pod = _pod("train", {P.RESOURCE_NAME: "1", P.RESOURCE_MEM: "8192", P.RESOURCE_CORES: "30"})
    pod["metadata"]["annotations"] = {S.ASSIGNED_NODE: "node-a", P.BIND_TIME: "1", P.BIND_PHASE: "allocating",
                                      P.TO_ALLOCATE: f"{uuid0},NVIDIA,8192,30:;", P.ALLOCATED: f"{uuid0},NVIDIA,8192,30:;"}
    apisrv = FakeApiServer([{"metadata": {"name": "node-a", "annotations": {}}}], [pod])
    stub = KubeletStub(str(sock)); stub.start()
    env = dict(os.environ, PYTHONPATH=ROOT, LD_LIBRARY_PATH=FAKE + ":" + os.environ.get("LD_LIBRARY_PATH", ""), FAKE_GPU_COUNT="2",
               FAKE_NVML_XID="1:79@2500", NodeName="node-a")
    hook = tmp_path / "hook"
    proc = subprocess.Popen([sys.executable, "-m", "k8s_device_plugin_b200.plugin", "device-plugin", "--apiserver", apisrv.url, "--socket-dir", str(sock),
                             "--node-name", "node-a", "--device-split-count", "3", "--hook-path", str(hook), "--node-config-file", str(tmp_path / "none.json"),
                             "--mig-strategy=none", "--disable-core-limit=false", "-v=false"],
                            env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        assert stub.registered.wait(30), proc.poll()
        assert stub.request.resource_name == "nvidia.com/gpu" and stub.request.endpoint == "nvidia-gpu.sock"
        first = stub.list_and_watch_once(1)[0]
        ids = [d.ID for d in first.devices]
        assert len(ids) == 6 and ids[0] == uuid0 + "-0" and all(d.health == A.HEALTHY for d in first.devices)
        deadline = time.time() + 10
        while P.REGISTER not in apisrv.nodes["node-a"]["metadata"]["annotations"] and time.time() < deadline:
            time.sleep(0.05)
        reg = apisrv.nodes["node-a"]["metadata"]["annotations"][P.REGISTER]
        assert reg.startswith(f"{uuid0},3,183359,100,NVIDIA-NVIDIA B200 (fake),0,true:")
        resp = stub.allocate([[ids[0]]])
        envs = dict(resp.container_responses[0].envs)
        assert envs["CUDA_DEVICE_MEMORY_LIMIT_0"] == "8192m" and envs["CUDA_DEVICE_SM_LIMIT"] == "30" and envs["NVIDIA_VISIBLE_DEVICES"] == uuid0
        assert {m.container_path for m in resp.container_responses[0].mounts} >= {"/etc/ld.so.preload", "/tmp/vgpulock"}
        annos = apisrv.pods[("default", "train")]["metadata"]["annotations"]
        assert annos[P.BIND_PHASE] == "success" and annos[P.TO_ALLOCATE] == ";"
        # the Xid event for GPU 1 arrives ~2.5 s after the health checker starts: its three shares go Unhealthy
        two = stub.list_and_watch_once(2, timeout=20)
        bad = [d.ID for d in two[-1].devices if d.health == A.UNHEALTHY]
        assert len(bad) == 3 and all(b.startswith("GPU-b3a29180") for b in bad)
    finally:
        proc.terminate()
        try:
            proc.wait(5)
        except subprocess.TimeoutExpired:
            proc.kill()
        stub.stop(); apisrv.stop()


def test_monitor_cli_exports_region_and_host_metrics(tmp_path):
    from conftest import FAKE, HOOK_SO, OREF
    uid, ctr = "uid-train", "main"
    cdir = tmp_path / "containers" / f"{uid}_{ctr}"
    cdir.mkdir(parents=True)
    trace = tmp_path / "t.txt"
    trace.write_text("A 0 %d\nS 6000\n" % (100 << 20))
    henv = dict(os.environ, LD_LIBRARY_PATH=FAKE + ":" + os.environ.get("LD_LIBRARY_PATH", ""), LD_PRELOAD=HOOK_SO, LIBCUDA_LOG_LEVEL="0",
                CUDA_DEVICE_MEMORY_LIMIT_0="1g", CUDA_DEVICE_MEMORY_SHARED_CACHE=str(cdir / "x.cache"), FAKE_GPU_CTX_MIB="16")
    app = subprocess.Popen([os.path.join(OREF, "trace_replay"), str(trace)], env=henv, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    pod = _pod("train", {P.RESOURCE_NAME: "1"})
    pod["spec"]["nodeName"] = "node-a"
    apisrv = FakeApiServer([], [pod])
    port = _free_port()
    # the chart starts `vGPUmonitor` without arguments and with HOOK_PATH=<gpuHookPath>/vgpu (daemonsetnvidia.yaml:85-97):
    # the containers directory is $HOOK_PATH/containers (pathmonitor.go:31-36)
    env = dict(os.environ, PYTHONPATH=ROOT, LD_LIBRARY_PATH=FAKE + ":" + os.environ.get("LD_LIBRARY_PATH", ""), HOOK_PATH=str(tmp_path))
    mon = subprocess.Popen([sys.executable, "-m", "k8s_device_plugin_b200.plugin", "monitor", "--apiserver", apisrv.url, "--port", str(port)],
                           env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        text, deadline = "", time.time() + 30
        while time.time() < deadline:
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
                c.request("GET", "/metrics")
                text = c.getresponse().read().decode()
                if "vGPU_device_memory_usage_in_bytes{" in text:
                    break
            except OSError:
                pass
            time.sleep(0.3)
        base = 'podnamespace="default",podname="train",ctrname="main",vdeviceid="0"'
        assert f"vGPU_device_memory_usage_in_bytes{{{base}" in text and "} %s" % float((100 << 20) + (16 << 20)) in text
        assert f"vGPU_device_memory_limit_in_bytes{{{base}" in text and "HostGPUMemoryUsage{" in text and "HostCoreUtilization{" in text
    finally:
        for p in (mon, app):
            p.terminate()
            try:
                p.wait(5)
            except subprocess.TimeoutExpired:
                p.kill()
        apisrv.stop()


def test_cli_flag_syntax_follows_the_go_binaries(monkeypatch, capsys):
    """Go's flag packages take -name and --name alike and booleans as --flag / --flag=false; the chart relies on both
    (`--disable-core-limit=false`, `-v=false`, `-v=4`, `--debug`). Unsupported strategies are refused loudly, and the
    monitor insists on HOOK_PATH like validation.go:13-20."""
    from k8s_device_plugin_b200.plugin import __main__ as M
    assert M.go_style_argv(["-resource-name=x", "--http_bind=a", "-v=4", "-v", "--debug"]) == ["--resource-name=x", "--http_bind=a", "-v=4", "-v", "--debug"]
    assert M._gobool("T") is True and M._gobool("0") is False
    with pytest.raises(Exception):
        M._gobool("yes")
    monkeypatch.delenv("HOOK_PATH", raising=False)
    assert M.main(["monitor", "--apiserver", "http://127.0.0.1:1"]) == 1
    assert "HOOK_PATH" in capsys.readouterr().err
    assert M.main(["device-plugin", "--apiserver", "http://127.0.0.1:1", "--mig-strategy=mixed"]) == 1
    assert "mig-strategy" in capsys.readouterr().err
    assert M.main(["device-plugin", "--apiserver", "http://127.0.0.1:1", "-device-id-strategy=index"]) == 1
    assert M.main(["device-plugin", "--apiserver", "http://127.0.0.1:1", "--version"]) == 0
