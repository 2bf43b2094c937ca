"""RestKube — the apiserver client behind the scheduler extender, the device plugin and the monitor.

Reference: pkg/util/client/client.go:30-47 and pkg/k8sutil/client.go:27-41 (in-cluster config first, then kubeconfig),
pkg/util/util.go:46-76,273-318 (GetNode, GetPendingPod, PatchNodeAnnotations / PatchPodAnnotations: strategic merge
patches of metadata.annotations), pkg/scheduler/scheduler.go:106-121 (pod informer: list, then watch, handlers on
add/update/delete), :340 (pods/binding sub-resource), pkg/util/nodelock/nodelock.go (full node Update).

Plain HTTPS with the service-account token (urllib + ssl, nothing to install); only the handful of verbs the reference
uses. Objects are the apiserver's JSON, as everywhere else in this package.
"""
import json
import os
import ssl
import threading
import time
import urllib.error
import urllib.parse
import urllib.request

from .scheduler import KubeClient
from .server import ASSIGNED_NODE, BIND_ALLOCATING, BIND_PHASE, BIND_TIME, Container, Pod, PodSource

SA_DIR = "/var/run/secrets/kubernetes.io/serviceaccount"


class ApiError(RuntimeError):
    def __init__(self, code, body):
        super().__init__(f"apiserver answered {code}: {body[:300]}")
        self.code = code


class RestKube(KubeClient):
    def __init__(self, base_url, token=None, ca_file=None, insecure=False, timeout=30):
        self.base = base_url.rstrip("/")
        self.token = token
        self.timeout = timeout
        self.ctx = None
        if self.base.startswith("https"):
            self.ctx = ssl.create_default_context(cafile=ca_file) if ca_file else ssl.create_default_context()
            if insecure:
                self.ctx.check_hostname = False
                self.ctx.verify_mode = ssl.CERT_NONE

    @classmethod
    def in_cluster(cls):
        """rest.InClusterConfig(): KUBERNETES_SERVICE_HOST/PORT + the mounted service-account token and CA."""
        host, port = os.environ.get("KUBERNETES_SERVICE_HOST"), os.environ.get("KUBERNETES_SERVICE_PORT")
        if not host or not port:
            raise RuntimeError("unable to load in-cluster configuration, KUBERNETES_SERVICE_HOST and KUBERNETES_SERVICE_PORT must be defined")
        token = open(os.path.join(SA_DIR, "token")).read().strip()
        if ":" in host:
            host = f"[{host}]"
        return cls(f"https://{host}:{port}", token=token, ca_file=os.path.join(SA_DIR, "ca.crt"))

    # ---- transport
    def _request(self, method, path, body=None, content_type="application/json", query=None, stream=False, timeout=None):
        url = self.base + path + ("?" + urllib.parse.urlencode(query) if query else "")
        data = json.dumps(body).encode() if body is not None else None
        req = urllib.request.Request(url, data=data, method=method)
        req.add_header("Accept", "application/json")
        if data is not None:
            req.add_header("Content-Type", content_type)
        if self.token:
            req.add_header("Authorization", "Bearer " + self.token)
        try:
            resp = urllib.request.urlopen(req, timeout=timeout or self.timeout, context=self.ctx)
        except urllib.error.HTTPError as e:
            raise ApiError(e.code, e.read().decode(errors="replace")) from None
        if stream:
            return resp
        with resp:
            raw = resp.read()
        return json.loads(raw) if raw else {}

    # ---- KubeClient
    def get_pod(self, namespace, name):
        return self._request("GET", f"/api/v1/namespaces/{namespace}/pods/{name}")

    def list_pods(self, field_selector=None):
        q = {"fieldSelector": field_selector} if field_selector else None
        return self._request("GET", "/api/v1/pods", query=q)

    def patch_pod_annotations(self, namespace, name, annotations):
        return self._request("PATCH", f"/api/v1/namespaces/{namespace}/pods/{name}", {"metadata": {"annotations": annotations}},
                             "application/strategic-merge-patch+json")

    def bind_pod(self, namespace, name, uid, node):
        body = {"apiVersion": "v1", "kind": "Binding", "metadata": {"name": name, "uid": uid}, "target": {"kind": "Node", "name": node}}
        return self._request("POST", f"/api/v1/namespaces/{namespace}/pods/{name}/binding", body)

    def list_nodes(self):
        return self._request("GET", "/api/v1/nodes").get("items") or []

    def get_node(self, name):
        return self._request("GET", f"/api/v1/nodes/{name}")

    def patch_node_annotations(self, name, annotations):
        return self._request("PATCH", f"/api/v1/nodes/{name}", {"metadata": {"annotations": annotations}}, "application/strategic-merge-patch+json")

    def update_node(self, node):
        return self._request("PUT", f"/api/v1/nodes/{node['metadata']['name']}", node)

    # ---- informer: list, then watch from the list's resourceVersion; relist when the watch ends or expires (410)
    def watch_pods(self, on_add, on_update, on_delete, stop, relist_backoff=1.0):
        known = {}
        while not stop.is_set():
            try:
                lst = self._request("GET", "/api/v1/pods")
                seen = set()
                for p in lst.get("items") or []:
                    uid = p["metadata"].get("uid")
                    seen.add(uid)
                    if uid in known:
                        on_update(known[uid], p)
                    else:
                        on_add(p)
                    known[uid] = p
                for uid in [u for u in known if u not in seen]:
                    on_delete(known.pop(uid))
                rv = (lst.get("metadata") or {}).get("resourceVersion", "")
                resp = self._request("GET", "/api/v1/pods", query={"watch": "true", "resourceVersion": rv, "allowWatchBookmarks": "true"},
                                     stream=True, timeout=3600)
                with resp:
                    for line in resp:
                        if stop.is_set():
                            return
                        line = line.strip()
                        if not line:
                            continue
                        ev = json.loads(line)
                        typ, obj = ev.get("type"), ev.get("object") or {}
                        if typ == "ERROR":
                            break                                  # e.g. 410 Gone: relist
                        uid = (obj.get("metadata") or {}).get("uid")
                        if typ == "ADDED":
                            on_add(obj); known[uid] = obj
                        elif typ == "MODIFIED":
                            on_update(known.get(uid, obj), obj); known[uid] = obj
                        elif typ == "DELETED":
                            known.pop(uid, None); on_delete(obj)
            except (ApiError, OSError, ValueError):
                pass
            stop.wait(relist_backoff)


class KubePodSource(PodSource):
    """The device plugin's view of the cluster (util.GetPendingPod util.go:51-76, PatchPodAnnotations, nodelock release)."""

    def __init__(self, kube):
        self.kube = kube

    def get_pending_pod(self, node):
        for p in self.kube.list_pods().get("items") or []:
            a = (p.get("metadata") or {}).get("annotations") or {}
            if BIND_TIME not in a or a.get(BIND_PHASE) != BIND_ALLOCATING or a.get(ASSIGNED_NODE) != node:
                continue
            ctrs = [Container(c.get("name", ""), {e.get("name"): e.get("value", "") for e in c.get("env") or []})
                    for c in (p.get("spec") or {}).get("containers") or []]
            pod = Pod(UID=p["metadata"].get("uid", ""), Name=p["metadata"].get("name", ""), Annotations=dict(a), Containers=ctrs)
            pod.Namespace = p["metadata"].get("namespace", "default")
            return pod
        raise LookupError(f"no binding pod found on node {node}")

    def patch_pod_annotations(self, pod, annos):
        self.kube.patch_pod_annotations(getattr(pod, "Namespace", "default"), pod.Name, annos)
        pod.Annotations.update(annos)

    def release_node_lock(self, node):
        from .scheduler import release_node_lock
        try:
            release_node_lock(self.kube, node)
        except Exception:
            pass


def start_pod_informer(kube, scheduler, stop=None):
    """Scheduler.Start (scheduler.go:106-121): feed the pod manager from the apiserver."""
    stop = stop or threading.Event()
    t = threading.Thread(target=kube.watch_pods, args=(scheduler.on_add_pod, scheduler.on_update_pod, scheduler.on_del_pod, stop), daemon=True)
    t.start()
    return stop, t


def run_registration_loop(scheduler, stop, interval=15.0):
    """RegisterFromNodeAnnotatons' outer loop (scheduler.go:123,239): one pass every 15 s."""
    while not stop.is_set():
        try:
            scheduler.register_from_node_annotations_once()
        except Exception:
            pass
        stop.wait(interval)


def wait_forever(stop):
    while not stop.is_set():
        time.sleep(0.5)
