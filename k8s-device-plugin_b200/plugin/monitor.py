"""Node monitor for the vGPU path: the consumer of the shared region on the host side (reference: cmd/vGPUmonitor —
pathmonitor.go scanning, metrics.go Prometheus exposition, feedback.go priority arbitration). The region is read and
written through the C ABI (include/vgpu.h: vgpu_region_*, vgpu_monitor_observe) instead of a hand-copied struct;
this module is the thin shell around it: directory scanning, label plumbing and an HTTP endpoint."""
import os
import shutil
import threading
import time
from http.server import BaseHTTPRequestHandler, HTTPServer

import k8s_device_plugin_b200 as v


def is_valid_pod(name, pod_uids):
    """isVaildPod (pathmonitor.go:73-80): the directory name contains the UID of a live pod."""
    return any(uid in name for uid in pod_uids)


def check_files(dirpath):
    """checkfiles (pathmonitor.go:38-71): at most two entries; the first *.cache that is not the hook library."""
    files = sorted(os.listdir(dirpath))
    if len(files) > 2:
        raise ValueError("cache num not matched")
    for name in files:
        if "libvgpu.so" in name or ".cache" not in name:
            continue
        try:
            return v.Region(os.path.join(dirpath, name))
        except OSError:
            continue
    return None


def nvml_host_gpus():
    """HostGPUMemoryUsage / HostCoreUtilization inputs through NVML (metrics.go:152-195)."""
    import pynvml as nv
    nv.nvmlInit()
    try:
        out = []
        for i in range(nv.nvmlDeviceGetCount()):
            h = nv.nvmlDeviceGetHandleByIndex(i)
            uuid = nv.nvmlDeviceGetUUID(h)
            out.append((i, uuid.decode() if isinstance(uuid, bytes) else uuid, nv.nvmlDeviceGetMemoryInfo(h).used, nv.nvmlDeviceGetUtilizationRates(h).gpu))
        return out
    finally:
        nv.nvmlShutdown()


class PodInfo:
    def __init__(self, uid, namespace, name, containers):
        self.uid, self.namespace, self.name, self.containers = uid, namespace, name, list(containers)


class Monitor:
    """containers_path = <HOOK_PATH>/containers (pathmonitor.go:31-36). list_pods() returns PodInfo objects (client-go in
    the reference)."""

    GC_SECONDS = 300                                   # pathmonitor.go:96

    def __init__(self, containers_path, list_pods, host_gpus=None):
        self.path = containers_path
        self.list_pods = list_pods
        self.host_gpus = host_gpus                     # () -> [(index, uuid, used_bytes, sm_util_percent)]; nvml_host_gpus on a GPU node
        self.regions = {}                              # dir -> (idstr, Region)
        self.lock = threading.Lock()

    def monitor_path(self, now=None):
        """monitorpath (pathmonitor.go:82-128): adopt new container dirs, drop + delete dirs of vanished pods after 300 s."""
        now = time.time() if now is None else now
        uids = [p.uid for p in self.list_pods()]
        with self.lock:
            for name in sorted(os.listdir(self.path)):
                d = os.path.join(self.path, name)
                if not os.path.isdir(d):
                    continue
                if not is_valid_pod(name, uids):
                    if os.path.getmtime(d) + self.GC_SECONDS < now:
                        ent = self.regions.pop(d, None)
                        if ent:
                            ent[1].close()
                        shutil.rmtree(d, ignore_errors=True)
                    continue
                if d not in self.regions:
                    try:
                        r = check_files(d)
                    except ValueError:
                        continue
                    if r is not None:
                        self.regions[d] = (name, r)

    def observe(self):
        """watchAndFeedback body (feedback.go:257-270): rescan, then one Observe pass."""
        self.monitor_path()
        with self.lock:
            return v.monitor_observe([r for _, r in self.regions.values()])

    def collect(self):
        """Prometheus text exposition with the reference's metric names and label sets (metrics.go:66-97, 227-252)."""
        self.monitor_path()
        pods = {p.uid: p for p in self.list_pods()}
        out = ["# HELP vGPU_device_memory_usage_in_bytes vGPU device usage", "# TYPE vGPU_device_memory_usage_in_bytes gauge",
               "# HELP vGPU_device_memory_limit_in_bytes vGPU device limit", "# TYPE vGPU_device_memory_limit_in_bytes gauge",
               "# HELP Device_memory_desc_of_container Container device meory description", "# TYPE Device_memory_desc_of_container counter"]
        for idx, uuid, used, util in (self.host_gpus() if self.host_gpus else []):     # metrics.go:152-195 (NVML walk)
            out.append(f'HostGPUMemoryUsage{{deviceidx="{idx}",deviceuuid="{uuid}"}} {float(used)}')
            out.append(f'HostCoreUtilization{{deviceidx="{idx}",deviceuuid="{uuid}"}} {float(util)}')
        with self.lock:
            for idstr, region in self.regions.values():
                parts = idstr.split("_")                 # parseidstr (metrics.go:108-115)
                if len(parts) < 2 or parts[0] not in pods:
                    continue
                pod, ctr = pods[parts[0]], parts[1]
                if ctr not in pod.containers:
                    continue
                snap = region.snapshot()
                procs = [region.proc(i) for i in range(snap.proc_num)]
                for i in range(int(snap.device_num)):
                    tot = {k: sum(getattr(p.used[i], k) for p in procs) for k in ("context_size", "module_size", "buffer_size", "offset", "total")}
                    uuid = bytes(snap.uuids[i]).split(b"\0")[0].decode()[:40]
                    base = f'podnamespace="{pod.namespace}",podname="{pod.name}",ctrname="{ctr}",vdeviceid="{i}",deviceuuid="{uuid}"'
                    out.append(f"vGPU_device_memory_usage_in_bytes{{{base}}} {float(tot['total'])}")
                    out.append(f"vGPU_device_memory_limit_in_bytes{{{base}}} {float(snap.limit[i])}")
                    out.append(f'Device_memory_desc_of_container{{{base},context="{tot["context_size"]}",module="{tot["module_size"]}",'
                               f'data="{tot["buffer_size"]}",offset="{tot["offset"]}"}} {float(tot["total"])}')
                    sw = region.swap_counters(i)         # extension: swap engine counters
                    if sw and sw["processes"]:
                        for key, name in (("page_out_bytes", "vGPU_swap_page_out_bytes_total"), ("page_in_bytes", "vGPU_swap_page_in_bytes_total"),
                                          ("faults", "vGPU_swap_faults_total"), ("evictions", "vGPU_swap_evictions_total"),
                                          ("resident_bytes", "vGPU_swap_resident_bytes"), ("live_bytes", "vGPU_swap_live_bytes"),
                                          ("host_bytes", "vGPU_swap_host_pool_bytes")):
                            out.append(f"{name}{{{base}}} {float(sw[key])}")
        return "\n".join(out) + "\n"

    def serve(self, port=9394, host="127.0.0.1"):       # metrics.go:309
        mon = self

        class H(BaseHTTPRequestHandler):
            def do_GET(self):
                body = mon.collect().encode()
                self.send_response(200)
                self.send_header("Content-Type", "text/plain; version=0.0.4")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def log_message(self, *a):
                pass

        srv = HTTPServer((host, port), H)
        threading.Thread(target=srv.serve_forever, daemon=True).start()
        return srv
