"""Entry points of the host-side services, flag for flag after the reference binaries:

  python -m k8s_device_plugin_b200.plugin scheduler     (cmd/scheduler/main.go:48-63)
  python -m k8s_device_plugin_b200.plugin device-plugin (cmd/device-plugin/nvidia/main.go + vgpucfg.go:21-66)
  python -m k8s_device_plugin_b200.plugin monitor       (cmd/vGPUmonitor/main.go)
"""
import argparse
import os
import sys
import threading

from . import api, kube as K, monitor as M, rm, scheduler as S, server as P


def _kube(args):
    if args.apiserver:
        return K.RestKube(args.apiserver, token=args.token, insecure=args.insecure)
    return K.RestKube.in_cluster()


def _common(p):
    p.add_argument("--apiserver", default=os.environ.get("VGPU_APISERVER", ""), help="apiserver URL (default: in-cluster config)")
    p.add_argument("--token", default=os.environ.get("VGPU_APISERVER_TOKEN"))
    p.add_argument("--insecure", action="store_true")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="k8s_device_plugin_b200.plugin")
    sub = ap.add_subparsers(dest="cmd", required=True)

    sc = sub.add_parser("scheduler", help="kubernetes vgpu scheduler (extender + webhook)")
    _common(sc)
    sc.add_argument("--http_bind", default="127.0.0.1:8080")
    sc.add_argument("--cert_file", default="")
    sc.add_argument("--key_file", default="")
    sc.add_argument("--scheduler-name", default="")
    sc.add_argument("--default-mem", type=int, default=0)
    sc.add_argument("--default-cores", type=int, default=0)
    sc.add_argument("--metrics-bind-address", default=":9395")
    sc.add_argument("--resource-name", default=P.RESOURCE_NAME)
    sc.add_argument("--resource-mem", default=P.RESOURCE_MEM)
    sc.add_argument("--resource-mem-percentage", default=P.RESOURCE_MEM_PERCENTAGE)
    sc.add_argument("--resource-cores", default=P.RESOURCE_CORES)
    sc.add_argument("--resource-priority", default="vgputaskpriority")
    sc.add_argument("--multi-container", action="store_true", help="per-container device bookkeeping (vgpu_sched.h mode 1) instead of the reference's")

    dp = sub.add_parser("device-plugin", help="NVIDIA device plugin for Kubernetes (vGPU path)")
    _common(dp)
    dp.add_argument("--node-name", default=os.environ.get("NodeName", ""))
    dp.add_argument("--device-split-count", type=int, default=int(os.environ.get("DEVICE_SPLIT_COUNT", "2")))
    dp.add_argument("--device-memory-scaling", type=float, default=float(os.environ.get("DEVICE_MEMORY_SCALING", "1.0")))
    dp.add_argument("--device-cores-scaling", type=float, default=float(os.environ.get("DEVICE_CORES_SCALING", "1.0")))
    dp.add_argument("--disable-core-limit", action="store_true", default=os.environ.get("DISABLE_CORE_LIMIT", "") in ("1", "true"))
    dp.add_argument("--resource-name", default=P.RESOURCE_NAME)
    dp.add_argument("--config-file", default="/config/config.json")
    dp.add_argument("--socket-dir", default=api.DEVICE_PLUGIN_PATH)
    dp.add_argument("--hook-path", default=os.environ.get("HOOK_PATH", "/usr/local"))

    mo = sub.add_parser("monitor", help="vGPU monitor: region metrics + priority feedback")
    _common(mo)
    mo.add_argument("--containers-path", default=os.path.join(os.environ.get("HOOK_PATH", "/usr/local"), "vgpu", "containers"))
    mo.add_argument("--port", type=int, default=9394)
    mo.add_argument("--node-name", default=os.environ.get("NODE_NAME", ""))

    args = ap.parse_args(argv)
    stop = threading.Event()

    if args.cmd == "scheduler":
        cfg = S.Config(HttpBind=args.http_bind, SchedulerName=args.scheduler_name, DefaultMem=args.default_mem, DefaultCores=args.default_cores,
                       MetricsBindAddress=args.metrics_bind_address, ResourceName=args.resource_name, ResourceMem=args.resource_mem,
                       ResourceMemPercentage=args.resource_mem_percentage, ResourceCores=args.resource_cores,
                       ResourcePriority=args.resource_priority, MultiContainer=args.multi_container)
        kube = _kube(args)
        sch = S.Scheduler(kube, cfg)
        K.start_pod_informer(kube, sch, stop)
        threading.Thread(target=K.run_registration_loop, args=(sch, stop), daemon=True).start()
        S.serve(sch, args.metrics_bind_address if args.metrics_bind_address.rpartition(":")[0] else "0.0.0.0" + args.metrics_bind_address)
        S.serve(sch, args.http_bind, args.cert_file, args.key_file)
        K.wait_forever(stop)
        return 0

    if args.cmd == "device-plugin":
        kube = _kube(args)
        split, mscale, cscale = rm.read_node_config(args.config_file, args.node_name, args.device_split_count, args.device_memory_scaling,
                                                    args.device_cores_scaling)

        def make():
            return [P.NvidiaDevicePlugin(rm.nvml_devices(), K.KubePodSource(kube), node_name=args.node_name, resource_name=args.resource_name,
                                         socket_dir=args.socket_dir, host_hook_path=args.hook_path, device_split_count=split,
                                         device_memory_scaling=mscale, device_cores_scaling=cscale, disable_core_limit=args.disable_core_limit)]

        mgr = rm.PluginManager(make, os.path.join(args.socket_dir, "kubelet.sock"))
        mgr.install_signal_handlers()

        def side_loops():
            while not mgr.plugins and not stop.is_set():
                stop.wait(0.2)
            p = mgr.plugins[0]
            threading.Thread(target=rm.watch_and_register, args=(p, kube, stop), daemon=True).start()
            rm.check_health(stop, p.devices, p.mark_unhealthy, rm.NvmlEventSource())

        threading.Thread(target=side_loops, daemon=True).start()
        mgr.run()
        stop.set()
        return 0

    if args.cmd == "monitor":
        kube = _kube(args)

        def list_pods():
            sel = f"spec.nodeName={args.node_name}" if args.node_name else None
            return [M.PodInfo(p["metadata"].get("uid", ""), p["metadata"].get("namespace", ""), p["metadata"].get("name", ""),
                              [c.get("name", "") for c in (p.get("spec") or {}).get("containers") or []])
                    for p in kube.list_pods(sel).get("items") or []]

        mon = M.Monitor(args.containers_path, list_pods, host_gpus=M.nvml_host_gpus)
        mon.serve(args.port, host="0.0.0.0")
        while not stop.is_set():          # watchAndFeedback (feedback.go:257-270): every 5 s
            try:
                mon.observe()
            except Exception as e:
                print("observe failed:", e, file=sys.stderr)
            stop.wait(5)
        return 0


if __name__ == "__main__":
    sys.exit(main())
