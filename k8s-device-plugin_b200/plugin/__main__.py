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


def _gobool(v):
    """Go's strconv.ParseBool, as the flag packages apply it to --flag=value."""
    if v in ("1", "t", "T", "TRUE", "true", "True"):
        return True
    if v in ("0", "f", "F", "FALSE", "false", "False"):
        return False
    raise argparse.ArgumentTypeError(f"invalid boolean value {v!r}")


def _bool_flag(p, *names, default=False, env=None, help=None):
    """A Go-style boolean flag: --flag, --flag=true, --flag=false; optional environment fallback (urfave/cli EnvVars)."""
    if env and os.environ.get(env, "") != "":
        try:
            default = _gobool(os.environ[env])
        except argparse.ArgumentTypeError:
            pass
    p.add_argument(*names, type=_gobool, nargs="?", const=True, default=default, help=help)


def go_style_argv(argv):
    """Go's flag package (and cobra / urfave/cli on top of it) accepts -name and --name alike; argparse wants --name for
    multi-letter flags. `-v=4` / `-v=false` stay as they are (single letter)."""
    out = []
    for a in argv:
        if a.startswith("-") and not a.startswith("--") and len(a.split("=", 1)[0]) > 2:
            a = "-" + a
        out.append(a)
    return out


def _klog_flags(p):
    """klog.InitFlags (pkg/util/util.go:321-327): accepted so that the chart's `-v=4` works; only -v is looked at."""
    p.add_argument("-v", "--v", default="0", help="log level verbosity")
    for name in ("logtostderr", "alsologtostderr", "add_dir_header", "skip_headers", "skip_log_headers", "one_output"):
        _bool_flag(p, "--" + name, default=(name == "logtostderr"))
    for name in ("log_dir", "log_file", "log_file_max_size", "stderrthreshold", "vmodule", "log_backtrace_at"):
        p.add_argument("--" + name, default="")


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
    _bool_flag(sc, "--debug", help="debug mode (pkg/device/devices.go:98)")
    _klog_flags(sc)
    sc.add_argument("--multi-container", action="store_true", help="per-container device bookkeeping (vgpu_sched.h mode 1) instead of the reference's")

    dp = sub.add_parser("device-plugin", help="NVIDIA device plugin for Kubernetes (vGPU path)")
    _common(dp)
    dp.add_argument("--node-name", default=os.environ.get("NodeName", ""))
    dp.add_argument("--device-split-count", type=int, default=int(os.environ.get("DEVICE_SPLIT_COUNT", "2")))
    dp.add_argument("--device-memory-scaling", type=float, default=float(os.environ.get("DEVICE_MEMORY_SCALING", "1.0")))
    dp.add_argument("--device-cores-scaling", type=float, default=float(os.environ.get("DEVICE_CORES_SCALING", "1.0")))
    _bool_flag(dp, "--disable-core-limit", env="DISABLE_CORE_LIMIT", help="If set, the core utilization limit will be ignored (vgpucfg.go:42)")
    # flags of the upstream NVIDIA plugin the reference keeps (cmd/device-plugin/nvidia/main.go:49-119); the vGPU path
    # needs none of them, but the chart passes --mig-strategy, so they parse — and what is not supported says so
    dp.add_argument("--mig-strategy", default=os.environ.get("MIG_STRATEGY", "none"))
    _bool_flag(dp, "--fail-on-init-error", default=True, env="FAIL_ON_INIT_ERROR")
    dp.add_argument("--nvidia-driver-root", default=os.environ.get("NVIDIA_DRIVER_ROOT", "/"))
    _bool_flag(dp, "--pass-device-specs", env="PASS_DEVICE_SPECS")
    dp.add_argument("--device-list-strategy", default=os.environ.get("DEVICE_LIST_STRATEGY", "envvar"))
    dp.add_argument("--device-id-strategy", default=os.environ.get("DEVICE_ID_STRATEGY", "uuid"))
    _bool_flag(dp, "--gds-enabled", env="GDS_ENABLED")
    _bool_flag(dp, "--mofed-enabled", env="MOFED_ENABLED")
    dp.add_argument("--cdi-annotation-prefix", default=os.environ.get("CDI_ANNOTATION_PREFIX", "cdi.k8s.io/"))
    dp.add_argument("--nvidia-ctk-path", default=os.environ.get("NVIDIA_CTK_PATH", "/usr/bin/nvidia-ctk"))
    dp.add_argument("--container-driver-root", default=os.environ.get("CONTAINER_DRIVER_ROOT", "/driver-root"))
    _bool_flag(dp, "-v", "--version", help="urfave/cli's built-in; the chart passes -v=false")
    dp.add_argument("--resource-name", default=P.RESOURCE_NAME)
    dp.add_argument("--config-file", default=os.environ.get("CONFIG_FILE", ""), help="upstream plugin config file (unused on the vGPU path)")
    dp.add_argument("--node-config-file", default="/config/config.json", help="per-node overrides (vgpucfg.go:80: fixed path in the reference)")
    dp.add_argument("--socket-dir", default=api.DEVICE_PLUGIN_PATH)
    dp.add_argument("--hook-path", default=os.environ.get("HOOK_PATH", "/usr/local"))

    mo = sub.add_parser("monitor", help="vGPU monitor: region metrics + priority feedback")
    _common(mo)
    # pathmonitor.go:31-36: containerPath = $HOOK_PATH/containers, and HOOK_PATH is required (validation.go:8-20); the chart
    # sets HOOK_PATH=<gpuHookPath>/vgpu for this container (daemonsetnvidia.yaml:96-97)
    mo.add_argument("--containers-path", default=os.path.join(os.environ["HOOK_PATH"], "containers") if "HOOK_PATH" in os.environ else None)
    mo.add_argument("--port", type=int, default=9394)
    mo.add_argument("--node-name", default=os.environ.get("NODE_NAME", ""))

    args = ap.parse_args(go_style_argv(sys.argv[1:] if argv is None else list(argv)))
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
        if args.version is True:
            print("k8s_device_plugin_b200 device-plugin")
            return 0
        if args.mig_strategy != "none":
            print(f"--mig-strategy={args.mig_strategy}: MIG devices are outside this plugin's path (DESIGN.md §8); only \"none\" is supported", file=sys.stderr)
            return 1
        if args.device_list_strategy != "envvar" or args.device_id_strategy != "uuid":
            print("only --device-list-strategy=envvar and --device-id-strategy=uuid are supported (the vGPU path of server.go uses no other)", file=sys.stderr)
            return 1
        split, mscale, cscale = rm.read_node_config(args.node_config_file, args.node_name, args.device_split_count, args.device_memory_scaling,
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
        if not args.containers_path:
            print("required environment variable HOOK_PATH not set", file=sys.stderr)     # ValidateEnvVars (validation.go:13-20)
            return 1
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
