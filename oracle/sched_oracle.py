"""TEST INFRASTRUCTURE (oracle/). Pure-Python restatement of the reference scheduler's scoring, statement by statement
after pkg/scheduler/score.go:36-226 and pkg/device/nvidia/device.go:69-118 — the checker for csrc/sched_core.cc in
tests/test_scheduler.py (randomised differential test). Plain dicts/lists, no native code, nothing shared with the
product. Only tests may import this.

Device: dict(Id, Index, Used, Count, Usedmem, Totalmem, Totalcore, Usedcores, Numa, Type, Health)
Request: dict(Nums, Type, Memreq, MemPercentagereq, Coresreq)
"""
GPU_IN_USE = "nvidia.com/use-gputype"
GPU_NO_USE = "nvidia.com/nouse-gputype"
NUMA_BIND = "nvidia.com/numa-bind"


def _i32(x):
    return ((x + 2 ** 31) % 2 ** 32) - 2 ** 31


def _f32(x):
    import struct
    return struct.unpack("<f", struct.pack("<f", x))[0]


def sort_devices(devs):                      # score.go:36-49; Go's sort is an insertion sort (stable) up to 12 elements
    devs.sort(key=lambda d: (d["Numa"], d["Count"] - d["Used"]))


def check_gpu_type(annos, cardtype):         # device.go:69-104
    if GPU_IN_USE in annos:
        return any(v.upper() in cardtype.upper() for v in annos[GPU_IN_USE].split(","))
    if GPU_NO_USE in annos:
        return not any(v.upper() in cardtype.upper() for v in annos[GPU_NO_USE].split(","))
    return True


def assert_numa(annos):                      # device.go:106-115 (strconv.ParseBool)
    return annos.get(NUMA_BIND) in ("1", "t", "T", "TRUE", "true", "True")


def check_type(annos, d, n):                 # score.go:72-85 + device.go:117-122
    if n["Type"] not in d["Type"]:
        return False, False
    if n["Type"] == "NVIDIA":
        return check_gpu_type(annos, d["Type"]), assert_numa(annos)
    return False, False


def fit_in_certain_device(devs, request, annos):      # score.go:87-161
    k = dict(request)
    origin = k["Nums"]
    prevnuma = -1
    tmp = []
    for i in range(len(devs) - 1, -1, -1):
        d = devs[i]
        found, numa = check_type(annos, d, k)
        if not found:
            continue
        if numa and prevnuma != d["Numa"]:
            k["Nums"] = origin
            prevnuma = d["Numa"]
            tmp = []
        memreq = 0
        if d["Count"] <= d["Used"]:
            continue
        if k["Coresreq"] > 100:
            return False, tmp
        if k["Memreq"] > 0:
            memreq = k["Memreq"]
        if k["MemPercentagereq"] != 101 and k["Memreq"] == 0:
            prod = _i32(d["Totalmem"] * k["MemPercentagereq"])
            memreq = int(prod / 100)         # Go integer division truncates toward zero
        if d["Totalmem"] - d["Usedmem"] < memreq:
            continue
        if d["Totalcore"] - d["Usedcores"] < k["Coresreq"]:
            continue
        if d["Totalcore"] == 100 and k["Coresreq"] == 100 and d["Used"] > 0:
            continue
        if d["Totalcore"] != 0 and d["Usedcores"] == d["Totalcore"] and k["Coresreq"] == 0:
            continue
        if k["Nums"] > 0:
            k["Nums"] -= 1
            tmp.append({"Idx": i, "UUID": d["Id"], "Type": k["Type"], "Usedmem": memreq, "Usedcores": k["Coresreq"]})
        if k["Nums"] == 0:
            return True, tmp
    return False, tmp


def fit_in_devices(devs, requests, annos):            # score.go:163-195 (one vendor per container here)
    total = free = sums = 0
    out = []
    for k in requests:
        sums += k["Nums"]
        if k["Nums"] > len(devs):
            return False, 0.0, []
        sort_devices(devs)
        fit, tmp = fit_in_certain_device(devs, k, annos)
        if not fit:
            return False, 0.0, []
        for val in tmp:
            d = devs[val["Idx"]]
            total += d["Count"]
            free += d["Count"] - d["Used"]
            d["Used"] += 1
            d["Usedcores"] += val["Usedcores"]
            d["Usedmem"] += val["Usedmem"]
        out += tmp
    return True, _f32(_f32(_f32(total) / _f32(free)) + _f32(len(devs) - sums)), out


def score_node(devs, ctr_requests, annos):            # calcScore body, score.go:199-224, reference bookkeeping
    """Returns ('fit', score, [[devices of container] ...]) | ('nofit',) | ('panic',)."""
    score = 0.0
    lists = None                                       # score.devices["NVIDIA"]; None = key absent
    for ctrid, req in enumerate(ctr_requests):
        if req is None or req["Nums"] == 0:
            if lists is not None:
                if ctrid >= len(lists):
                    return ("panic",)
                lists[ctrid].append({})
            continue
        fit, s, devices = fit_in_devices(devs, [req], annos)
        if not fit:
            break
        score = _f32(score + s)
        lists = (lists or []) + [devices]
    if (0 if lists is None else 1) == len(ctr_requests):
        return ("fit", score, lists)
    return ("nofit",)
