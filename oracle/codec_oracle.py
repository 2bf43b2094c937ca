"""TEST INFRASTRUCTURE (oracle/). Pure-Python restatement of the reference's annotation codec, statement by statement
after pkg/util/util.go:78-271 — the checker for csrc/plugin_core.cc in tests/test_plugin.py (randomised differential
test). Go semantics that matter are spelled out: strings.Split keeps empty fields, strconv.Atoi / ParseInt(…, 32) give 0
on error (the reference ignores the error), ParseBool accepts exactly 1,t,T,TRUE,true,True / 0,f,F,FALSE,false,False."""


class CodecError(ValueError):
    pass


def _atoi(s, bits=64):
    import re
    if not re.fullmatch(r"[+-]?[0-9]+", s):
        return 0
    v = int(s)
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    if v < lo or v > hi:
        return hi if v > 0 else lo        # ParseInt returns the clamped limit next to ErrRange; the reference ignores the error
    return v


def _i32(x):
    return ((x + 2 ** 31) % 2 ** 32) - 2 ** 31


def _parse_bool(s):
    return s in ("1", "t", "T", "TRUE", "true", "True")


def decode_node_devices(text):                           # util.go:78-109
    if ":" not in text:
        raise CodecError("node annotations not decode successfully")
    out = []
    for val in text.split(":"):
        if "," in val:
            items = val.split(",")
            if len(items) != 7:
                raise CodecError("node annotations not decode successfully")
            numa = _i32(_atoi(items[5]))          # Go keeps a 64-bit int; the C ABI carries int32
            out.append(dict(Id=items[0], Count=_i32(_atoi(items[1])), Devmem=_i32(_atoi(items[2])), Devcore=_i32(_atoi(items[3])), Type=items[4],
                            Numa=numa, Health=_parse_bool(items[6])))
    return out


def encode_node_devices(devs):                           # util.go:111-118
    return "".join(f"{d['Id']},{d['Count']},{d['Devmem']},{d['Devcore']},{d['Type']},{d['Numa']},{'true' if d['Health'] else 'false'}:" for d in devs)


def decode_container_devices(text):                      # util.go:162-191
    if len(text) == 0:
        return []
    out = []
    for val in text.split(":"):
        if "," in val:
            f = val.split(",")
            if len(f) < 4:
                raise CodecError("pod annotation format error")
            out.append(dict(UUID=f[0], Type=f[1], Usedmem=_i32(_atoi(f[2], 32)), Usedcores=_i32(_atoi(f[3], 32))))
    return out


def encode_container_devices(cd):                        # util.go:120-128
    return "".join(f"{d['UUID']},{d['Type']},{d['Usedmem']},{d['Usedcores']}:" for d in cd)


def encode_pod_single_device(pd):                        # util.go:142-150: ONE ';' for the whole pod
    return "".join(encode_container_devices(c) for c in pd) + ";"


def decode_pod_single_device(text):                      # the per-vendor body of DecodePodDevices, util.go:203-210
    return [decode_container_devices(s) for s in text.split(";")]


def next_device_request(annotation):                     # GetNextDeviceRequest, util.go:216-236
    for idx, ctr in enumerate(decode_pod_single_device(annotation)):
        if len(ctr) > 0:
            return idx, ctr
    raise LookupError("device request not found")


def erase_next_device_request(annotation):               # EraseNextDeviceTypeFromAnnotation, util.go:244-271
    res, found = [], False
    for val in decode_pod_single_device(annotation):
        if found:
            res.append(val)
        elif len(val) > 0:
            found = True
            res.append([])
        else:
            res.append(val)
    return encode_pod_single_device(res)
