// plugin_core.cc — host-side device-plugin logic for the vGPU path behind include/vgpu_plugin.h.
// Mirrors, function for function: pkg/util/util.go (annotation codec), rm/devices.go:144-167 (vGPU fan-out),
// plugin/register.go:128-153 (advertised memory/cores) and plugin/server.go:288-411 (Allocate's env/mount contract).
#include "vgpu_plugin.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define VGPU_API extern "C" __attribute__((visibility("default")))

namespace {

std::vector<std::string> split(const std::string &s, char sep) {   // strings.Split semantics (keeps empty fields)
    std::vector<std::string> out;
    size_t start = 0;
    for (;;) {
        size_t p = s.find(sep, start);
        if (p == std::string::npos) { out.push_back(s.substr(start)); break; }
        out.push_back(s.substr(start, p - start));
        start = p + 1;
    }
    return out;
}
int put(const std::string &s, char *buf, size_t cap) {
    if (s.size() + 1 > cap) return -2;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return 0;
}
void copy_str(char *dst, const std::string &s) {
    std::snprintf(dst, VGPU_PLUGIN_MAX_STR, "%s", s.c_str());
}
// Go's strconv with the error IGNORED, as the reference uses it (`v, _ := strconv.Atoi(x)`): text that is not
// [+-]?[0-9]+ (no white space, no prefixes) gives 0; a value outside the target width gives the CLAMPED limit of that
// width (ParseInt returns it next to ErrRange). Atoi targets int (64-bit) and the reference then truncates with int32(…);
// ParseInt(…, 10, 32) clamps to int32 directly.
bool go_decimal(const std::string &s, bool *neg, std::string *digits) {
    size_t i = 0;
    *neg = false;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) { *neg = s[i] == '-'; i++; }
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++) if (s[k] < '0' || s[k] > '9') return false;
    *digits = s.substr(i);
    return true;
}
int64_t go_parse_int(const std::string &s, int bits) {
    bool neg; std::string dg;
    if (!go_decimal(s, &neg, &dg)) return 0;
    const uint64_t maxpos = bits == 32 ? (uint64_t)INT32_MAX : (uint64_t)INT64_MAX;
    uint64_t acc = 0;
    bool over = false;
    for (char c : dg) {
        uint64_t d = (uint64_t)(c - '0');
        if (acc > (UINT64_MAX - d) / 10) { over = true; break; }
        acc = acc * 10 + d;
    }
    if (over || acc > maxpos + (neg ? 1 : 0)) return neg ? -(int64_t)maxpos - 1 : (int64_t)maxpos;
    return neg ? (int64_t)(0 - acc) : (int64_t)acc;
}
int32_t atoi_to_i32(const std::string &s) { return (int32_t)(uint32_t)(uint64_t)go_parse_int(s, 64); }   // int32(strconv.Atoi(s))
int32_t atoi32(const std::string &s) { return (int32_t)go_parse_int(s, 32); }                              // ParseInt(s, 10, 32)
std::string enc_container(const vgpu_container_device_t *d, int n) {   // EncodeContainerDevices util.go:120-128
    std::string t;
    for (int i = 0; i < n; i++)
        t += std::string(d[i].uuid) + "," + d[i].type + "," + std::to_string(d[i].usedmem) + "," + std::to_string(d[i].usedcores) + ":";
    return t;
}
int dec_container(const std::string &s, std::vector<vgpu_container_device_t> *out) {   // DecodeContainerDevices util.go:162-191
    if (s.empty()) return 0;
    for (const std::string &val : split(s, ':')) {
        if (val.find(',') == std::string::npos) continue;
        std::vector<std::string> f = split(val, ',');
        if (f.size() < 4) return -1;   // "pod annotation format error; information missing"
        vgpu_container_device_t d;
        std::memset(&d, 0, sizeof d);
        copy_str(d.uuid, f[0]);
        copy_str(d.type, f[1]);
        d.usedmem = atoi32(f[2]);
        d.usedcores = atoi32(f[3]);
        out->push_back(d);
    }
    return 0;
}
int dec_pod(const std::string &s, std::vector<std::vector<vgpu_container_device_t>> *out) {   // DecodePodDevices util.go:204-210
    for (const std::string &c : split(s, ';')) {
        std::vector<vgpu_container_device_t> cd;
        if (dec_container(c, &cd) != 0) return -1;
        out->push_back(cd);
    }
    return 0;
}
std::string enc_pod(const std::vector<std::vector<vgpu_container_device_t>> &pd) {   // EncodePodSingleDevice util.go:142-150
    std::string res;
    for (const auto &c : pd) res += enc_container(c.data(), (int)c.size());
    res += ";";   // ONE ';' after all containers — the reference's behaviour at this commit (Appendix E)
    return res;
}

}  // namespace

VGPU_API int vgpu_codec_encode_node_devices(const vgpu_node_device_t *d, int n, char *buf, size_t cap) {
    std::string t;
    for (int i = 0; i < n; i++)
        t += std::string(d[i].id) + "," + std::to_string(d[i].count) + "," + std::to_string(d[i].devmem) + "," + std::to_string(d[i].devcore) +
             "," + d[i].type + "," + std::to_string(d[i].numa) + "," + (d[i].health ? "true" : "false") + ":";
    return put(t, buf, cap);
}

VGPU_API int vgpu_codec_decode_node_devices(const char *s, vgpu_node_device_t *out, int cap, int *n) {
    if (!s || !n) return -1;
    std::string str(s);
    *n = 0;
    if (str.find(':') == std::string::npos) return -1;   // "node annotations not decode successfully"
    for (const std::string &val : split(str, ':')) {
        if (val.find(',') == std::string::npos) continue;
        std::vector<std::string> f = split(val, ',');
        if (f.size() != 7) return -1;
        if (*n >= cap) return -2;
        vgpu_node_device_t &d = out[(*n)++];
        std::memset(&d, 0, sizeof d);
        copy_str(d.id, f[0]);
        d.count = atoi_to_i32(f[1]); d.devmem = atoi_to_i32(f[2]); d.devcore = atoi_to_i32(f[3]);
        copy_str(d.type, f[4]);
        d.numa = atoi_to_i32(f[5]);   // Go keeps a 64-bit int here; NUMA node numbers fit 32 bits
        // strconv.ParseBool accepts 1,t,T,TRUE,true,True (error ignored -> false)
        d.health = (f[6] == "1" || f[6] == "t" || f[6] == "T" || f[6] == "TRUE" || f[6] == "true" || f[6] == "True");
    }
    return 0;
}

VGPU_API int vgpu_codec_encode_container_devices(const vgpu_container_device_t *d, int n, char *buf, size_t cap) {
    return put(enc_container(d, n), buf, cap);
}
VGPU_API int vgpu_codec_decode_container_devices(const char *s, vgpu_container_device_t *out, int cap, int *n) {
    if (!s || !n) return -1;
    std::vector<vgpu_container_device_t> v;
    if (dec_container(s, &v) != 0) return -1;
    if ((int)v.size() > cap) return -2;
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    *n = (int)v.size();
    return 0;
}
VGPU_API int vgpu_codec_encode_pod_single_device(const vgpu_container_device_t *devs, const int *counts, int n_ctrs, char *buf, size_t cap) {
    std::vector<std::vector<vgpu_container_device_t>> pd;
    int k = 0;
    for (int c = 0; c < n_ctrs; c++) { pd.emplace_back(devs + k, devs + k + counts[c]); k += counts[c]; }
    return put(enc_pod(pd), buf, cap);
}
VGPU_API int vgpu_codec_decode_pod_single_device(const char *s, vgpu_container_device_t *out, int dev_cap, int *counts, int ctr_cap, int *n_ctrs) {
    if (!s || !n_ctrs) return -1;
    std::vector<std::vector<vgpu_container_device_t>> pd;
    if (dec_pod(s, &pd) != 0) return -1;
    if ((int)pd.size() > ctr_cap) return -2;
    int k = 0;
    for (size_t c = 0; c < pd.size(); c++) {
        counts[c] = (int)pd[c].size();
        for (auto &d : pd[c]) { if (k >= dev_cap) return -2; out[k++] = d; }
    }
    *n_ctrs = (int)pd.size();
    return 0;
}
VGPU_API int vgpu_codec_next_device_request(const char *anno, int *ctr_index, vgpu_container_device_t *out, int cap, int *n) {
    if (!anno || !n) return -1;
    std::vector<std::vector<vgpu_container_device_t>> pd;
    if (dec_pod(anno, &pd) != 0) return -1;
    for (size_t c = 0; c < pd.size(); c++) {
        if (pd[c].empty()) continue;
        if ((int)pd[c].size() > cap) return -2;
        for (size_t i = 0; i < pd[c].size(); i++) out[i] = pd[c][i];
        *n = (int)pd[c].size();
        if (ctr_index) *ctr_index = (int)c;
        return 0;
    }
    return -3;   // "device request not found"
}
VGPU_API int vgpu_codec_erase_next_device_request(const char *anno, char *buf, size_t cap) {
    if (!anno) return -1;
    std::vector<std::vector<vgpu_container_device_t>> pd, res;
    if (dec_pod(anno, &pd) != 0) return -1;
    bool found = false;
    for (auto &c : pd) {
        if (!found && !c.empty()) { found = true; res.emplace_back(); }
        else res.push_back(c);
    }
    return put(enc_pod(res), buf, cap);
}

VGPU_API int vgpu_plugin_device_id(const char *uuid, unsigned index, char *buf, size_t cap) {
    if (!uuid) return -1;
    return put(std::string(uuid) + "-" + std::to_string(index), buf, cap);
}
VGPU_API int32_t vgpu_plugin_registered_mem(uint64_t total_bytes, double scaling) {
    int32_t mib = (int32_t)(total_bytes / 1024 / 1024);
    if (scaling != 1) mib = (int32_t)((double)mib * scaling);
    return mib;
}
VGPU_API int32_t vgpu_plugin_registered_cores(double scaling) { return (int32_t)(scaling * 100); }

static void add_env(vgpu_allocate_out_t *o, const std::string &k, const std::string &v) {
    if (o->n_envs >= 32) return;
    std::snprintf(o->envs[o->n_envs].key, sizeof o->envs[0].key, "%s", k.c_str());
    std::snprintf(o->envs[o->n_envs].value, sizeof o->envs[0].value, "%s", v.c_str());
    o->n_envs++;
}
static void add_mount(vgpu_allocate_out_t *o, const std::string &c, const std::string &h, bool ro) {
    if (o->n_mounts >= 8) return;
    std::snprintf(o->mounts[o->n_mounts].container_path, 512, "%s", c.c_str());
    std::snprintf(o->mounts[o->n_mounts].host_path, 512, "%s", h.c_str());
    o->mounts[o->n_mounts].read_only = ro;
    o->n_mounts++;
}

VGPU_API int vgpu_plugin_allocate(const vgpu_allocate_in_t *in, vgpu_allocate_out_t *out) {
    if (!in || !out || !in->host_hook_path || !in->pod_uid || !in->container_name) return -1;
    std::memset(out, 0, sizeof *out);
    if (in->n_devices != in->n_requested_ids) return -4;   // "device allocate number not matched" (server.go:328-331)
    const std::string hook = in->host_hook_path;
    // getAllocateResponse -> apiEnvs(deviceListEnvvar, deviceIDs): NVIDIA_VISIBLE_DEVICES=<uuid,uuid,...>
    std::string ids;
    for (int i = 0; i < in->n_devices; i++) { if (i) ids += ","; ids += in->devices[i].uuid; }
    add_env(out, in->device_list_envvar ? in->device_list_envvar : "NVIDIA_VISIBLE_DEVICES", ids);
    for (int i = 0; i < in->n_devices; i++)                                                  // server.go:343-345
        add_env(out, "CUDA_DEVICE_MEMORY_LIMIT_" + std::to_string(i), std::to_string(in->devices[i].usedmem) + "m");
    add_env(out, "CUDA_DEVICE_SM_LIMIT", std::to_string(in->n_devices ? in->devices[0].usedcores : 0));   // :354
    std::string cache = in->cache_uuid ? in->cache_uuid : (std::string(in->pod_uid) + "-" + in->container_name);
    add_env(out, "CUDA_DEVICE_MEMORY_SHARED_CACHE", hook + "/vgpu/" + cache + ".cache");       // :355
    if (in->device_memory_scaling > 1) add_env(out, "CUDA_OVERSUBSCRIBE", "true");            // :356-358
    if (in->disable_core_limit) add_env(out, "GPU_CORE_UTILIZATION_POLICY", "disable");       // :359-361 (api.CoreLimitSwitch)
    std::string dir = hook + "/vgpu/containers/" + in->pod_uid + "_" + in->container_name;    // :362
    std::snprintf(out->cache_host_dir, sizeof out->cache_host_dir, "%s", dir.c_str());
    add_mount(out, hook + "/vgpu/libvgpu.so", hook + "/vgpu/libvgpu.so", true);               // :368-379
    add_mount(out, hook + "/vgpu", dir, false);
    add_mount(out, "/tmp/vgpulock", "/tmp/vgpulock", false);
    if (!in->container_sets_disable_control)                                                   // :380-391
        add_mount(out, "/etc/ld.so.preload", hook + "/vgpu/ld.so.preload", true);
    if (in->license_present) {                                                                 // :392-404
        add_mount(out, "/vgpu/", hook + "/vgpu/license", true);
        add_mount(out, "/usr/bin/vgpuvalidator", hook + "/vgpu/vgpuvalidator", true);
    }
    return 0;
}
