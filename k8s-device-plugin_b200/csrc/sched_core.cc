// Scoring core of the scheduler extender (include/vgpu_sched.h). Behaviour follows the reference's Go code statement by
// statement (pkg/scheduler/score.go, pkg/device/nvidia/device.go:69-118); nothing here touches CUDA.
#include "vgpu_sched.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define VGPU_API extern "C" __attribute__((visibility("default")))

namespace {

const char kNvidia[] = "NVIDIA";   // nvidia.NvidiaGPUDevice (device.go:18)

std::string upper(const char *s) {
    std::string r(s ? s : "");
    for (auto &c : r) c = (char)std::toupper((unsigned char)c);
    return r;
}

bool contains(const std::string &hay, const std::string &needle) { return hay.find(needle) != std::string::npos; }

std::vector<std::string> split(const std::string &s, char sep) {   // strings.Split: n separators -> n+1 fields
    std::vector<std::string> out;
    size_t from = 0;
    for (;;) {
        size_t at = s.find(sep, from);
        if (at == std::string::npos) { out.push_back(s.substr(from)); break; }
        out.push_back(s.substr(from, at - from));
        from = at + 1;
    }
    return out;
}

// checkGPUtype (device.go:69-104): use-list wins over nouse-list; matching is case-insensitive substring, and an empty
// list element matches every card (strings.Contains(x, "") is true) — kept.
bool check_gpu_type(const vgpu_sched_annotations_t *a, const char *cardtype) {
    std::string card = upper(cardtype);
    if (a && a->use_gputype) {
        for (const auto &v : split(a->use_gputype, ','))
            if (contains(card, upper(v.c_str()))) return true;
        return false;
    }
    if (a && a->nouse_gputype) {
        for (const auto &v : split(a->nouse_gputype, ','))
            if (contains(card, upper(v.c_str()))) return false;
        return true;
    }
    return true;
}

// assertNuma (device.go:106-115): strconv.ParseBool accepts exactly these spellings
bool assert_numa(const vgpu_sched_annotations_t *a) {
    if (!a || !a->numa_bind) return false;
    static const char *truthy[] = {"1", "t", "T", "TRUE", "true", "True"};
    for (const char *t : truthy)
        if (std::strcmp(a->numa_bind, t) == 0) return true;
    return false;
}

void copy_str(char *dst, const char *src) {
    std::snprintf(dst, VGPU_PLUGIN_MAX_STR, "%s", src);
}

}  // namespace

VGPU_API void vgpu_sched_sort_devices(vgpu_device_usage_t *devs, int n) {
    std::stable_sort(devs, devs + n, [](const vgpu_device_usage_t &a, const vgpu_device_usage_t &b) {
        if (a.numa == b.numa) return a.count - a.used < b.count - b.used;
        return a.numa < b.numa;
    });
}

VGPU_API int vgpu_sched_check_type(const vgpu_sched_annotations_t *annos, const vgpu_device_usage_t *d, const vgpu_device_request_t *req,
                                   int *pass, int *numa_assert) {
    if (pass) *pass = 0;
    if (numa_assert) *numa_assert = 0;
    // general type check (score.go:74): the device type must contain the request type
    if (!contains(d->type, req->type)) return 0;
    if (std::strcmp(req->type, kNvidia) != 0) return 0;   // the only vendor on this path
    if (pass) *pass = check_gpu_type(annos, d->type);
    if (numa_assert) *numa_assert = assert_numa(annos);
    return 1;
}

VGPU_API int vgpu_sched_fit_in_certain_device(const vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *req,
                                              const vgpu_sched_annotations_t *annos, vgpu_sched_assignment_t *out, int cap, int *n_out) {
    int32_t nums = req->nums;
    const int32_t origin = req->nums;
    int prevnuma = -1;
    int cnt = 0;
    auto done = [&](int rc) { if (n_out) *n_out = cnt; return rc; };
    for (int i = n - 1; i >= 0; i--) {
        const vgpu_device_usage_t &d = devs[i];
        int pass = 0, numa = 0;
        vgpu_sched_check_type(annos, &d, req, &pass, &numa);
        if (!pass) continue;                                   // "card type mismatch"
        if (numa && prevnuma != d.numa) {                      // NUMA binding: start over on every NUMA boundary
            nums = origin;
            prevnuma = d.numa;
            cnt = 0;
        }
        if (d.count <= d.used) continue;
        if (req->coresreq > 100) return done(0);               // "core limit can't exceed 100"
        int32_t memreq = 0;
        if (req->memreq > 0) memreq = req->memreq;
        if (req->mem_percentage_req != 101 && req->memreq == 0)
            memreq = (int32_t)((uint32_t)d.totalmem * (uint32_t)req->mem_percentage_req) / 100;   // int32 arithmetic like Go
        if (d.totalmem - d.usedmem < memreq) continue;
        if (d.totalcore - d.usedcores < req->coresreq) continue;
        if (d.totalcore == 100 && req->coresreq == 100 && d.used > 0) continue;            // exclusive card wanted
        if (d.totalcore != 0 && d.usedcores == d.totalcore && req->coresreq == 0) continue; // core=0 job on a full GPU
        if (nums > 0) {
            nums--;
            if (cnt < cap) {
                vgpu_sched_assignment_t &a = out[cnt];
                std::memset(&a, 0, sizeof a);
                a.idx = i;
                copy_str(a.dev.uuid, d.id);
                copy_str(a.dev.type, req->type);
                a.dev.usedmem = memreq;
                a.dev.usedcores = req->coresreq;
            }
            cnt++;
        }
        if (nums == 0) return done(1);
    }
    return done(0);
}

VGPU_API int vgpu_sched_fit_in_devices(vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *reqs, int n_req,
                                       const vgpu_sched_annotations_t *annos, vgpu_sched_assignment_t *out, int cap, int *n_out, float *score) {
    int32_t total = 0, free_ = 0;
    int sums = 0, cnt = 0;
    if (n_out) *n_out = 0;
    if (score) *score = 0;
    for (int r = 0; r < n_req; r++) {
        const vgpu_device_request_t &k = reqs[r];
        sums += k.nums;
        if (k.nums > n) return 0;                              // more devices than the node has
        vgpu_sched_sort_devices(devs, n);
        int got = 0;
        if (!vgpu_sched_fit_in_certain_device(devs, n, &k, annos, out + cnt, cap - cnt, &got)) return 0;
        if (cnt + got > cap) return 0;
        for (int j = 0; j < got; j++) {
            vgpu_device_usage_t &d = devs[out[cnt + j].idx];
            total += d.count;
            free_ += d.count - d.used;
            d.used++;
            d.usedcores += out[cnt + j].dev.usedcores;
            d.usedmem += out[cnt + j].dev.usedmem;
        }
        cnt += got;
    }
    if (n_out) *n_out = cnt;
    if (score) *score = (float)total / (float)free_ + (float)(n - sums);
    return 1;
}

VGPU_API int vgpu_sched_score_node(vgpu_device_usage_t *devs, int n, const vgpu_device_request_t *reqs, int n_ctrs,
                                   const vgpu_sched_annotations_t *annos, int mode, vgpu_sched_assignment_t *out, int cap, int *n_out,
                                   float *score) {
    int cnt = 0;
    float total_score = 0;
    int vendor_lists = 0;       // len(score.devices): number of vendor keys (0 or 1 here)
    int list_len = 0;           // len(score.devices["NVIDIA"]): container slots appended so far
    bool all_fit = true;
    if (n_out) *n_out = 0;
    if (score) *score = 0;
    for (int c = 0; c < n_ctrs; c++) {
        if (reqs[c].nums == 0) {
            if (mode == 0) {
                // score.devices[idx][ctrid] = append(score.devices[idx][ctrid], ...) — indexes slot ctrid of every vendor list
                if (vendor_lists > 0 && c >= list_len) return -5;   // Go: index out of range -> panic
            } else {
                list_len++;
            }
            continue;
        }
        int got = 0;
        float s = 0;
        if (!vgpu_sched_fit_in_devices(devs, n, &reqs[c], 1, annos, out + cnt, cap - cnt, &got, &s)) { all_fit = false; break; }
        for (int j = 0; j < got; j++) out[cnt + j].container = c;
        cnt += got;
        total_score += s;
        vendor_lists = 1;
        list_len++;
    }
    if (n_out) *n_out = cnt;
    if (score) *score = total_score;
    if (mode == 0) return vendor_lists == n_ctrs ? 1 : 0;     // len(score.devices) == len(nums)
    return all_fit && vendor_lists > 0 ? 1 : 0;
}

VGPU_API void vgpu_sched_charge(vgpu_device_usage_t *devs, int n, const vgpu_container_device_t *pod_devs, int n_pod_devs) {
    for (int j = 0; j < n_pod_devs; j++)
        for (int i = 0; i < n; i++)
            if (std::strcmp(devs[i].id, pod_devs[j].uuid) == 0) {
                devs[i].used++;
                devs[i].usedmem += pod_devs[j].usedmem;
                devs[i].usedcores += pod_devs[j].usedcores;
            }
}
