// lib.cu — library-level state: last-error string, launch counter, device properties.
#include <string.h>
#include <vector>
#include <mutex>
#include <utility>
#include <stdlib.h>
#include "common.cuh"

namespace vst {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VST_PDL"); v = e ? atoi(e) : 1; }
    return v != 0;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ------------------------------------------------------------------------------------------
// per-launch profiler
// ------------------------------------------------------------------------------------------
struct ProfRec {
    char cls[40];
    cudaEvent_t a, b;
    double flops, bytes;
};
static bool g_prof = false;
static std::vector<ProfRec> g_recs;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_pool;
static std::mutex g_prof_mu;

bool prof_on() { return g_prof; }

void prof_begin(cudaStream_t st, const char* cls, double flops, double bytes) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    snprintf(r.cls, sizeof(r.cls), "%s", cls);
    if (!g_pool.empty()) {
        r.a = g_pool.back().first; r.b = g_pool.back().second; g_pool.pop_back();
    } else {
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    }
    r.flops = flops; r.bytes = bytes;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}
void prof_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}

}  // namespace vst

extern "C" int vst_profile_enable(int on) {
    vst::g_prof = on != 0;
    return 0;
}
extern "C" int vst_profile_collect(vst_profile_entry* out, int max_entries, int* n_out) {
    using namespace vst;
    VST_REQUIRE(out && n_out && max_entries > 0, "vst_profile_collect: bad arguments");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int n = 0;
    for (ProfRec& r : g_recs) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.a, r.b);
        g_pool.push_back({r.a, r.b});
        if (e != cudaSuccess) continue;
        int k = 0;
        for (; k < n; ++k) if (strncmp(out[k].name, r.cls, sizeof(out[k].name)) == 0) break;
        if (k == n) {
            if (n == max_entries) continue;
            memset(&out[n], 0, sizeof(out[n]));
            snprintf(out[n].name, sizeof(out[n].name), "%s", r.cls);
            ++n;
        }
        out[k].ms += ms; out[k].launches += 1; out[k].flops += r.flops; out[k].bytes += r.bytes;
    }
    g_recs.clear();
    *n_out = n;
    return 0;
}

extern "C" const char* vst_last_error(void) { return vst::g_err; }
extern "C" int vst_version(void) { return 100; }
extern "C" unsigned long long vst_launch_count(void) { return vst::g_launches.load(); }
