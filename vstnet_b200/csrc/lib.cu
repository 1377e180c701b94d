// lib.cu — library-level state: last-error string, launch counter, device properties.
#include <string.h>
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

}  // namespace vst

extern "C" const char* vst_last_error(void) { return vst::g_err; }
extern "C" int vst_version(void) { return 100; }
extern "C" unsigned long long vst_launch_count(void) { return vst::g_launches.load(); }
