// common.cuh — shared helpers for libvstb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/vstb200.h"

namespace vst {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// returns non-zero (and records the message) if the last launch failed to enqueue
inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    count_launch();
    return 0;
}

#define VST_CUDA_OK(expr)                                                            \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            vst::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define VST_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            vst::set_error(__VA_ARGS__);  \
            return 2;                     \
        }                                 \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int num_sms();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: remember per device
// (not per process) that it has been set, so a process that uses a second GPU opts in there too.
struct PerDeviceOnce {
    std::atomic<bool> done[64];
    PerDeviceOnce() { for (auto& d : done) d.store(false, std::memory_order_relaxed); }
};
template <typename K>
inline cudaError_t ensure_dyn_smem(PerDeviceOnce& once, K kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && once.done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && tracked) once.done[dev].store(true, std::memory_order_release);
    return e;
}

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains, so that its
// prologue (barrier init, TMEM allocation, weight / bias staging — nothing that depends on the predecessor's output)
// overlaps the predecessor's tail.  Such a kernel MUST execute pdl_wait() before its first access to dependent memory.
// VST_PDL=0 (developer knob) launches plainly.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- optional per-launch device timing (cudaEvent pairs on the launch stream), aggregated by
// kernel class; enabled by bench.py to obtain the live per-kernel durations of the roofline.
bool prof_on();
void prof_begin(cudaStream_t st, const char* cls, double flops, double bytes);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfScope(cudaStream_t s, const char* cls, double flops, double bytes) : st(s), on(prof_on()) {
        if (on) prof_begin(st, cls, flops, bytes);
    }
    ~ProfScope() {
        if (on) prof_end(st);
    }
};

}  // namespace vst
