// cwct.cu — Cholesky whitening / colouring transform on the device, with no host round trips.
//
// Replaces models/cWCT.py: whitening :134-149, coloring :152-164, cholesky_dec :111-132,
// _transfer :24-47, _transfer_seg :49-109 (+ compute_label_info :166-189, get_index :199-204),
// interpolation :206-262.   Closed form, per label l:
//     out = T_l (x - mu_c,l) + beta_l
//     T_l    = (1-a_c) (sum_k a_k Ls_k) Lc^-1 + a_c I
//     beta_l = (1-a_c) sum_k a_k mu_s,k + a_c mu_c
//
// Three kernels: (1) stats  — per-label count / sum / Gram in one pass over [C, n] features,
// fp32 products, fp64 cross-tile accumulation, pivot-shifted to avoid cancellation;
// (2) factor — one CTA per label: covariance, Cholesky with the cumulative eps*I retry,
// triangular solve, all in fp64 shared memory; (3) apply — label-indexed C x C transform.
#include <stdlib.h>
#include "kernels.cuh"

namespace vst {

// ------------------------------------------------------------------------------------------
// stats block layout (doubles unless noted):  count[L] | sum[L*C] | gram[L*C*C] | pivot[C] (float)
// ------------------------------------------------------------------------------------------
struct StatsView {
    double* count;
    double* sum;
    double* gram;
    float* pivot;
};
__host__ __device__ inline size_t stats_doubles(int C, int L) { return (size_t)L * (1 + C + (size_t)C * C); }
__host__ __device__ inline StatsView stats_view(void* p, int C, int L) {
    StatsView v;
    v.count = (double*)p;
    v.sum = v.count + L;
    v.gram = v.sum + (size_t)L * C;
    v.pivot = (float*)(v.gram + (size_t)L * C * C);
    return v;
}

// pivot[c] = mean of up to 4096 evenly spaced samples of channel c
__global__ void pivot_kernel(const float* __restrict__ feat, long long n, float* __restrict__ pivot) {
    const int c = blockIdx.x;
    const long long S = n < 4096 ? n : 4096;
    float s = 0.f;
    for (long long k = threadIdx.x; k < S; k += blockDim.x) s += __ldg(feat + (size_t)c * n + (k * n) / S);
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) pivot[c] = s / (float)S;
    }
}

// the same for the latent that the P4 half-state x1 spreads to: channel c of sub-position 0 = state channel c of x1
__global__ void pivot_state_kernel(const float* __restrict__ x1, int h, int w, float* __restrict__ pivot) {
    const int c = blockIdx.x;
    const long long n = (long long)h * w, S = n < 4096 ? n : 4096;
    const float* plane = x1 + (size_t)(c >> 2) * p4_plane_px(h, w) * 4 + (c & 3);
    float s = 0.f;
    for (long long k = threadIdx.x; k < S; k += blockDim.x) {
        const long long p = (k * n) / S;
        const int y = (int)(p / w), x = (int)(p - (long long)y * w);
        s += __ldg(plane + ((size_t)(y + 1) * (w + 2) + x + 1) * 4);
    }
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) pivot[c] = s / (float)S;
    }
}

// ---- C <= 32: every warp owns a private 32x32 Gram in registers (lane: 4 rows x 8 cols) and
// streams its own contiguous pixel range through a private smem sub-tile.  fp32 products are summed
// in fp32 over runs of at most 256 pixels and folded into fp64; the unmasked kernel keeps the fp64
// partials in registers, reduces the 8 warps of the CTA through shared memory and issues ONE set of
// global fp64 atomics per CTA (the per-warp flushes of an earlier version serialised on ~1000 hot
// addresses and cost 4x the HBM time of the pass).
template <bool MASKED>
__global__ void __launch_bounds__(256) gram32_kernel(const float* __restrict__ feat, int C, long long n,
                                                     const uint8_t* __restrict__ labels, int L, StatsView sv,
                                                     long long chunk) {
    __shared__ __align__(16) float tile[8][8][33][4];
    __shared__ float piv[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) piv[threadIdx.x] = threadIdx.x < C ? sv.pivot[threadIdx.x] : 0.f;
    __syncthreads();
    const long long start = ((long long)blockIdx.x * 8 + warp) * chunk;
    const long long end = start + chunk < n ? start + chunk : n;
    const int gi = lane >> 2, gj = (lane & 3) * 2;   // row group, first col group

    float acc[4][8];
    float sacc = 0.f;
    double dacc[MASKED ? 1 : 4][MASKED ? 1 : 8];     // unmasked: fp64 partial Gram of this warp
    double dsum = 0.0, dcnt = 0.0;
    int cnt = 0, cur = MASKED ? -1 : 0, since = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            acc[a][b] = 0.f;
            if (!MASKED) dacc[a][b] = 0.0;
        }

    auto flush = [&](int l) {
        if (MASKED) {
            if (l >= 0 && l < L && cnt > 0) {
                double* g = sv.gram + (size_t)l * C * C;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        int i = gi * 4 + a, j = gj * 4 + b;
                        if (i < C && j < C) atomicAdd(g + (size_t)i * C + j, (double)acc[a][b]);
                    }
                if (lane < C) atomicAdd(sv.sum + (size_t)l * C + lane, (double)sacc);
                if (lane == 0) atomicAdd(sv.count + l, (double)cnt);
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) dacc[a][b] += (double)acc[a][b];
            dsum += (double)sacc;
            dcnt += (double)cnt;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
        sacc = 0.f; cnt = 0; since = 0;
    };

    for (long long base = start; base < end; base += 32) {
        const long long p = base + lane;
        const bool ok = p < end;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float4 v;
            v.x = (ok && 4 * g + 0 < C) ? __ldg(feat + (size_t)(4 * g + 0) * n + p) - piv[4 * g + 0] : 0.f;
            v.y = (ok && 4 * g + 1 < C) ? __ldg(feat + (size_t)(4 * g + 1) * n + p) - piv[4 * g + 1] : 0.f;
            v.z = (ok && 4 * g + 2 < C) ? __ldg(feat + (size_t)(4 * g + 2) * n + p) - piv[4 * g + 2] : 0.f;
            v.w = (ok && 4 * g + 3 < C) ? __ldg(feat + (size_t)(4 * g + 3) * n + p) - piv[4 * g + 3] : 0.f;
            *reinterpret_cast<float4*>(tile[warp][g][lane]) = v;
        }
        int lab = 0;
        if (MASKED) lab = ok ? (int)labels[p] : -1;
        __syncwarp();
        const int npx = (int)(end - base < 32 ? end - base : 32);
        for (int q = 0; q < npx; ++q) {
            if (MASKED) {
                int l = __shfl_sync(0xffffffffu, lab, q);
                if (l != cur) { flush(cur); cur = l; }
            }
            const float4 a4 = *reinterpret_cast<const float4*>(tile[warp][gi][q]);
            const float4 b0 = *reinterpret_cast<const float4*>(tile[warp][gj][q]);
            const float4 b1 = *reinterpret_cast<const float4*>(tile[warp][gj + 1][q]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            sacc += tile[warp][lane >> 2][q][lane & 3];
            ++cnt;
        }
        __syncwarp();
        since += 32;
        if (since >= (MASKED ? 1024 : 256)) flush(cur);   // bound the fp32 run length; the cross-run sum is fp64
    }
    flush(cur);

    if (!MASKED) {
        // CTA-level fp64 reduction through the (now idle) tile storage, then one set of global atomics
        __syncthreads();
        double* red = reinterpret_cast<double*>(&tile[0][0][0][0]);          // 8*8*33*4 floats = 33 KB >= (1024 + 33) doubles
        for (int w = 0; w < 8; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const int idx = (gi * 4 + a) * 32 + gj * 4 + b;
                        red[idx] = (w == 0 ? 0.0 : red[idx]) + dacc[a][b];
                    }
                red[1024 + lane] = (w == 0 ? 0.0 : red[1024 + lane]) + dsum;
                if (lane == 0) red[1056] = (w == 0 ? 0.0 : red[1056]) + dcnt;
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < 1024; e += 256) {
            const int i = e >> 5, j = e & 31;
            if (i < C && j < C) atomicAdd(sv.gram + (size_t)i * C + j, red[e]);
        }
        if (threadIdx.x < C) atomicAdd(sv.sum + threadIdx.x, red[1024 + threadIdx.x]);
        if (threadIdx.x == 0) atomicAdd(sv.count, red[1056]);
    }
}

// ---- 32 < C <= 128: the CTA owns one 128x128 Gram (thread: 8 rows x 8 cols) and streams a
// contiguous pixel range through a shared [group][pixel][4] tile.
template <bool MASKED>
__global__ void __launch_bounds__(256) gram128_kernel(const float* __restrict__ feat, int C, long long n,
                                                      const uint8_t* __restrict__ labels, int L, StatsView sv,
                                                      long long chunk) {
    __shared__ __align__(16) float tile[32][33][4];
    __shared__ float piv[128];
    __shared__ int labs[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 128) piv[tid] = tid < C ? sv.pivot[tid] : 0.f;
    __syncthreads();
    const long long start = (long long)blockIdx.x * chunk;
    const long long end = start + chunk < n ? start + chunk : n;
    const int ti = tid >> 4, tj = tid & 15;   // rows 8ti..8ti+7 ; cols 4tj..4tj+3 and 64+4tj..64+4tj+3

    float acc[8][8];
    float sacc = 0.f;
    int cnt = 0, cur = MASKED ? -1 : 0, since = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    auto flush = [&](int l) {
        if (l >= 0 && l < L && cnt > 0) {
            double* g = sv.gram + (size_t)l * C * C;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    int i = ti * 8 + a, j = (b < 4) ? tj * 4 + b : 64 + tj * 4 + (b - 4);
                    if (i < C && j < C) atomicAdd(g + (size_t)i * C + j, (double)acc[a][b]);
                }
            if (tid < C) atomicAdd(sv.sum + (size_t)l * C + tid, (double)sacc);
            if (tid == 0) atomicAdd(sv.count + l, (double)cnt);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
        sacc = 0.f; cnt = 0; since = 0;
    };

    for (long long base = start; base < end; base += 32) {
        const long long p = base + lane;
        const bool ok = p < end;
        __syncthreads();   // previous tile fully consumed
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
            const int g = warp * 4 + gg;
            float4 v;
            v.x = (ok && 4 * g + 0 < C) ? __ldg(feat + (size_t)(4 * g + 0) * n + p) - piv[4 * g + 0] : 0.f;
            v.y = (ok && 4 * g + 1 < C) ? __ldg(feat + (size_t)(4 * g + 1) * n + p) - piv[4 * g + 1] : 0.f;
            v.z = (ok && 4 * g + 2 < C) ? __ldg(feat + (size_t)(4 * g + 2) * n + p) - piv[4 * g + 2] : 0.f;
            v.w = (ok && 4 * g + 3 < C) ? __ldg(feat + (size_t)(4 * g + 3) * n + p) - piv[4 * g + 3] : 0.f;
            *reinterpret_cast<float4*>(tile[g][lane]) = v;
        }
        if (MASKED && warp == 0) labs[lane] = ok ? (int)labels[p] : -1;
        __syncthreads();
        const int npx = (int)(end - base < 32 ? end - base : 32);
        for (int q = 0; q < npx; ++q) {
            if (MASKED) {
                int l = labs[q];
                if (l != cur) { flush(cur); cur = l; }
            }
            const float4 a0 = *reinterpret_cast<const float4*>(tile[2 * ti][q]);
            const float4 a1 = *reinterpret_cast<const float4*>(tile[2 * ti + 1][q]);
            const float4 b0 = *reinterpret_cast<const float4*>(tile[tj][q]);
            const float4 b1 = *reinterpret_cast<const float4*>(tile[tj + 16][q]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
            if (tid < 128) sacc += tile[tid >> 2][q][tid & 3];
            ++cnt;
        }
        since += 32;
        if (since >= 1024) flush(cur);
    }
    flush(cur);
}

// ------------------------------------------------------------------------------------------
// factor: one CTA per label
// ------------------------------------------------------------------------------------------
struct FactorArgs {
    const void* cstats;
    const void* sstats[VST_MAX_STYLES];
    float alpha_s[VST_MAX_STYLES];
    int n_styles;
    float alpha_c, eps;
    int C, L, masked, use_double;
    int mode;                    // 0: T = Mix Lc^-1 (transfer)   1: T = Lc^-1, beta = 0 (whitening, cWCT.py:134-149)
                                 // 2: T = Mix, mu = 0, beta = mixed style mean (coloring, cWCT.py:152-164)
    float *T, *mu, *beta;
    int *valid, *status;
};

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

// covariance (+ jitter) of one label from a stats block into packed-lower A; returns nothing
__device__ void build_cov(double* A, const StatsView& sv, int l, int C, double n, double jitter, bool /*round32*/) {
    const double* G = sv.gram + (size_t)l * C * C;
    const double* s = sv.sum + (size_t)l * C;
    for (int e = threadIdx.x; e < C * (C + 1) / 2; e += blockDim.x) {
        int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
        while (tri(i + 1, 0) <= e) ++i;
        while (tri(i, 0) > e) --i;
        int j = e - tri(i, 0);
        // average the two triangles: the Gram is accumulated unsymmetrised
        double g = 0.5 * (G[(size_t)i * C + j] + G[(size_t)j * C + i]);
        // The covariance stays in fp64 (the Gram was accumulated in fp64): rounding it to fp32 "as the reference's matmul
        // result would be" only adds the reference's own rounding noise a second time — on an ill-conditioned label the
        // result moved 2x further from the fp64 evaluation than the reference's.  `round32` (fp32 reference arithmetic)
        // now only selects the failure rule of the factorisation (chol_retry).
        double c = (g - s[i] * s[j] / n) / (n - 1.0);
        if (i == j) c += jitter;
        A[e] = c;
    }
}

// in-place packed-lower Cholesky; returns false (uniformly) when a pivot is not safely positive
__device__ bool cholesky_packed(double* A, int C, int* flag) {
    for (int k = 0; k < C; ++k) {
        if (threadIdx.x == 0) {
            double d = A[tri(k, k)];
            // LAPACK potrf fails on d <= 0 or NaN; fp32 LAPACK additionally cannot tell a pivot
            // below its rounding noise from zero, so treat those as failures too.
            if (!(d > 0.0)) *flag = 1;
            else A[tri(k, k)] = sqrt(d);
        }
        __syncthreads();
        if (*flag) return false;
        const double inv = 1.0 / A[tri(k, k)];
        __syncthreads();
        for (int i = k + 1 + threadIdx.x; i < C; i += blockDim.x) A[tri(i, k)] *= inv;
        __syncthreads();
        const int m = C - k - 1;   // trailing size
        for (int e = threadIdx.x; e < m * (m + 1) / 2; e += blockDim.x) {
            int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while (tri(a + 1, 0) <= e) ++a;
            while (tri(a, 0) > e) --a;
            int b = e - tri(a, 0);
            int i = k + 1 + a, j = k + 1 + b;
            A[tri(i, j)] -= A[tri(i, k)] * A[tri(j, k)];
        }
        __syncthreads();
    }
    return true;
}

// cov -> L with the reference's retry rule (cWCT.py:115-128): first try as is, then add eps, then
// 2 eps more, ... (cumulative eps*k(k+1)/2).  Returns #retries, or -1 if it never succeeded.
__device__ int chol_retry(double* A, const StatsView& sv, int l, int C, double n, double eps, bool round32, int* flag) {
    for (int k = 0; k <= 64; ++k) {
        __syncthreads();
        if (threadIdx.x == 0) *flag = 0;
        build_cov(A, sv, l, C, n, eps * (double)(k * (k + 1) / 2), round32);
        __syncthreads();
        // rank-deficient labels (n <= C): the exact pivot is 0 and the computed one is rounding
        // noise of either sign; the reference (fp32 LAPACK on an fp32 matrix) sees noise ~1e-8 *
        // |diag|.  Make the outcome deterministic: a pivot below that noise floor is a failure.
        if (cholesky_packed(A, C, flag)) {
            // fp64 factor arithmetic (use_double): torch.linalg.cholesky in fp64 accepts any positive pivot, so do we
            if (!round32) return k;
            __shared__ int bad;
            if (threadIdx.x == 0) bad = 0;
            __syncthreads();
            // relative check of the factor's diagonal against the matrix diagonal
            const StatsView& s2 = sv;
            for (int i = threadIdx.x; i < C; i += blockDim.x) {
                double g = s2.gram[(size_t)l * C * C + (size_t)i * C + i];
                double si = s2.sum[(size_t)l * C + i];
                double cii = (g - si * si / n) / (n - 1.0) + eps * (double)(k * (k + 1) / 2);
                double lii = A[tri(i, i)];
                if (!(lii * lii > 1e-7 * cii)) bad = 1;
            }
            __syncthreads();
            if (!bad) return k;
        }
    }
    return -1;
}

__global__ void __launch_bounds__(256) factor_kernel(FactorArgs fa) {
    extern __shared__ __align__(16) double sm[];
    const int C = fa.C, L = fa.L, l = blockIdx.x;
    const int TRI = C * (C + 1) / 2;
    double* Lc = sm;             // packed lower
    double* Ls = sm + TRI;       // packed lower (per style)
    double* Mix = sm + 2 * TRI;  // packed lower: sum_k a_k Ls_k
    double* mu_c = sm + 3 * TRI; // [C]
    double* mu_m = mu_c + C;     // [C] mixed style mean
    __shared__ int flag;
    __shared__ int ok_s;

    const bool has_c = fa.mode != 2, has_s = fa.mode != 1;
    const StatsView cs = stats_view(const_cast<void*>(fa.cstats), C, L);
    const double nc = has_c ? cs.count[l] : 2.0;
    float* T = fa.T + (size_t)l * C * C;

    bool ok = nc >= 2.0;
    for (int k = 0; has_s && k < fa.n_styles && ok; ++k) {
        const double ns = stats_view(const_cast<void*>(fa.sstats[k]), C, L).count[l];
        ok = ns >= 2.0;
        if (fa.masked)   // cWCT.py:178
            ok = ok && nc > 10.0 && ns > 10.0 && nc / ns < 100.0 && ns / nc < 100.0;
    }
    int retries = 0;
    if (ok && has_c) {
        int r = chol_retry(Lc, cs, l, C, nc, (double)fa.eps, !fa.use_double, &flag);
        if (r < 0) ok = false; else retries += r;
    }
    for (int e = threadIdx.x; e < TRI; e += blockDim.x) {
        Mix[e] = 0.0;
        if (!has_c) Lc[e] = 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        mu_c[i] = (ok && has_c) ? (double)cs.pivot[i] + cs.sum[(size_t)l * C + i] / nc : 0.0;
        mu_m[i] = 0.0;
        if (!has_c) Lc[tri(i, i)] = 1.0;          // coloring only: Lc = I
        if (!has_s) Mix[tri(i, i)] = 1.0;         // whitening only: Mix = I
    }
    __syncthreads();
    for (int k = 0; has_s && k < fa.n_styles && ok; ++k) {
        const StatsView ss = stats_view(const_cast<void*>(fa.sstats[k]), C, L);
        const double ns = ss.count[l];
        int r = chol_retry(Ls, ss, l, C, ns, (double)fa.eps, !fa.use_double, &flag);
        if (r < 0) { ok = false; break; }
        retries += r;
        const double a = (double)fa.alpha_s[k];
        for (int e = threadIdx.x; e < TRI; e += blockDim.x) Mix[e] += a * Ls[e];
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            mu_m[i] += a * ((double)ss.pivot[i] + ss.sum[(size_t)l * C + i] / ns);
        __syncthreads();
    }
    if (threadIdx.x == 0) ok_s = ok ? 1 : 0;
    __syncthreads();
    ok = ok_s != 0;

    const double ac = (double)fa.alpha_c;
    if (ok) {
        // row r of X = Mix Lc^-1 solves  X[r,:] Lc = Mix[r,:]   (back substitution, j = r..0)
        for (int r = threadIdx.x; r < C; r += blockDim.x) {
            // reuse Ls storage row r (length r+1) for the solution; Mix row r is the rhs
            double* x = Ls + tri(r, 0);
            for (int j = r; j >= 0; --j) {
                double s = Mix[tri(r, j)];
                for (int k = j + 1; k <= r; ++k) s -= x[k] * Lc[tri(k, j)];
                x[j] = s / Lc[tri(j, j)];
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
            int i = e / C, j = e - i * C;
            double v = (j <= i) ? (1.0 - ac) * Ls[tri(i, j)] : 0.0;
            if (i == j) v += ac;
            T[e] = (float)v;
        }
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            fa.mu[(size_t)l * C + i] = (float)mu_c[i];
            fa.beta[(size_t)l * C + i] = (float)((1.0 - ac) * mu_m[i] + ac * mu_c[i]);
        }
    } else {
        for (int e = threadIdx.x; e < C * C; e += blockDim.x) T[e] = (e / C == e % C) ? 1.f : 0.f;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            fa.mu[(size_t)l * C + i] = 0.f;
            fa.beta[(size_t)l * C + i] = 0.f;
        }
    }
    if (threadIdx.x == 0) {
        fa.valid[l] = ok ? 1 : 0;
        fa.status[l] = ok ? retries : ((nc >= 2.0 && !fa.masked) ? -1 : 0);
    }
}

// ------------------------------------------------------------------------------------------
// cholesky_dec of a given C x C matrix (cWCT.py:111-132): L = chol(cov) with the cumulative eps*I retry,
// optionally inverted (torch.inverse(L): forward substitution, column per thread).  One CTA.
// ------------------------------------------------------------------------------------------
template <typename TIO>
__global__ void __launch_bounds__(256) cholesky_dec_kernel(const TIO* __restrict__ cov, int C, double eps, int invert,
                                                           int round32, TIO* __restrict__ out, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm[];
    const int TRI = C * (C + 1) / 2;
    double* A = sm;              // packed lower factor
    double* X = sm + TRI;        // packed lower inverse
    __shared__ int flag;
    __shared__ int bad;
    int k = 0;
    for (; k <= 64; ++k) {
        __syncthreads();
        if (threadIdx.x == 0) { flag = 0; bad = 0; }
        const double jit = eps * (double)(k * (k + 1) / 2);
        for (int e = threadIdx.x; e < TRI; e += blockDim.x) {
            int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
            while (tri(i + 1, 0) <= e) ++i;
            while (tri(i, 0) > e) --i;
            const int j = e - tri(i, 0);
            double c = (double)cov[(size_t)i * C + j];     // potrf reads the lower triangle
            if (i == j) {
                c += jit;
                if (round32) c = (double)(float)c;         // conv + iden * eps is an fp32 sum in the reference
            }
            A[e] = c;
        }
        __syncthreads();
        if (!cholesky_packed(A, C, &flag)) continue;
        if (!round32) break;
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            const double cii = (double)cov[(size_t)i * C + i] + jit, lii = A[tri(i, i)];
            if (!(lii * lii > 1e-7 * cii)) bad = 1;
        }
        __syncthreads();
        if (!bad) break;
    }
    const bool ok = k <= 64;
    __syncthreads();
    if (ok && invert) {
        // column j of X = L^-1: X[j][j] = 1 / L[j][j];  X[i][j] = -(sum_{m=j..i-1} L[i][m] X[m][j]) / L[i][i]
        for (int j = threadIdx.x; j < C; j += blockDim.x) {
            X[tri(j, j)] = 1.0 / A[tri(j, j)];
            for (int i = j + 1; i < C; ++i) {
                double s = 0.0;
                for (int m = j; m < i; ++m) s += A[tri(i, m)] * X[tri(m, j)];
                X[tri(i, j)] = -s / A[tri(i, i)];
            }
        }
        __syncthreads();
    }
    const double* R = (ok && invert) ? X : A;
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) {
        const int i = e / C, j = e - i * C;
        double v = ok ? (j <= i ? R[tri(i, j)] : 0.0) : nan("");
        out[e] = (TIO)v;
    }
    if (threadIdx.x == 0) *status = ok ? k : -1;
}

// Packed fp32 FMA (Blackwell FFMA2): (d0, d1) += (a, a) * (b0, b1), IEEE fma per half — bit-identical to two fmaf().
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

// ------------------------------------------------------------------------------------------
// apply:  out[:,p] = T[l(p)] (x[:,p] - mu[l(p)]) + beta[l(p)]
// thread tile = 8 output channels x 4 pixels; CTA tile = CP channels x PX pixels
// ------------------------------------------------------------------------------------------
template <int CP>
struct ApplyCfg {
    static constexpr int NCG = CP / 8;          // channel groups
    static constexpr int NPG = 256 / NCG;       // pixel groups
    static constexpr int PX = NPG * 4;          // pixels per tile
    static constexpr size_t SMEM = (size_t)(CP * PX + CP * CP + 2 * CP) * sizeof(float) + PX * sizeof(int);
};

template <int CP>
__global__ void __launch_bounds__(256) apply_kernel(const float* __restrict__ feat, float* __restrict__ out, int C,
                                                    long long n, const uint8_t* __restrict__ labels, int L,
                                                    const float* __restrict__ T, const float* __restrict__ mu,
                                                    const float* __restrict__ beta, const int* __restrict__ valid,
                                                    long long n_tiles) {
    using Cfg = ApplyCfg<CP>;
    constexpr int PX = Cfg::PX, NPG = Cfg::NPG;
    extern __shared__ __align__(16) float smf[];
    float* xs = smf;                   // [CP][PX]
    float* Tt = xs + CP * PX;          // [k][c] (transposed)
    float* mu_s = Tt + CP * CP;        // [CP]
    float* be_s = mu_s + CP;           // [CP]
    int* lab_s = (int*)(be_s + CP);    // [PX]
    const int tid = threadIdx.x;
    const int pg = tid % NPG, cg = tid / NPG;
    int cached = -1;

    __shared__ unsigned int pres[8];   // labels present in the tile (bit l)

    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long p0 = t * PX;
        const int npx = (int)(n - p0 < PX ? n - p0 : PX);
        __syncthreads();   // previous tile's smem fully consumed
        // ---- labels of this tile and the set of labels present
        if (tid < 8) pres[tid] = (labels || tid) ? 0u : 1u;          // unmasked: label 0 only
        __syncthreads();
        if (labels) {
            for (int i = tid; i < npx; i += 256) {
                const int l = (int)labels[p0 + i];
                lab_s[i] = l;
                atomicOr(&pres[l >> 5], 1u << (l & 31));
            }
        }
        // ---- stage x (raw); 16-byte loads when the rows are 16-byte aligned and the tile is full
        const bool vec = ((n & 3) == 0) && npx == PX && ((((uintptr_t)feat | (uintptr_t)out) & 15) == 0);
        if (vec) {
            // all loads of the tile are issued before the first shared-memory store (memory-level parallelism)
            constexpr int NIT = CP * (PX / 4) / 256;
            float4 v[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = it * 256 + tid, k = i / (PX / 4), p4 = i - k * (PX / 4);
                v[it] = (k < C) ? __ldg(reinterpret_cast<const float4*>(feat + (size_t)k * n + p0) + p4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < NIT; ++it) {
                const int i = it * 256 + tid, k = i / (PX / 4), p4 = i - k * (PX / 4);
                reinterpret_cast<float4*>(xs + k * PX)[p4] = v[it];
            }
        } else {
            for (int i = tid; i < CP * PX; i += 256) {
                int k = i / PX, px = i - k * PX;
                xs[i] = (k < C && px < npx) ? __ldg(feat + (size_t)k * n + p0 + px) : 0.f;
            }
        }
        __syncthreads();
        int n_present = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) n_present += __popc(pres[w8]);
        const bool uniform = n_present == 1;

        // ---- one pass per label present (one for almost every tile; two or three where regions meet): the label's
        //      transform goes to shared memory and the whole tile is multiplied; a mixed tile stores only the pixels
        //      that carry the label.  (The first version fell back to a per-pixel product with T read from global
        //      memory for every mixed tile: 4x the time of the unmasked kernel on blocky masks.)
        for (int w8 = 0; w8 < 8; ++w8) {
            unsigned int bits = pres[w8];
            while (bits) {
                const int lbl = w8 * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                const bool v = lbl < L && valid[lbl < L ? lbl : 0];
                if (!v) {   // invalid label: content features pass through untouched (cWCT.py:80-84)
                    if (out != feat)
                        for (int i = tid; i < CP * PX; i += 256) {
                            int k = i / PX, px = i - k * PX;
                            if (k < C && px < npx && (uniform || lab_s[px] == lbl)) out[(size_t)k * n + p0 + px] = xs[i];
                        }
                    continue;
                }
                if (lbl != cached) {
                    __syncthreads();           // the previous label's product is done with Tt
                    for (int i = tid; i < CP * CP; i += 256) {
                        int k = i / CP, c = i - k * CP;     // Tt[k][c] = T[c][k]
                        Tt[i] = (c < C && k < C) ? __ldg(T + ((size_t)lbl * C + c) * C + k) : 0.f;
                    }
                    for (int i = tid; i < CP; i += 256) {
                        mu_s[i] = i < C ? mu[(size_t)lbl * C + i] : 0.f;
                        be_s[i] = i < C ? beta[(size_t)lbl * C + i] : 0.f;
                    }
                    cached = lbl;
                    __syncthreads();
                }
                float acc[8][4];
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
                for (int k = 0; k < CP; ++k) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xs + k * PX + pg * 4);
                    const float4 t0 = *reinterpret_cast<const float4*>(Tt + k * CP + cg * 8);
                    const float4 t1 = *reinterpret_cast<const float4*>(Tt + k * CP + cg * 8 + 4);
                    const float m = mu_s[k];
                    const float xv[4] = {x4.x - m, x4.y - m, x4.z - m, x4.w - m};
                    const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                    for (int a = 0; a < 8; ++a) {       // packed FFMA2: two pixels per issue slot (a plain FFMA issues every
                        fma2(acc[a][0], acc[a][1], tv[a], xv[0], xv[1]);   // other cycle per sub-partition)
                        fma2(acc[a][2], acc[a][3], tv[a], xv[2], xv[3]);
                    }
                }
                bool mine[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) mine[b] = (pg * 4 + b < npx) && (uniform || lab_s[pg * 4 + b] == lbl);
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int c = cg * 8 + a;
                    if (c >= C) continue;
                    const float bt = be_s[c];
                    if (vec && uniform) {
                        *reinterpret_cast<float4*>(out + (size_t)c * n + p0 + pg * 4) =
                            make_float4(acc[a][0] + bt, acc[a][1] + bt, acc[a][2] + bt, acc[a][3] + bt);
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if (mine[b]) out[(size_t)c * n + p0 + pg * 4 + b] = acc[a][b] + bt;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// apply on the P4 half-states, in place (fused video path: the latent is never materialised):
//   for every state pixel p and sub-position j:  s[jC .. jC+C) <- T (s[jC .. jC+C) - mu) + beta
// which is exactly spread -> apply -> gather (models/RevResNet.py:140-146, cWCT.py:147,161-162, RevResNet.py:149-152).
// Tile = one sub-position x PX consecutive interior pixels; same thread tile (8 channels x 4 pixels, FFMA2) as above.
// ------------------------------------------------------------------------------------------
template <int CP>
__global__ void __launch_bounds__(256, CP == 32 ? 4 : 1) apply_state_kernel(float* __restrict__ x1, float* __restrict__ x2, int Ch, int h, int w,
                                                          const float* __restrict__ T, const float* __restrict__ mu,
                                                          const float* __restrict__ beta, const int* __restrict__ valid,
                                                          int tiles_per_sub, int n_tiles) {
    using Cfg = ApplyCfg<CP>;
    constexpr int PX = Cfg::PX, NPG = Cfg::NPG, G = CP / 4;
    if (!valid[0]) return;                  // factorisation failed: the content features pass through (cWCT.py:80-84)
    extern __shared__ __align__(16) float smf[];
    float* xs = smf;                   // [CP][PX]
    float* Tt = xs + CP * PX;          // [k][c] (transposed)
    float* mu_s = Tt + CP * CP;        // [CP]
    float* be_s = mu_s + CP;           // [CP]
    const int tid = threadIdx.x;
    const int pg = tid % NPG, cg = tid / NPG;
    const int n_px = h * w, Wp = w + 2, gph = Ch / 4;
    const size_t plane = p4_plane_px(h, w);
    for (int i = tid; i < CP * CP; i += 256) {
        const int k = i / CP, c = i - k * CP;     // Tt[k][c] = T[c][k]
        Tt[i] = __ldg(T + (size_t)c * CP + k);
    }
    for (int i = tid; i < CP; i += 256) { mu_s[i] = mu[i]; be_s[i] = beta[i]; }

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int j = t / tiles_per_sub, p0 = (t - j * tiles_per_sub) * PX;
        const int npx = min(PX, n_px - p0);
        const int g0 = j * G;                                           // first state group of this sub-position
        float4* base = reinterpret_cast<float4*>(g0 < gph ? x1 : x2) + (size_t)(g0 < gph ? g0 : g0 - gph) * plane;
        __syncthreads();   // previous tile's smem fully consumed (and Tt staged)
        {
            // item (it, tid): group g = (it * 256 + tid) / PX, pixel px = tid % PX — the pixel is fixed per thread, so
            // one division per tile; ALL loads are issued before the first shared-memory store (a load followed by its
            // dependent store serialises one DRAM latency per item)
            constexpr int NIT = G * PX / 256;
            const int px = tid % PX, gq = tid / PX;
            const int p = p0 + px, y = p / w, x = p - y * w;
            const float4* src = base + (size_t)(y + 1) * Wp + x + 1;
            float4 v[NIT];
#pragma unroll
            for (int it = 0; it < NIT; ++it)
                v[it] = (px < npx) ? src[(size_t)(it * (256 / PX) + gq) * plane] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int it = 0; it < NIT; ++it) {                 // staged centred: x - mu
                const int g = it * (256 / PX) + gq;
                const float4 m4 = *reinterpret_cast<const float4*>(mu_s + 4 * g);
                xs[(4 * g + 0) * PX + px] = v[it].x - m4.x; xs[(4 * g + 1) * PX + px] = v[it].y - m4.y;
                xs[(4 * g + 2) * PX + px] = v[it].z - m4.z; xs[(4 * g + 3) * PX + px] = v[it].w - m4.w;
            }
        }
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        // this thread's 4 pixels are pg, pg + NPG, pg + 2 NPG, pg + 3 NPG: consecutive lanes own consecutive pixels, so
        // the 16-byte P4 stores below are contiguous across the warp (4 pixels per lane would write half-used sectors)
#pragma unroll 4
        for (int k = 0; k < CP; ++k) {
            const float* xr = xs + k * PX + pg;
            const float4 t0 = *reinterpret_cast<const float4*>(Tt + k * CP + cg * 8);
            const float4 t1 = *reinterpret_cast<const float4*>(Tt + k * CP + cg * 8 + 4);
            const float xv[4] = {xr[0], xr[NPG], xr[2 * NPG], xr[3 * NPG]};
            const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                fma2(acc[a][0], acc[a][1], tv[a], xv[0], xv[1]);
                fma2(acc[a][2], acc[a][3], tv[a], xv[2], xv[3]);
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int px = pg + b * NPG;
            if (px >= npx) continue;
            const int p = p0 + px, y = p / w, x = p - y * w;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c = cg * 8 + half * 4;
                const float4 o = make_float4(acc[half * 4 + 0][b] + be_s[c], acc[half * 4 + 1][b] + be_s[c + 1],
                                             acc[half * 4 + 2][b] + be_s[c + 2], acc[half * 4 + 3][b] + be_s[c + 3]);
                p4_store(base + (size_t)(c >> 2) * plane, h, w, y, x, o);
            }
        }
    }
}

template <int CP>
static int launch_apply_state_cfg(float* x1, float* x2, int Ch, int h, int w, const float* T, const float* mu,
                                  const float* beta, const int* valid, cudaStream_t st) {
    using Cfg = ApplyCfg<CP>;
    static PerDeviceOnce smem_once;
    auto kern = apply_state_kernel<CP>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    const int nsub = 2 * Ch / CP;
    const int tiles_per_sub = cdiv((long long)h * w, Cfg::PX), n_tiles = nsub * tiles_per_sub;
    const int grid = std::min(n_tiles, num_sms() * (CP == 32 ? 6 : 2));
    const double n = (double)nsub * h * w;
    ProfScope prof(st, CP == 32 ? "cwct_apply_state c32" : "cwct_apply_state c128", 2.0 * CP * CP * n, 8.0 * CP * n);
    kern<<<grid, 256, Cfg::SMEM, st>>>(x1, x2, Ch, h, w, T, mu, beta, valid, tiles_per_sub, n_tiles);
    return check_launch("cwct_apply_state");
}

// in-place transform of the latent held as P4 half-states x1 | x2 (Ch channels each), C latent channels, label 0
int launch_apply_state(float* x1, float* x2, int C, int Ch, int h, int w, const float* T, const float* mu, const float* beta,
                       const int* valid, cudaStream_t st) {
    VST_REQUIRE(C == 32 || C == 128, "apply_state: C = %d not supported", C);
    if (C == 32) return launch_apply_state_cfg<32>(x1, x2, Ch, h, w, T, mu, beta, valid, st);
    return launch_apply_state_cfg<128>(x1, x2, Ch, h, w, T, mu, beta, valid, st);
}

// statistics of that latent (one label): zero-initialises `stats`, pivot from sub-position 0, tensor-core Gram
int launch_stats_state(const float* x1, const float* x2, int C, int Ch, int h, int w, void* stats, cudaStream_t st) {
    StatsView sv = stats_view(stats, C, 1);
    VST_CUDA_OK(cudaMemsetAsync(stats, 0, stats_doubles(C, 1) * sizeof(double), st));
    count_launch();
    pivot_state_kernel<<<C, 256, 0, st>>>(x1, h, w, sv.pivot);
    if (check_launch("cwct_pivot_state")) return 1;
    const double n = (double)(2 * Ch / C) * h * w;
    ProfScope prof(st, C <= 32 ? "cwct_gram_state c32" : "cwct_gram_state c128", 2.0 * C * C * n, 4.0 * C * n);
    return launch_gram_tc_state(x1, x2, sv.pivot, sv.count, sv.sum, sv.gram, C, Ch, h, w, st);
}

template <int CP>
static int launch_apply(const float* feat, float* out, int C, long long n, const uint8_t* labels, int L, const float* T,
                        const float* mu, const float* beta, const int* valid, cudaStream_t st) {
    using Cfg = ApplyCfg<CP>;
    static PerDeviceOnce smem_once;
    auto kern = apply_kernel<CP>;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, kern, (int)Cfg::SMEM));
    long long tiles = (n + Cfg::PX - 1) / Cfg::PX;
    int grid = (int)std::min<long long>(tiles, (long long)num_sms() * (CP == 32 ? 6 : 2));
    ProfScope prof(st, CP == 32 ? "cwct_apply c32" : "cwct_apply c128", 2.0 * C * C * (double)n,
                   8.0 * C * (double)n + (labels ? (double)n : 0.0));
    kern<<<grid, 256, Cfg::SMEM, st>>>(feat, out, C, n, labels, L, T, mu, beta, valid, tiles);
    return check_launch("cwct_apply");
}

}  // namespace vst

using namespace vst;

extern "C" size_t vst_cwct_stats_bytes(int C, int n_labels) {
    if (C < 1 || n_labels < 1) return 0;
    return align_up(stats_doubles(C, n_labels) * sizeof(double) + (size_t)C * sizeof(float), 16);
}

static int cwct_stats_impl(const float* feat, int C, long long n, int H, int W, const uint8_t* labels, int n_labels, void* stats,
                           void* stream);

extern "C" int vst_cwct_stats(const float* feat, int C, long long n, const uint8_t* labels, int n_labels, void* stats,
                              void* stream) {
    return cwct_stats_impl(feat, C, n, 0, 0, labels, n_labels, stats, stream);
}

extern "C" int vst_cwct_stats2d(const float* feat, int C, int H, int W, const uint8_t* labels, int n_labels, void* stats,
                                void* stream) {
    VST_REQUIRE(H >= 1 && W >= 1, "vst_cwct_stats2d: empty feature map");
    return cwct_stats_impl(feat, C, (long long)H * W, H, W, labels, n_labels, stats, stream);
}

static int cwct_stats_impl(const float* feat, int C, long long n, int H, int W, const uint8_t* labels, int n_labels, void* stats,
                           void* stream) {
    VST_REQUIRE(feat && stats, "vst_cwct_stats: null argument");
    VST_REQUIRE(C >= 1 && C <= 128, "cWCT supports 1 <= C <= 128 channels (got %d)", C);
    VST_REQUIRE(n >= 1, "vst_cwct_stats: empty feature map");
    VST_REQUIRE(n_labels >= 1 && n_labels <= VST_MAX_LABELS, "n_labels %d out of range", n_labels);
    VST_REQUIRE(labels || n_labels == 1, "unmasked stats need n_labels == 1");
    VST_REQUIRE(((uintptr_t)stats & 7) == 0, "stats buffer must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    StatsView sv = stats_view(stats, C, n_labels);
    VST_CUDA_OK(cudaMemsetAsync(stats, 0, stats_doubles(C, n_labels) * sizeof(double), st));
    count_launch();
    pivot_kernel<<<C, 256, 0, st>>>(feat, n, sv.pivot);
    if (check_launch("cwct_pivot")) return 1;
    const int sms = num_sms();
    ProfScope prof(st, C <= 32 ? "cwct_gram c32" : "cwct_gram c128", 2.0 * C * C * (double)n,
                   4.0 * C * (double)n + (labels ? (double)n : 0.0));
    static int use_tc = -1;
    if (use_tc < 0) { const char* e = getenv("VST_GRAM_TC"); use_tc = e ? atoi(e) : 1; }
    if (!labels && use_tc && gram_tc_eligible(C, n) && (((uintptr_t)feat) & 15) == 0)
        return launch_gram_tc(feat, sv.pivot, sv.count, sv.sum, sv.gram, C, n, st);     // tensor cores (gram_tc.cu)
    if (labels && use_tc && H > 0 && gram_tc_masked_eligible(C, H, W) && (((uintptr_t)feat) & 15) == 0)
        return launch_gram_tc_masked(feat, labels, sv.pivot, sv.count, sv.sum, sv.gram, C, n_labels, H, W, st);
    if (C <= 32) {
        int grid = labels ? sms * 3 : sms * 2;
        long long warps = (long long)grid * 8;
        long long chunk = ((n + warps - 1) / warps + 31) / 32 * 32;
        grid = (int)((n + chunk * 8 - 1) / (chunk * 8));
        if (labels) gram32_kernel<true><<<grid, 256, 0, st>>>(feat, C, n, labels, n_labels, sv, chunk);
        else gram32_kernel<false><<<grid, 256, 0, st>>>(feat, C, n, labels, n_labels, sv, chunk);
    } else {
        int grid = sms * 3;
        long long chunk = ((n + grid - 1) / grid + 31) / 32 * 32;
        grid = (int)((n + chunk - 1) / chunk);
        if (labels) gram128_kernel<true><<<grid, 256, 0, st>>>(feat, C, n, labels, n_labels, sv, chunk);
        else gram128_kernel<false><<<grid, 256, 0, st>>>(feat, C, n, labels, n_labels, sv, chunk);
    }
    return check_launch("cwct_gram");
}

namespace vst {
size_t cwct_stats_bytes(int C, int n_labels) { return align_up(stats_doubles(C, n_labels) * sizeof(double) + (size_t)C * sizeof(float), 16); }

int launch_factor(int mode, const void* content_stats, const void* const* style_stats, const float* alpha_s,
                         int n_styles, float alpha_c, float eps, int C, int n_labels, int masked, int use_double, float* T,
                         float* mu, float* beta, int* valid, int* status, cudaStream_t st) {
    VST_REQUIRE(T && mu && beta && valid && status, "vst_cwct_factor: null output");
    VST_REQUIRE(mode == 2 || content_stats, "vst_cwct_factor: null content statistics");
    VST_REQUIRE(mode == 1 || (style_stats && alpha_s), "vst_cwct_factor: null style statistics");
    VST_REQUIRE(C >= 1 && C <= 128, "cWCT supports 1 <= C <= 128 channels (got %d)", C);
    VST_REQUIRE(mode == 1 || (n_styles >= 1 && n_styles <= VST_MAX_STYLES), "n_styles %d out of range", n_styles);
    VST_REQUIRE(n_labels >= 1 && n_labels <= VST_MAX_LABELS, "n_labels %d out of range", n_labels);
    VST_REQUIRE(!masked || n_styles == 1, "masked transfer takes exactly one style (cWCT.py:49)");
    FactorArgs fa;
    fa.cstats = content_stats;
    fa.n_styles = mode == 1 ? 0 : n_styles;
    for (int k = 0; k < fa.n_styles; ++k) {
        VST_REQUIRE(style_stats[k], "style_stats[%d] is null", k);
        fa.sstats[k] = style_stats[k];
        fa.alpha_s[k] = alpha_s[k];
    }
    fa.alpha_c = alpha_c; fa.eps = eps;
    fa.C = C; fa.L = n_labels; fa.masked = masked; fa.use_double = use_double; fa.mode = mode;
    fa.T = T; fa.mu = mu; fa.beta = beta; fa.valid = valid; fa.status = status;
    const size_t smem = ((size_t)3 * (C * (C + 1) / 2) + 2 * C) * sizeof(double);
    static PerDeviceOnce smem_once;
    VST_CUDA_OK(ensure_dyn_smem(smem_once, factor_kernel, 3 * (128 * 129 / 2) * 8 + 2 * 128 * 8));
    ProfScope prof(st, "cwct_factor", 0.0, 0.0);
    factor_kernel<<<n_labels, 256, smem, st>>>(fa);
    return check_launch("cwct_factor");
}
}  // namespace vst

extern "C" int vst_cwct_factor(const void* content_stats, const void* const* style_stats, const float* alpha_s,
                               int n_styles, float alpha_c, float eps, int C, int n_labels, int masked, int use_double,
                               float* T, float* mu, float* beta, int* valid, int* status, void* stream) {
    return launch_factor(0, content_stats, style_stats, alpha_s, n_styles, alpha_c, eps, C, n_labels, masked, use_double, T,
                         mu, beta, valid, status, (cudaStream_t)stream);
}

extern "C" int vst_cwct_whiten_factor(const void* stats, float eps, int C, int use_double, float* T, float* mu, float* beta,
                                      int* valid, int* status, void* stream) {
    return launch_factor(1, stats, nullptr, nullptr, 0, 0.f, eps, C, 1, 0, use_double, T, mu, beta, valid, status,
                         (cudaStream_t)stream);
}

extern "C" int vst_cwct_color_factor(const void* stats, float eps, int C, int use_double, float* T, float* mu, float* beta,
                                     int* valid, int* status, void* stream) {
    const void* ss[1] = {stats};
    const float one[1] = {1.f};
    return launch_factor(2, nullptr, ss, one, 1, 0.f, eps, C, 1, 0, use_double, T, mu, beta, valid, status,
                         (cudaStream_t)stream);
}

extern "C" int vst_cwct_cholesky(const void* cov, int C, int is_double, float eps, int invert, void* out, int* status,
                                 void* stream) {
    VST_REQUIRE(cov && out && status, "vst_cwct_cholesky: null argument");
    VST_REQUIRE(C >= 1 && C <= 128, "cWCT supports 1 <= C <= 128 channels (got %d)", C);
    const size_t smem = (size_t)2 * (C * (C + 1) / 2) * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_double) {
        static PerDeviceOnce once;
        VST_CUDA_OK(ensure_dyn_smem(once, cholesky_dec_kernel<double>, 2 * (128 * 129 / 2) * 8));
        cholesky_dec_kernel<double><<<1, 256, smem, st>>>((const double*)cov, C, (double)eps, invert, 0, (double*)out, status);
    } else {
        static PerDeviceOnce once;
        VST_CUDA_OK(ensure_dyn_smem(once, cholesky_dec_kernel<float>, 2 * (128 * 129 / 2) * 8));
        cholesky_dec_kernel<float><<<1, 256, smem, st>>>((const float*)cov, C, (double)eps, invert, 1, (float*)out, status);
    }
    return check_launch("cwct_cholesky");
}

extern "C" int vst_cwct_apply(const float* feat, float* out, int C, long long n, const uint8_t* labels, int n_labels,
                              const float* T, const float* mu, const float* beta, const int* valid, void* stream) {
    VST_REQUIRE(feat && out && T && mu && beta && valid, "vst_cwct_apply: null argument");
    VST_REQUIRE(C >= 1 && C <= 128, "cWCT supports 1 <= C <= 128 channels (got %d)", C);
    VST_REQUIRE(n >= 1, "vst_cwct_apply: empty feature map");
    VST_REQUIRE(n_labels >= 1 && n_labels <= VST_MAX_LABELS, "n_labels %d out of range", n_labels);
    if (C <= 32) return launch_apply<32>(feat, out, C, n, labels, n_labels, T, mu, beta, valid, (cudaStream_t)stream);
    return launch_apply<128>(feat, out, C, n, labels, n_labels, T, mu, beta, valid, (cudaStream_t)stream);
}
