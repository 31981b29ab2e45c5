// Shared device helpers: warp/block reductions, the log-sum-exp monoid, control block, error handling.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

#include "fastmath.cuh"

namespace smcb {

#define SMCB_FULL 0xffffffffu

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define SMCB_CUDA(x) ::smcb::cuda_check((x), #x, __FILE__, __LINE__)

// Programmatic dependent launch (sm_90+): the kernels of a filter step form a chain in one stream; launched with this helper,
// the blocks of the next kernel are scheduled onto the SMs while the previous kernel drains, and wait in
// pdl_wait_for_predecessor() until all of its memory is visible -- the launch latency between dependent kernels
// disappears.  A kernel launched the ordinary way passes both calls as no-ops.  Disabled with SMCB200_NO_PDL=1.
__device__ __forceinline__ void pdl_wait_for_predecessor() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release_successor() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SMCB200_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
template <class... KArgs, class... Args>
inline void launch_chained(void (*kernel)(KArgs...), int grid, int block, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    SMCB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// Resampling modes / history modes: numeric values are the reference's (inst/include/sampler.h:45-68).
enum ResampleMode : int { MULTINOMIAL = 0, RESIDUAL = 1, STRATIFIED = 2, SYSTEMATIC = 3 };
enum HistoryMode : int { HISTORY_NONE = 0, HISTORY_RAM = 1, HISTORY_AL = 2 };

__device__ __forceinline__ double shfl_down_d(double v, int off) { return __shfl_down_sync(SMCB_FULL, v, off); }
__device__ __forceinline__ double shfl_xor_d(double v, int off) { return __shfl_xor_sync(SMCB_FULL, v, off); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_xor_d(v, o));
    return v;
}

// (m, s, q, g[]) with s = sum exp(lw - m), q = sum exp(2(lw - m)), g_j = sum exp(lw - m) h_j: an associative,
// commutative monoid under max-rescaled addition.  It carries everything the reference
// derives from stableLogSumWeights (population.h:46-50): logNC increment = m + log s (sampler.h:498,710),
// ESS = s^2 / q (sampler.h:413-417), weighted means g/s (sampler.h:559-571).
template <int G>
struct LseAcc {
    double m, s, q;
    double g[G > 0 ? G : 1];
    __device__ __forceinline__ void clear() {
        m = -INFINITY; s = 0.0; q = 0.0;
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = 0.0;
    }
    // add one element: exactly one exp, straight-line (selects, no branches) so that a thread's particles interleave
    __device__ __forceinline__ void add(double lw, const double* h) {
        const double d = lw - m;            // +inf for the first finite element, NaN when both are -inf
        const bool up = d > 0.0;            // lw becomes the new running maximum
        double e = fm_exp(-fabs(d));
        e = (lw == -INFINITY) ? 0.0 : e;    // weightless particle contributes nothing (also covers the NaN case)
        const double e2 = e * e;
        s = up ? fma(s, e, 1.0) : s + e;
        q = up ? fma(q, e2, 1.0) : q + e2;
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = up ? fma(g[j], e, h[j]) : fma(e, h[j], g[j]);
        m = up ? lw : m;
    }
    __device__ __forceinline__ void merge(const LseAcc& o) {
        if (o.m == -INFINITY) return;
        if (m == -INFINITY) { *this = o; return; }
        double M = fmax(m, o.m);
        double ea = fm_exp(m - M), eb = fm_exp(o.m - M);
        s = s * ea + o.s * eb;
        q = q * (ea * ea) + o.q * (eb * eb);
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = g[j] * ea + o.g[j] * eb;
        m = M;
    }
    __device__ __forceinline__ LseAcc shfl_xor(int off) const {
        LseAcc r;
        r.m = shfl_xor_d(m, off); r.s = shfl_xor_d(s, off); r.q = shfl_xor_d(q, off);
#pragma unroll
        for (int j = 0; j < G; ++j) r.g[j] = shfl_xor_d(g[j], off);
        return r;
    }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { LseAcc t = shfl_xor(o); merge(t); }
    }
};

// Sums of weights exp(lw - c) against a shift c that is common to the whole step: s = sum e, q = sum e^2, g_j = sum e h_j, and
// the running maximum of lw (which validates c afterwards and predicts the next one).  One exp per element, no selects; the
// exponential is returned so that the caller can store it -- the resampling scan then normalises with one multiplication
// instead of a second exp.  logNC increment = c + log s (population.h:46-50), ESS = s^2 / q (sampler.h:413-417).
template <int G>
struct WAcc {
    double s, q, mx;
    double g[G > 0 ? G : 1];
    __device__ __forceinline__ void clear() {
        s = 0.0; q = 0.0; mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = 0.0;
    }
    __device__ __forceinline__ double add(double lw, double c, const double* h) {
        double e = fm_exp(lw - c);          // 0 for lw = -inf
        e = (lw == lw) ? e : lw;            // a NaN log-weight poisons the sums, as it does in the reference
        s += e;
        q = fma(e, e, q);
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = fma(e, h[j], g[j]);
        mx = fmax(mx, lw);
        return e;
    }
    __device__ __forceinline__ void warp_reduce() {
        s = warp_sum(s); q = warp_sum(q); mx = warp_max(mx);
#pragma unroll
        for (int j = 0; j < G; ++j) g[j] = warp_sum(g[j]);
    }
};

}  // namespace smcb
