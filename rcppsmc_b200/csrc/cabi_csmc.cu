// Conditional SMC on device: the conditional bootstrap particle filter of compareNCestimates and its plain counterpart.
// Replaces smc::conditionalSampler (reference inst/include/conditionalSampler.h:166-235 Initialise, :303-361 IterateEss,
// :377-655 conditionalResample, :657-661 MoveReferenceParticle; moveset.h:170-180 DoConditionalMove) on the LGSS model of
// src/cSMCexamples.cpp:141-218, the driver compareNCestimates_imp (:36-139) and the ancestry getters of the AL history
// (sampler.h:445-476 GetALineInd / GetALineSpace).
//
// Shape of the work.  The reference workload is many small filters (N ~ 10^3 particles, 4 resampling schemes x simNum
// simulations), each strictly sequential in time.  One CTA runs one CHAIN of conditional filters start to finish -- every
// time step, its reductions, scans and searches stay inside the CTA (no launches, no grid synchronisation) -- and chains
// run side by side on the 148 SMs.  A chain is the unit of sequential dependence: the reference keeps ONE sampler object
// for all runs, and its referenceTrajectoryIndices carry over from run to run, so runs of one chain execute
// in order on shared index state.  Every conditional resampling scheme is "scan, then one independent binary search per
// offspring" (the reference walks the cumulative weights sequentially, O(N^2) worst case through arma::find(tail...)).
#include <cmath>
#include <functional>
#include <vector>

#include "cabi_common.cuh"
#include "models.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {

constexpr int CS_THREADS = 512;

struct CsmcArgs {
    LGSS::Params mp;
    long long T, n;
    long long n_chains, runs_per_chain;
    const int* modes;        // [n_chains * runs_per_chain]
    const double* ref;       // [n_chains * runs_per_chain][T]
    unsigned long long seed;
    const double* zin;       // [runs][T][n] or null
    const double* uin;       // [runs][T][n + 2] or null: {K_t draw, systematic V, offspring i -> 2 + i}
    // per-chain scratch
    double* xa; double* xb; double* lw; double* w; double* W; double* P;
    int* L; uint32_t* anc; uint32_t* refidx;
    // outputs
    double* lognc;           // [runs]
    uint32_t* refidx_out;    // [runs][T]
    uint32_t* anc_out;       // [runs][T][n] or null  (AL history)
    double* val_out; double* lw_out;
    int* res_out;            // [runs][T] or null
};

struct Shared {
    double red[3][CS_THREADS / 32];
    double scan_d[CS_THREADS];
    int scan_i[CS_THREADS];
    double bc[4];
    long long bi[2];
};

// in-place inclusive scan of a[0..n) by the whole CTA (thread t owns a contiguous chunk); returns the total
template <class V>
__device__ V cta_scan(V* a, long long n, V* part) {
    const int B = blockDim.x, tid = threadIdx.x;
    const long long c = (n + B - 1) / B, lo = tid * c, hi = lo + c < n ? lo + c : n;
    V s = 0;
    for (long long i = lo; i < hi; ++i) s += a[i];
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < B; off <<= 1) {
        V v = tid >= off ? part[tid - off] : (V)0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    V run = tid > 0 ? part[tid - 1] : (V)0;
    const V total = part[B - 1];
    for (long long i = lo; i < hi; ++i) { run += a[i]; a[i] = run; }
    __syncthreads();
    return total;
}

// Rcpp::sample(n, 1, ., p) on the running sums C of p (restated as in the oracle: first k < n-1 with u <= C_k / C_{n-1}, else n-1)
__device__ __forceinline__ long long sample_cdf(const double* C, long long n, double tot, double u) {
    long long lo = 0, hi = n - 1;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (u <= C[mid] / tot) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// first j with W[j] > u, clamped to n-1 (arma::find(W.tail(...) > u, 1, "first"), conditionalSampler.h:476,550)
__device__ __forceinline__ long long first_above(const double* W, long long n, double u) {
    long long lo = 0, hi = n - 1;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (W[mid] > u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ double cs_uniform(const CsmcArgs& a, long long run, long long t, long long slot) {
    if (a.uin) return a.uin[((size_t)run * (size_t)a.T + (size_t)t) * (size_t)(a.n + 2) + (size_t)slot];
    return u64_to_unit(philox_draw(a.seed, STREAM_RESAMPLE, (uint32_t)t, (uint64_t)slot, (uint32_t)run).a);
}

__global__ void __launch_bounds__(CS_THREADS) k_csmc_chain(CsmcArgs a) {
    __shared__ Shared sh;
    const int tid = threadIdx.x, B = CS_THREADS, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n, T = a.T, chain = blockIdx.x;
    double* xs[2] = {a.xa + chain * n, a.xb + chain * n};
    double* lw = a.lw + chain * n;
    double* w = a.w + chain * n;
    double* W = a.W + chain * n;
    double* P = a.P + chain * n;
    int* Lx = a.L + chain * n;
    uint32_t* anc = a.anc + chain * n;
    uint32_t* refidx = a.refidx + chain * T;
    const double log_n = log((double)n), thr = 0.5 * (double)n, nd = (double)n;
    for (long long t = tid; t < T; t += B) refidx[t] = 0u;   // a fresh sampler object: arma::fill::zeros (conditionalSampler.h:96)
    __syncthreads();

    for (long long r = 0; r < a.runs_per_chain; ++r) {
        const long long run = chain * a.runs_per_chain + r;
        const int mode = a.modes[run];
        const double* ref = a.ref + (size_t)run * (size_t)T;
        double path = 0.0, lse_prev = 0.0;
        int resampled = 0, cur = 0;
        for (long long t = 0; t < T; ++t) {
            // ---- DoInit / DoMove over every slot (through the previous step's ancestors), then the pin of the reference slot
            if (t == 0 && tid == 0) refidx[0] = 0u;
            __syncthreads();
            const uint32_t pin = refidx[t];
            const double* src = xs[cur];
            double* dst = xs[cur ^ 1];
            LseAcc<0> acc; acc.clear();
            for (long long i = tid; i < n; i += B) {
                double z;
                if (a.zin) z = a.zin[((size_t)run * (size_t)T + (size_t)t) * (size_t)n + (size_t)i];
                else { double z1; philox_normal2(a.seed, t == 0 ? STREAM_INIT : STREAM_MOVE, (uint32_t)t, (uint64_t)i, (uint32_t)run, z, z1); }
                double x, l;
                if (t == 0) {
                    x = a.mp.phi * a.mp.x0 + (0.0 + a.mp.sd_evol * z);
                    l = LGSS::loglik(a.mp, a.mp.y[0], x);
                } else {
                    const double xp = resampled ? src[anc[i]] : src[i];
                    l = resampled ? -log_n : lw[i] - lse_prev;
                    x = a.mp.phi * xp + (0.0 + a.mp.sd_evol * z);
                    l += LGSS::loglik(a.mp, a.mp.y[t], x);
                }
                if ((uint32_t)i == pin) { x = ref[t]; l += LGSS::loglik(a.mp, a.mp.y[t], x); }   // pfWeight adds (cSMCexamples.cpp:203-210)
                if (t == 0) l -= log_n;
                dst[i] = x; lw[i] = l;
                acc.add(l, nullptr);
            }
            cur ^= 1;
            // ---- logNC increment, ESS (sampler.h:241, :413-417)
            acc.warp_reduce();
            if (lane == 0) { sh.red[0][warp] = acc.m; sh.red[1][warp] = acc.s; sh.red[2][warp] = acc.q; }
            __syncthreads();
            if (warp == 0) {
                LseAcc<0> b; b.clear();
                if (lane < B / 32) { b.m = sh.red[0][lane]; b.s = sh.red[1][lane]; b.q = sh.red[2][lane]; }
                b.warp_reduce();
                if (lane == 0) { sh.bc[0] = b.m + log(b.s); sh.bc[1] = (b.s * b.s) / b.q; }
            }
            __syncthreads();
            const double lse = sh.bc[0], ess = sh.bc[1];
            path += lse;
            lse_prev = lse;
            resampled = ess < thr ? 1 : 0;
            const long long wr = t == 0 ? 1 : t;                      // referenceTrajectoryIndices.at(T + 1), the sampler's own T
            const uint32_t kt1 = refidx[t == 0 ? 0 : t - 1];
            double* x = xs[cur];
            if (!resampled) {
                if (tid == 0 && wr < T) refidx[wr] = kt1;             // conditionalSampler.h:329-333
            } else {
                for (long long i = tid; i < n; i += B) { w[i] = exp(lw[i] - lse); W[i] = w[i]; }
                __syncthreads();
                if (mode == MULTINOMIAL) {                            // :385-414
                    const double tot = cta_scan(W, n, sh.scan_d);
                    for (long long i = tid; i < n; i += B)
                        anc[i] = i == 0 ? kt1 : (uint32_t)sample_cdf(W, n, tot, cs_uniform(a, run, t, 2 + i));
                    if (tid == 0 && wr < T) refidx[wr] = 0u;
                } else if (mode == STRATIFIED || mode == SYSTEMATIC) { // :416-556
                    cta_scan(W, n, sh.scan_d);
                    const double hi = W[kt1], lo = kt1 > 0 ? W[kt1 - 1] : 0.0, wk = w[kt1];
                    for (long long i = tid; i < n; i += B) {
                        const double up = (double)(i + 1) / nd, dn = (double)i / nd;
                        const double s = fmin(hi, up) - fmax(lo, dn);
                        P[i] = (s > 0.0 ? s : 0.0) / wk;
                    }
                    __syncthreads();
                    const double ptot = cta_scan(P, n, sh.scan_d);
                    if (tid == 0) {
                        const long long Kt = sample_cdf(P, n, ptot, cs_uniform(a, run, t, 0));
                        sh.bi[0] = Kt;
                        if (mode == SYSTEMATIC) {
                            const double vhi = W[Kt], vlo = Kt == 0 ? 0.0 : W[Kt - 1];
                            double V = vlo + (vhi - vlo) * cs_uniform(a, run, t, 1);
                            sh.bc[2] = nd * V - floor(nd * V);
                        }
                    }
                    __syncthreads();
                    const long long Kt = sh.bi[0];
                    const double V = sh.bc[2];
                    for (long long i = tid; i < n; i += B) {
                        if (i == Kt) { anc[i] = kt1; continue; }
                        double u = mode == SYSTEMATIC ? V + (double)i : cs_uniform(a, run, t, 2 + i) + (double)i;
                        u /= nd;
                        anc[i] = (uint32_t)first_above(W, n, u);
                    }
                } else {                                              // RESIDUAL :558-642
                    for (long long i = tid; i < n; i += B) {
                        const int e = (int)floor(nd * w[i]);
                        const int card = e > 0 ? e : 0;
                        Lx[i] = card;
                        W[i] = w[i] * nd - (double)card / nd;         // residual weight (kept quirk: floor / N)
                        anc[i] = 0u;
                    }
                    __syncthreads();
                    const long long l = cta_scan(Lx, n, sh.scan_i);   // inclusive: deterministic offspring of parents 0..i
                    const long long dk_hi = Lx[kt1], dk_lo = kt1 > 0 ? Lx[kt1 - 1] : 0;
                    const double rk = W[kt1];
                    const double prob1 = 1.0 / (nd * w[kt1]);
                    const double prob2 = rk / (nd * w[kt1] * (double)(n - l));
                    for (long long j = tid; j < n; j += B) {
                        double p = (j >= dk_lo && j < dk_hi) ? prob1 : 0.0;
                        if (l < n ? j >= l : j == n - 1) p = prob2;
                        P[j] = p;
                        // deterministic offspring: slot j < l belongs to the first parent whose running count exceeds j
                        if (j < l) {
                            long long lo2 = 0, hi2 = n - 1;
                            while (lo2 < hi2) { long long mid = (lo2 + hi2) >> 1; if ((long long)Lx[mid] > j) hi2 = mid; else lo2 = mid + 1; }
                            anc[j] = (uint32_t)lo2;
                        }
                    }
                    __syncthreads();
                    const double ptot = cta_scan(P, n, sh.scan_d);
                    const double rtot = cta_scan(W, n, sh.scan_d);    // running sums of the residual weights
                    if (tid == 0) {
                        const long long Kt = sample_cdf(P, n, ptot, cs_uniform(a, run, t, 0));
                        sh.bi[0] = Kt;
                        if (wr < T) refidx[wr] = (uint32_t)Kt;        // :627
                        if (dk_hi == dk_lo) anc[Kt] = kt1;            // :629-633 (the "not in D" branch never fires)
                    }
                    __syncthreads();
                    const long long Kt = sh.bi[0];
                    if (l < n)
                        for (long long k = l + tid; k < n; k += B)
                            if (k != Kt) anc[k] = (uint32_t)sample_cdf(W, n, rtot, cs_uniform(a, run, t, 2 + k));
                }
                __syncthreads();
            }
            // ---- AL history row (post-resampling values, ancestors, normalised log-weights)
            if (a.anc_out || a.val_out || a.lw_out) {
                const size_t row = ((size_t)run * (size_t)T + (size_t)t) * (size_t)n;
                for (long long i = tid; i < n; i += B) {
                    const uint32_t p = resampled ? anc[i] : (uint32_t)i;
                    if (a.anc_out) a.anc_out[row + i] = p;
                    if (a.val_out) a.val_out[row + i] = x[p];
                    if (a.lw_out) a.lw_out[row + i] = resampled ? -log_n : lw[i] - lse;
                }
            }
            if (a.res_out && tid == 0) a.res_out[run * T + t] = resampled;
            __syncthreads();
        }
        if (tid == 0) a.lognc[run] = path;
        for (long long t = tid; t < T; t += B) a.refidx_out[run * T + t] = refidx[t];
        __syncthreads();
    }
}

// sampler::GetALineInd(n) for every n and GetALineSpace (sampler.h:445-476): one thread walks one particle's line backwards.
__global__ void __launch_bounds__(256) k_ancestor_lines(const uint32_t* anc, const double* values, long long T, long long n, uint32_t* lines,
                                                        double* line_values) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        uint32_t p = anc[(size_t)(T - 1) * (size_t)n + (size_t)i];
        lines[(size_t)(T - 1) * (size_t)n + (size_t)i] = (uint32_t)i;
        if (line_values) line_values[(size_t)(T - 1) * (size_t)n + (size_t)i] = values[(size_t)(T - 1) * (size_t)n + (size_t)i];
        for (long long t = T - 2; t >= 0; --t) {
            p = anc[(size_t)t * (size_t)n + p];
            lines[(size_t)t * (size_t)n + (size_t)i] = p;
            if (line_values) line_values[(size_t)t * (size_t)n + (size_t)i] = values[(size_t)t * (size_t)n + p];
        }
    }
}

LGSS::Params lgss_params(const double* y_dev, double phi, double var_evol, double x0, double var_obs) {
    const double sd_obs = std::sqrt(var_obs);
    return LGSS::Params{y_dev, phi, std::sqrt(var_evol), x0, sd_obs, std::log(sd_obs)};
}

void run_chains(const double* y_host, long T, long particles, double phi, double var_evol, double x0, double var_obs, long n_chains,
                long runs_per_chain, const int* modes_host, const double* ref_host, uint64_t seed, const double* noise_host,
                const double* unif_host, double* lognc, uint32_t* refidx_out, uint32_t* anc_out, double* val_out, double* lw_out, int* res_out,
                const std::function<void()>& overlap = nullptr) {
    if (!y_host || T < 2 || particles < 2 || n_chains < 1 || runs_per_chain < 1 || !modes_host || !ref_host || !lognc)
        throw std::invalid_argument("conditionalBPF: need data with T >= 2, particles >= 2, chains, modes, reference trajectories and an output");
    if (particles >= (1LL << 31)) throw std::invalid_argument("conditionalBPF: particles must be < 2^31");
    const long runs = n_chains * runs_per_chain;
    for (long r = 0; r < runs; ++r) if (modes_host[r] < 0 || modes_host[r] > 3) throw std::invalid_argument("conditionalBPF: unknown resampling mode");
    require_device();
    const long long n = particles;
    DevBuf y(sizeof(double) * T), modes(sizeof(int) * runs), ref(sizeof(double) * runs * T);
    SMCB_CUDA(cudaMemcpy(y.p, y_host, sizeof(double) * T, cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(modes.p, modes_host, sizeof(int) * runs, cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(ref.p, ref_host, sizeof(double) * runs * T, cudaMemcpyHostToDevice));
    const size_t cn = (size_t)n_chains * (size_t)n;
    DevBuf dbl(sizeof(double) * 6 * cn), Lb(sizeof(int) * cn), ancb(sizeof(uint32_t) * cn), ridx(sizeof(uint32_t) * n_chains * T);
    DevBuf d_nc(sizeof(double) * runs), d_ro(sizeof(uint32_t) * runs * T), zin, uin, d_anc, d_val, d_lw, d_res;
    const size_t hist = (size_t)runs * (size_t)T * (size_t)n;
    if (noise_host) { zin = DevBuf(sizeof(double) * hist); SMCB_CUDA(cudaMemcpy(zin.p, noise_host, sizeof(double) * hist, cudaMemcpyHostToDevice)); }
    if (unif_host) {
        const size_t nu = (size_t)runs * (size_t)T * (size_t)(n + 2);
        uin = DevBuf(sizeof(double) * nu); SMCB_CUDA(cudaMemcpy(uin.p, unif_host, sizeof(double) * nu, cudaMemcpyHostToDevice));
    }
    if (anc_out) d_anc = DevBuf(sizeof(uint32_t) * hist);
    if (val_out) d_val = DevBuf(sizeof(double) * hist);
    if (lw_out) d_lw = DevBuf(sizeof(double) * hist);
    if (res_out) d_res = DevBuf(sizeof(int) * runs * T);
    CsmcArgs a{};
    a.mp = lgss_params(y.as<double>(), phi, var_evol, x0, var_obs);
    a.T = T; a.n = n; a.n_chains = n_chains; a.runs_per_chain = runs_per_chain;
    a.modes = modes.as<int>(); a.ref = ref.as<double>(); a.seed = seed; a.zin = zin.as<double>(); a.uin = uin.as<double>();
    double* d = dbl.as<double>();
    a.xa = d; a.xb = d + cn; a.lw = d + 2 * cn; a.w = d + 3 * cn; a.W = d + 4 * cn; a.P = d + 5 * cn;
    a.L = Lb.as<int>(); a.anc = ancb.as<uint32_t>(); a.refidx = ridx.as<uint32_t>();
    a.lognc = d_nc.as<double>(); a.refidx_out = d_ro.as<uint32_t>();
    a.anc_out = d_anc.as<uint32_t>(); a.val_out = d_val.as<double>(); a.lw_out = d_lw.as<double>(); a.res_out = d_res.as<int>();
    // The chains run on their own non-blocking stream: independent work handed in by the caller (the plain filters of
    // compareNCestimates, which use the legacy stream) overlaps with them instead of queueing behind one long kernel.
    cudaStream_t cs = nullptr;
    SMCB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    k_csmc_chain<<<(unsigned)n_chains, CS_THREADS, 0, cs>>>(a);
    cudaError_t le = cudaGetLastError();
    try {
        if (le == cudaSuccess && overlap) overlap();
    } catch (...) {
        cudaStreamSynchronize(cs); cudaStreamDestroy(cs);
        throw;
    }
    cudaError_t se = cudaStreamSynchronize(cs);
    cudaStreamDestroy(cs);
    SMCB_CUDA(le);
    SMCB_CUDA(se);
    SMCB_CUDA(cudaMemcpy(lognc, d_nc.p, sizeof(double) * runs, cudaMemcpyDeviceToHost));
    if (refidx_out) SMCB_CUDA(cudaMemcpy(refidx_out, d_ro.p, sizeof(uint32_t) * runs * T, cudaMemcpyDeviceToHost));
    if (anc_out) SMCB_CUDA(cudaMemcpy(anc_out, d_anc.p, sizeof(uint32_t) * hist, cudaMemcpyDeviceToHost));
    if (val_out) SMCB_CUDA(cudaMemcpy(val_out, d_val.p, sizeof(double) * hist, cudaMemcpyDeviceToHost));
    if (lw_out) SMCB_CUDA(cudaMemcpy(lw_out, d_lw.p, sizeof(double) * hist, cudaMemcpyDeviceToHost));
    if (res_out) SMCB_CUDA(cudaMemcpy(res_out, d_res.p, sizeof(int) * runs * T, cudaMemcpyDeviceToHost));
}

// Plain bootstrap filter on the LGSS model through the general engine: Initialise(); IterateUntil(T-1); GetLogNCPath().
struct PlainLgss {
    Sampler<LGSS> s;
    DevBuf y, noise, unif;
    long T, n;
    PlainLgss(const double* y_host, long T_, long particles, uint64_t seed) : s(particles, T_, seed, 0), y(sizeof(double) * T_), T(T_), n(particles) {
        SMCB_CUDA(cudaMemcpy(y.p, y_host, sizeof(double) * T, cudaMemcpyHostToDevice));
    }
    bool allow_graph = false;   // set by callers that run many filters of one shape
    double run(int mode, double phi, double var_evol, double x0, double var_obs, uint64_t seed, const double* noise_host, const double* unif_host) {
        s.SetResampleParams(mode, 0.5);                               // cSMCexamples.cpp:82
        s.SetParams(lgss_params(y.as<double>(), phi, var_evol, x0, var_obs));
        s.SetSeed(seed);
        const size_t per_noise = (size_t)n * (size_t)T;
        const size_t per_unif = mode == SMCB200_SYSTEMATIC ? (size_t)T : mode == SMCB200_STRATIFIED ? (size_t)T * (size_t)(n + 1) : per_noise;
        const double* zn = nullptr; const double* un = nullptr;
        if (noise_host) {
            if (!noise.p) noise = DevBuf(sizeof(double) * per_noise);
            SMCB_CUDA(cudaMemcpy(noise.p, noise_host, sizeof(double) * per_noise, cudaMemcpyHostToDevice)); zn = noise.as<double>();
        }
        if (unif_host) {
            unif = DevBuf(sizeof(double) * per_unif);
            SMCB_CUDA(cudaMemcpy(unif.p, unif_host, sizeof(double) * per_unif, cudaMemcpyHostToDevice)); un = unif.as<double>();
        }
        s.SetInjectedNoise(zn, mode == SMCB200_SYSTEMATIC ? un : nullptr, mode == SMCB200_STRATIFIED ? un : nullptr,
                           (mode == SMCB200_MULTINOMIAL || mode == SMCB200_RESIDUAL) ? un : nullptr);
        s.RunFilter(T, /*use_graph=*/allow_graph && !noise_host && !unif_host);
        return s.ReadCtrl().lognc_path;
    }
};

}  // namespace

extern "C" int smcb200_conditionalBPF(const double* data_host, long T, long particles, double phi, double varStateEvol, double stateInit,
                                      double varObs, long n_chains, long runs_per_chain, const int* modes_host,
                                      const double* reference_host, uint64_t seed, const double* noise_host, const double* uniforms_host,
                                      double* lognc_host, uint32_t* refidx_host, uint32_t* ancestors_host, double* values_host,
                                      double* logw_host, int* resampled_host) {
    return guarded([&]() -> int {
        run_chains(data_host, T, particles, phi, varStateEvol, stateInit, varObs, n_chains, runs_per_chain, modes_host, reference_host, seed,
                   noise_host, uniforms_host, lognc_host, refidx_host, ancestors_host, values_host, logw_host, resampled_host);
        return SMCB200_OK;
    });
}

extern "C" int smcb200_lgssBPF(const double* data_host, long T, long particles, double phi, double varStateEvol, double stateInit,
                               double varObs, int resample_mode, uint64_t seed, const double* noise_host, const double* uniforms_host,
                               double* lognc_host) {
    return guarded([&]() -> int {
        if (!data_host || T < 1 || particles < 1 || !lognc_host) throw std::invalid_argument("lgssBPF: data, T >= 1, particles >= 1 and an output required");
        if (resample_mode < 0 || resample_mode > 3) throw std::invalid_argument("lgssBPF: unknown resampling mode");
        require_device();
        PlainLgss f(data_host, T, particles, seed);
        *lognc_host = f.run(resample_mode, phi, varStateEvol, stateInit, varObs, seed, noise_host, uniforms_host);
        return SMCB200_OK;
    });
}

extern "C" int smcb200_compareNCestimates(const double* data_host, long T, long particles, int simNum, double phi, double varStateEvol,
                                          double stateInit, double varObs, const double* referenceTraj_host, uint64_t seed,
                                          int independent_sims, double* smcOut_host, double* csmcOut_host) {
    return guarded([&]() -> int {
        if (!data_host || T < 2 || particles < 2 || simNum < 1 || !referenceTraj_host || !smcOut_host || !csmcOut_host)
            throw std::invalid_argument("compareNCestimates: data (T >= 2), particles >= 2, simNum >= 1, referenceTraj and both outputs required");
        require_device();
        const int order[4] = {SMCB200_MULTINOMIAL, SMCB200_RESIDUAL, SMCB200_STRATIFIED, SMCB200_SYSTEMATIC};   // cSMCexamples.cpp:82-125
        // conditional filters: run (i, c) uses referenceTraj.col(i); one chain reproduces the reference's single sampler object
        const long runs = 4L * simNum;
        std::vector<int> modes(runs);
        std::vector<double> ref((size_t)runs * T), nc(runs);
        for (int i = 0; i < simNum; ++i)
            for (int c = 0; c < 4; ++c) {
                modes[4 * i + c] = order[c];
                for (long t = 0; t < T; ++t) ref[(size_t)(4 * i + c) * T + t] = referenceTraj_host[(size_t)i * T + t];
            }
        // plain bootstrap filters (column c of smcOut is order[c]) run while the chains are in flight
        auto plain = [&]() {
            PlainLgss f(data_host, T, particles, seed);
            f.allow_graph = simNum >= 16;
            for (int c = 0; c < 4; ++c)          // scheme-major: one captured graph per scheme serves all simulations
                for (int i = 0; i < simNum; ++i)
                    smcOut_host[(size_t)c * simNum + i] =
                        f.run(order[c], phi, varStateEvol, stateInit, varObs, seed + 0x9E3779B97F4A7C15ull * (uint64_t)(4 * i + c + 1), nullptr, nullptr);
        };
        // a single chain occupies one SM for the whole call: the plain filters overlap with it; many chains fill the GPU, so the
        // plain filters follow them instead of competing for the same SMs
        const long n_chains = independent_sims ? simNum : 1;
        const bool overlap = n_chains <= 8;
        run_chains(data_host, T, particles, phi, varStateEvol, stateInit, varObs, n_chains, runs / n_chains, modes.data(), ref.data(),
                   seed ^ 0xC5A3C5A3C5A3C5A3ull, nullptr, nullptr, nc.data(), nullptr, nullptr, nullptr, nullptr, nullptr,
                   overlap ? std::function<void()>(plain) : std::function<void()>());
        if (!overlap) plain();
        for (int i = 0; i < simNum; ++i)
            for (int c = 0; c < 4; ++c) csmcOut_host[(size_t)c * simNum + i] = nc[4 * i + c];
        return SMCB200_OK;
    });
}

extern "C" int smcb200_ancestor_lines(const uint32_t* ancestors_host, long T, long particles, const double* values_host,
                                      uint32_t* lines_host, double* line_values_host) {
    return guarded([&]() -> int {
        if (!ancestors_host || T < 1 || particles < 1 || !lines_host) throw std::invalid_argument("ancestor_lines: ancestors [T][N] and lines output required");
        if ((values_host == nullptr) != (line_values_host == nullptr)) throw std::invalid_argument("ancestor_lines: values and line_values go together");
        const size_t tn = (size_t)T * (size_t)particles;
        for (size_t k = 0; k < tn; ++k) if (ancestors_host[k] >= (uint32_t)particles) throw std::invalid_argument("ancestor_lines: ancestor index out of range");
        require_device();
        DevBuf anc(sizeof(uint32_t) * tn), lines(sizeof(uint32_t) * tn), val, lv;
        SMCB_CUDA(cudaMemcpy(anc.p, ancestors_host, sizeof(uint32_t) * tn, cudaMemcpyHostToDevice));
        if (values_host) {
            val = DevBuf(sizeof(double) * tn); lv = DevBuf(sizeof(double) * tn);
            SMCB_CUDA(cudaMemcpy(val.p, values_host, sizeof(double) * tn, cudaMemcpyHostToDevice));
        }
        int dev = 0, sms = 0;
        SMCB_CUDA(cudaGetDevice(&dev));
        SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const long long want = ((long long)particles + 255) / 256;
        k_ancestor_lines<<<(unsigned)(want < 8LL * sms ? want : 8LL * sms), 256>>>(anc.as<uint32_t>(), val.as<double>(), T, particles, lines.as<uint32_t>(),
                                                                               lv.as<double>());
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaDeviceSynchronize());
        SMCB_CUDA(cudaMemcpy(lines_host, lines.p, sizeof(uint32_t) * tn, cudaMemcpyDeviceToHost));
        if (line_values_host) SMCB_CUDA(cudaMemcpy(line_values_host, lv.p, sizeof(double) * tn, cudaMemcpyDeviceToHost));
        return SMCB200_OK;
    });
}
