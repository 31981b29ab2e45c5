// Conditional SMC on the device for ANY model written as device functors (smc::conditionalSampler<Space,Params>, reference
// inst/include/conditionalSampler.h:39-155 interface, :166-235 Initialise, :303-361 IterateEss, :377-655 conditionalResample,
// :657-661 MoveReferenceParticle; moveset.h:170-180 DoConditionalMove).  Model contract (static, __device__):
//   D, NZ_INIT, NZ_MOVE, struct Params
//   init(x, lw, params, z)        pfInitialise      move(t, x, lw, params, z)   pfMove
//   weight(t, x, lw, params)      pfWeight: the log-weight increment of a particle whose value has just been set
//
// Shape of the work.  The reference workload is many small filters (N ~ 10^3 particles), each strictly sequential in time.  One
// CTA runs one CHAIN of conditional filters start to finish -- every time step, its reductions, scans and searches stay inside
// the CTA (no launches, no grid synchronisation) -- and chains run side by side on the SMs.  A chain is the unit of sequential
// dependence: the reference keeps ONE sampler object for all runs, and its referenceTrajectoryIndices carry over from run to run
// (SURVEY A.11), so runs of one chain execute in order on shared index state.  Every conditional resampling scheme is "scan, then
// one independent binary search per offspring" (the reference walks the cumulative weights sequentially, O(N^2) worst case
// through arma::find(tail...)).  Rcpp::sample (third party, absent) is defined as sequential inverse CDF, one uniform per draw.
#pragma once
#include <array>
#include <stdexcept>
#include <vector>

#include "common.cuh"
#include "philox.cuh"

namespace smcb {

constexpr int CS_THREADS = 512;

template <class M>
struct CsmcArgs {
    typename M::Params mp;
    long long T, n;
    long long n_chains, runs_per_chain;
    const int* modes;        // [n_chains * runs_per_chain]
    const double* ref;       // [n_chains * runs_per_chain][T][D] reference trajectories
    unsigned long long seed;
    const double* zin;       // [runs][T][n][NZ] injected standard normals, or null
    const double* uin;       // [runs][T][n + 2] or null: {K_t draw, systematic V, offspring i -> 2 + i}
    long long t_cap;         // length of the reference trajectory (the slot-index array the steps write ahead into)
    double thr;              // absolute resampling threshold (ESS below it resamples)
    long long run0;          // index of the first run (Philox draws are keyed on the run index)
    int keep_refidx;         // 0: a fresh sampler object (reference-slot indices start at zero); 1: they carry over from earlier launches
    // per-chain scratch (states: [chain][D][n])
    double* xa; double* xb; double* lw; double* w; double* W; double* P;
    int* L; uint32_t* anc; uint32_t* refidx;
    // outputs
    double* lognc;           // [runs]
    uint32_t* refidx_out;    // [runs][T]
    uint32_t* anc_out;       // [runs][T][n] or null  (AL history)
    double* val_out; double* lw_out;   // [runs][T][D][n], [runs][T][n]
    int* res_out;            // [runs][T] or null
    double* ess_out;         // [runs][T] or null
};

struct CsmcShared {
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

template <class M>
__device__ __forceinline__ double cs_uniform(const CsmcArgs<M>& a, long long run, long long t, long long slot) {
    if (a.uin) return a.uin[((size_t)run * (size_t)a.T + (size_t)t) * (size_t)(a.n + 2) + (size_t)slot];
    return u64_to_unit(philox_draw(a.seed, STREAM_RESAMPLE, (uint32_t)t, (uint64_t)slot, (uint32_t)(a.run0 + run)).a);
}

template <class M>
__global__ void __launch_bounds__(CS_THREADS) k_csmc_chain(CsmcArgs<M> a) {
    constexpr int D = M::D;
    constexpr int NZ = (M::NZ_INIT > M::NZ_MOVE ? M::NZ_INIT : M::NZ_MOVE) > 0 ? (M::NZ_INIT > M::NZ_MOVE ? M::NZ_INIT : M::NZ_MOVE) : 1;
    __shared__ CsmcShared sh;
    const int tid = threadIdx.x, B = CS_THREADS, lane = tid & 31, warp = tid >> 5;
    const long long n = a.n, T = a.T, chain = blockIdx.x;
    double* xs[2] = {a.xa + chain * n * D, a.xb + chain * n * D};
    double* lw = a.lw + chain * n;
    double* w = a.w + chain * n;
    double* W = a.W + chain * n;
    double* P = a.P + chain * n;
    int* Lx = a.L + chain * n;
    uint32_t* anc = a.anc + chain * n;
    uint32_t* refidx = a.refidx + chain * T;
    const double log_n = log((double)n), thr = a.thr, nd = (double)n;
    if (!a.keep_refidx) for (long long t = tid; t < T; t += B) refidx[t] = 0u;   // a fresh sampler object: arma::fill::zeros (conditionalSampler.h:96)
    __syncthreads();

    for (long long r = 0; r < a.runs_per_chain; ++r) {
        const long long run = chain * a.runs_per_chain + r;
        const int mode = a.modes[run];
        const double* ref = a.ref + (size_t)run * (size_t)T * (size_t)D;
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
                double z[NZ];
                const int nz = t == 0 ? M::NZ_INIT : M::NZ_MOVE;
                for (int k = 0; k < nz; ++k) {
                    if (a.zin) z[k] = a.zin[(((size_t)run * (size_t)T + (size_t)t) * (size_t)n + (size_t)i) * (size_t)nz + (size_t)k];
                    else { double z1; philox_normal2(a.seed, t == 0 ? STREAM_INIT : STREAM_MOVE, (uint32_t)t, (uint64_t)i, (uint32_t)((a.run0 + run) * nz + k), z[k], z1); }
                }
                double x[D], l;
                if (t == 0) {
                    M::init(x, l, a.mp, z);                           // pfInitialise (moveset.h:133-138)
                } else {
                    const long long p = resampled ? (long long)anc[i] : i;
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = src[(size_t)d * n + p];
                    l = resampled ? -log_n : lw[i] - lse_prev;
                    M::move(t, x, l, a.mp, z);                        // pfMove (moveset.h:162-168)
                }
                if ((uint32_t)i == pin) {                             // DoConditionalMove: pin the reference value, pfWeight adds (moveset.h:170-180)
#pragma unroll
                    for (int d = 0; d < D; ++d) x[d] = ref[(size_t)t * D + d];
                    M::weight(t, x, l, a.mp);
                }
                if (t == 0) l -= log_n;
#pragma unroll
                for (int d = 0; d < D; ++d) dst[(size_t)d * n + i] = x[d];
                lw[i] = l;
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
                if (tid == 0 && wr < a.t_cap) refidx[wr] = kt1;             // conditionalSampler.h:329-333
            } else {
                for (long long i = tid; i < n; i += B) { w[i] = exp(lw[i] - lse); W[i] = w[i]; }
                __syncthreads();
                if (mode == MULTINOMIAL) {                            // :385-414
                    const double tot = cta_scan(W, n, sh.scan_d);
                    for (long long i = tid; i < n; i += B)
                        anc[i] = i == 0 ? kt1 : (uint32_t)sample_cdf(W, n, tot, cs_uniform(a, run, t, 2 + i));
                    if (tid == 0 && wr < a.t_cap) refidx[wr] = 0u;
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
                        if (wr < a.t_cap) refidx[wr] = (uint32_t)Kt;        // :627
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
                    if (a.val_out) {
#pragma unroll
                        for (int d = 0; d < D; ++d) a.val_out[row * D + (size_t)d * n + i] = x[(size_t)d * n + p];
                    }
                    if (a.lw_out) a.lw_out[row + i] = resampled ? -log_n : lw[i] - lse;
                }
            }
            if (a.res_out && tid == 0) a.res_out[run * T + t] = resampled;
            if (a.ess_out && tid == 0) a.ess_out[run * T + t] = ess;
            __syncthreads();
        }
        if (tid == 0) a.lognc[run] = path;
        for (long long t = tid; t < T; t += B) a.refidx_out[run * T + t] = refidx[t];
        __syncthreads();
    }
}



// ---- host side: smc::conditionalSampler<Model> (conditionalSampler.h:39-155) --------------------------------------------------
// One object = one chain: its reference-slot indices carry over from run to run, as the reference's do.  Execution
// is deferred: Initialise() / Iterate() / IterateUntil() move the evolution time, and the first getter launches ONE kernel that
// runs the filter from time 0 to the current time (the draws are counter-based, so running a longer prefix later reproduces the
// shorter one).  IterateEss() needs the step's ESS and therefore runs the filter up to that step each time it is called.
template <class M>
class ConditionalSampler {
public:
    static constexpr int D = M::D;
    using Params = typename M::Params;
    using Space = std::array<double, D>;

    ConditionalSampler(long lSize, int htHistoryMode, const std::vector<Space>& referenceTrajectoryInit, unsigned long long seed = 1, cudaStream_t stream = 0)
        : n_(lSize), hist_(htHistoryMode), ref_(referenceTrajectoryInit), seed_(seed), stream_(stream) {
        if (lSize < 2 || lSize >= (1LL << 31)) throw std::invalid_argument("conditionalSampler: particle count must be in [2, 2^31)");
        if (ref_.size() < 2) throw std::invalid_argument("conditionalSampler: the reference trajectory needs at least two time points");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) throw std::runtime_error("conditionalSampler: no CUDA device (there is no CPU fallback)");
        maxT_ = (long)ref_.size();
        const size_t n = (size_t)n_, T = (size_t)maxT_;
        alloc(&d_state_, sizeof(double) * (2 * D + 4) * n);
        alloc(&d_L_, sizeof(int) * n); alloc(&d_anc_, sizeof(uint32_t) * n); alloc(&d_refidx_, sizeof(uint32_t) * T);
        alloc(&d_ref_, sizeof(double) * T * D); alloc(&d_lognc_, sizeof(double)); alloc(&d_ro_, sizeof(uint32_t) * T);
        alloc(&d_res_, sizeof(int) * T); alloc(&d_ess_, sizeof(double) * T); alloc(&d_mode_, sizeof(int));
        if (hist_ != HISTORY_NONE) { alloc(&d_val_, sizeof(double) * T * D * n); alloc(&d_lw_, sizeof(double) * T * n); }
        if (hist_ == HISTORY_AL) alloc(&d_hanc_, sizeof(uint32_t) * T * n);
        SMCB_CUDA(cudaMemsetAsync(d_refidx_, 0, sizeof(uint32_t) * T, stream_));   // arma::fill::zeros (conditionalSampler.h:96)
        SetResampleParams(MULTINOMIAL, 0.5);                                        // sampler.h:259-267 defaults
    }
    ~ConditionalSampler() { for (void* p : allocs_) cudaFree(p); }
    ConditionalSampler(const ConditionalSampler&) = delete;
    ConditionalSampler& operator=(const ConditionalSampler&) = delete;

    void SetResampleParams(int mode, double threshold) {   // sampler.h:885-893
        if (mode < MULTINOMIAL || mode > SYSTEMATIC) throw std::invalid_argument("unknown resampling mode");
        mode_ = mode; thr_ = threshold < 1 ? threshold * (double)n_ : threshold; dirty_ = true;
    }
    void SetAlgParam(const Params& p) { prm_ = p; dirty_ = true; }
    void SetSeed(unsigned long long seed) { seed_ = seed; dirty_ = true; }
    std::vector<Space> GetReferenceTrajectory() const { return ref_; }
    Space GetReferenceValue(long t) const { return ref_.at((size_t)t); }
    void SetReferenceTrajectory(const std::vector<Space>& r) {
        if ((long)r.size() != maxT_) throw std::invalid_argument("conditionalSampler: the reference trajectory keeps its length");
        ref_ = r; dirty_ = true;
    }
    void SetReferenceValue(const Space& v, long t) { ref_.at((size_t)t) = v; dirty_ = true; }

    void Initialise() { if (started_) ++run_; started_ = true; T_ = 0; dirty_ = true; }   // conditionalSampler.h:162-227
    void Initialise(const std::vector<Space>& r) { SetReferenceTrajectory(r); Initialise(); }
    void Iterate() {                                                                     // :303-361
        if (!started_) throw std::runtime_error("conditionalSampler: Initialise() first");
        if (T_ + 1 >= maxT_) throw std::out_of_range("conditionalSampler: no reference value beyond the end of the trajectory");
        ++T_; dirty_ = true;
    }
    double IterateEss() { Iterate(); flush(); return host_ess_.at((size_t)T_); }
    void IterateUntil(long lTerminate) { while (T_ < lTerminate) Iterate(); }

    long GetNumber() const { return n_; }
    long GetTime() const { return T_; }
    double GetLogNCPath() { flush(); return lognc_; }
    double GetESS(long t) { flush(); return host_ess_.at((size_t)t); }
    int GetResampled() { flush(); return host_res_.at((size_t)T_); }
    int GetHistoryResampled(long t) { flush(); return host_res_.at((size_t)t); }
    long GetReferenceValueIndex(long lTime) { flush(); return (long)host_refidx_.at((size_t)lTime); }
    long GetHistoryLength() { return hist_ == HISTORY_NONE ? 0 : T_ + 1; }
    // generation t of the history (post-resampling population, history.h:63-120): values [D][N], normalised log-weights, and with
    // HistoryType::AL the ancestor indices of the generation's resampling (identity when it did not resample)
    std::vector<double> GetHistoryValues(long t) { need_hist(t, HISTORY_RAM); return copy_out<double>(d_val_ + (size_t)t * D * n_, (size_t)D * n_); }
    std::vector<double> GetHistoryLogWeights(long t) { need_hist(t, HISTORY_RAM); return copy_out<double>(d_lw_ + (size_t)t * n_, (size_t)n_); }
    std::vector<uint32_t> GetHistoryAncestors(long t) { need_hist(t, HISTORY_AL); return copy_out<uint32_t>(d_hanc_ + (size_t)t * n_, (size_t)n_); }
    // ancestral line of final particle i (sampler.h:445-460): slot indices at times 0 .. GetTime()
    std::vector<uint32_t> GetALineInd(long i) {
        need_hist(T_, HISTORY_AL);
        std::vector<uint32_t> anc = copy_out<uint32_t>(d_hanc_, (size_t)(T_ + 1) * n_), line((size_t)T_ + 1);
        uint32_t p = anc[(size_t)T_ * n_ + (size_t)i];   // the reference's recursion, as written (sampler.h:450-458)
        line[(size_t)T_] = (uint32_t)i;
        for (long t = T_ - 1; t >= 0; --t) { p = anc[(size_t)t * n_ + p]; line[(size_t)t] = p; }
        return line;
    }
    // members of the base class that the reference disables for conditional samplers (conditionalSampler.h:146-155)
    template <class A> void SetAdaptMethods(A*) { throw std::logic_error("Adaptation methods not currently supported for conditional sampler class."); }
    void SetMcmcRepeats(int) { throw std::logic_error("MCMC moves not currently supported for conditional sampler class."); }
    int GetAccepted() const { throw std::logic_error("MCMC moves not currently supported for conditional sampler class."); }
    int GetMcmcRepeats() const { throw std::logic_error("MCMC moves not currently supported for conditional sampler class."); }

private:
    void alloc(void* pp, size_t bytes) {
        void* p = nullptr;
        SMCB_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        allocs_.push_back(p);
        *reinterpret_cast<void**>(pp) = p;
    }
    template <class V> std::vector<V> copy_out(const V* dev, size_t count) {
        std::vector<V> h(count);
        SMCB_CUDA(cudaMemcpyAsync(h.data(), dev, sizeof(V) * count, cudaMemcpyDeviceToHost, stream_));
        SMCB_CUDA(cudaStreamSynchronize(stream_));
        return h;
    }
    void need_hist(long t, int level) {
        if (hist_ == HISTORY_NONE || (level == HISTORY_AL && hist_ != HISTORY_AL)) throw std::logic_error("conditionalSampler: this history mode does not keep that");
        if (t < 0 || t > T_) throw std::out_of_range("conditionalSampler: generation out of range");
        flush();
    }
    void flush() {
        if (!started_) throw std::runtime_error("conditionalSampler: Initialise() first");
        if (!dirty_) return;
        const size_t n = (size_t)n_;
        std::vector<double> flat((size_t)maxT_ * D);
        for (long t = 0; t < maxT_; ++t) for (int d = 0; d < D; ++d) flat[(size_t)t * D + d] = ref_[(size_t)t][(size_t)d];
        SMCB_CUDA(cudaMemcpyAsync(d_ref_, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice, stream_));
        SMCB_CUDA(cudaMemcpyAsync(d_mode_, &mode_, sizeof(int), cudaMemcpyHostToDevice, stream_));
        // the kernel writes the slot index of time t+1 while it works on time t: runs of one object see what earlier runs left,
        // but re-running THIS run from time 0 must start from the state the run began with
        if (run_ != saved_run_) {
            saved_refidx_ = copy_out<uint32_t>(d_refidx_, (size_t)maxT_);
            saved_run_ = run_;
        } else {
            SMCB_CUDA(cudaMemcpyAsync(d_refidx_, saved_refidx_.data(), sizeof(uint32_t) * (size_t)maxT_, cudaMemcpyHostToDevice, stream_));
        }
        CsmcArgs<M> a{};
        a.mp = prm_; a.T = T_ + 1; a.n = n_; a.n_chains = 1; a.runs_per_chain = 1; a.modes = d_mode_; a.ref = d_ref_; a.seed = seed_;
        a.t_cap = maxT_; a.thr = thr_; a.run0 = run_; a.keep_refidx = 1;
        a.xa = d_state_; a.xb = d_state_ + (size_t)D * n; a.lw = d_state_ + 2 * (size_t)D * n; a.w = a.lw + n; a.W = a.w + n; a.P = a.W + n;
        a.L = d_L_; a.anc = d_anc_; a.refidx = d_refidx_; a.lognc = d_lognc_; a.refidx_out = d_ro_;
        a.anc_out = d_hanc_; a.val_out = d_val_; a.lw_out = d_lw_; a.res_out = d_res_; a.ess_out = d_ess_;
        k_csmc_chain<M><<<1, CS_THREADS, 0, stream_>>>(a);
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpyAsync(&lognc_, d_lognc_, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        host_ess_ = copy_out<double>(d_ess_, (size_t)T_ + 1);
        host_res_ = copy_out<int>(d_res_, (size_t)T_ + 1);
        host_refidx_ = copy_out<uint32_t>(d_refidx_, (size_t)maxT_);
        dirty_ = false;
    }

    long n_, maxT_ = 0, T_ = 0;
    int hist_, mode_ = MULTINOMIAL;
    double thr_ = 0.0, lognc_ = 0.0;
    std::vector<Space> ref_;
    unsigned long long seed_;
    cudaStream_t stream_;
    Params prm_{};
    long long run_ = 0, saved_run_ = -1;
    bool started_ = false, dirty_ = true;
    double* d_state_ = nullptr; int* d_L_ = nullptr; uint32_t* d_anc_ = nullptr; uint32_t* d_refidx_ = nullptr; double* d_ref_ = nullptr;
    double* d_lognc_ = nullptr; uint32_t* d_ro_ = nullptr; int* d_res_ = nullptr; double* d_ess_ = nullptr; int* d_mode_ = nullptr;
    double* d_val_ = nullptr; double* d_lw_ = nullptr; uint32_t* d_hanc_ = nullptr;
    std::vector<double> host_ess_;
    std::vector<int> host_res_;
    std::vector<uint32_t> host_refidx_, saved_refidx_;
    std::vector<void*> allocs_;
};

}  // namespace smcb
