// Generic SMC engine behind the reference's plug-in surface: what smc::sampler<Space,Params> + smc::moveset<Space,Params> +
// smc::adaptMethods<Space,Params> (inst/include/sampler.h:133-237, moveset.h:73-96, adaptMethods.h:45-51) offer to a user
// model, for ANY model that can be written as a set of __host__ __device__ functors over a fixed-size state:
//
//   struct Model {
//       static constexpr int D = ...;                 // doubles per particle (Space); stored as D structure-of-arrays columns
//       struct Params { ... };                        // algParams: plain data, handed to every functor by value
//       SMCB_HD static void init(double (&x)[D], double& logw, const Params&, smcb::Rng&);           // pfInitialise
//       SMCB_HD static void move(long t, double (&x)[D], double& logw, const Params&, smcb::Rng&);   // pfMove
//       SMCB_HD static bool mcmc(long t, double (&x)[D], double& logw, const Params&, smcb::Rng&);   // pfMCMC   (optional)
//       SMCB_HD static void weight(long t, double (&x)[D], double& logw, const Params&);             // pfWeight (optional)
//   };
//
// The sequencing of a step is the reference's (sampler.h:481-550 Initialise, :701-764 IterateEss): move -> logNC increment ->
// normalise -> ESS -> resampling decision -> adaptation hook -> Resample -> DoMCMC -> renormalise -> adaptation hook ->
// history append -> T++.  The adaptation hooks are host callbacks, as in the reference; they see the population through
// smcb::Population, whose reductions (log-sum-exp, sums of user functors) run on the device.  Per step the host reads back
// one small struct (ESS, log-sum-exp, acceptance count); nothing of size N crosses PCIe unless a getter asks for it.
//
// Resampling is the engine's own (resample.cuh: all four ResampleType modes); it is applied lazily -- the next kernel that
// touches the population gathers its parents through the ancestor map -- unless an MCMC step or a history append needs the
// resampled population at once.  HistoryType RAM / AL keep every generation on the device (history.h:63-120), with
// IterateBack (sampler.h:684-699), the history getters (:153-177), ancestral lines (:445-476) and the path-sampling
// integrals (:602-669) on top.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "resample.cuh"

#ifndef SMCB_HD
#define SMCB_HD __host__ __device__
#endif

namespace smcb {

// ---- counter-based random stream handed to the user functors --------------------------------------------------------------
// One stream per (seed, kind of call, time step, particle); successive draws advance a slot counter, so a functor may draw as
// many variates as it likes and the values depend on nothing but (seed, stream, step, particle id, position in the sequence).
// Host and device produce the same uniforms bit for bit (Philox4x32-10 is integer arithmetic); normals differ by the last
// ulps of log/sqrt/sincos between libm and the device versions.
struct Rng {
    uint64_t seed, id;
    uint32_t stream, step, slot;
    uint64_t lane_b;
    double spare_z;
    bool have_b, have_z;
    SMCB_HD Rng(uint64_t seed_, uint32_t stream_, uint32_t step_, uint64_t id_)
        : seed(seed_), id(id_), stream(stream_), step(step_), slot(0u), lane_b(0ull), spare_z(0.0), have_b(false), have_z(false) {}
    SMCB_HD uint64_t bits64() {
        if (have_b) { have_b = false; return lane_b; }
        const Draw128 d = philox_draw(seed, stream, step, id, slot++);
        lane_b = d.b; have_b = true;
        return d.a;
    }
    SMCB_HD double uniform() { return u64_to_unit(bits64()); }          // [0, 1), 53 bits
    SMCB_HD double uniform_open() { return u64_to_open01(bits64()); }   // (0, 1), safe for log
    SMCB_HD double normal() {                                           // Box-Muller, two normals per Philox call
        if (have_z) { have_z = false; return spare_z; }
        have_b = false;
        const Draw128 d = philox_draw(seed, stream, step, id, slot++);
        const double u1 = u64_to_open01(d.a), u2 = u64_to_unit(d.b);
        double r, s, c;
#ifdef __CUDA_ARCH__
        r = fm_sqrt(fm_neg2log(u1));
        fm_sincos2pi(u2, s, c);
#else
        r = std::sqrt(-2.0 * std::log(u1));
        s = std::sin(6.283185307179586476925286766559 * u2);
        c = std::cos(6.283185307179586476925286766559 * u2);
#endif
        spare_z = r * s; have_z = true;
        return r * c;
    }
    SMCB_HD double exponential() { return -log(uniform_open()); }
};

// ---- optional members of a model -----------------------------------------------------------------------------------------
template <class M, class = void> struct has_mcmc : std::false_type {};
template <class M>
struct has_mcmc<M, std::void_t<decltype(M::mcmc(0L, std::declval<double (&)[M::D]>(), std::declval<double&>(),
                                                 std::declval<const typename M::Params&>(), std::declval<Rng&>()))>> : std::true_type {};
template <class M, class = void> struct has_weight : std::false_type {};
template <class M>
struct has_weight<M, std::void_t<decltype(M::weight(0L, std::declval<double (&)[M::D]>(), std::declval<double&>(),
                                                     std::declval<const typename M::Params&>()))>> : std::true_type {};

struct GenCtrl {                 // what a step reports back to the host (one D2H per kernel that ends a phase)
    double lse;                  // log-sum-exp of the stored log-weights (population.h:46-50)
    double ess;                  // (sum w)^2 / sum w^2 (sampler.h:413-417)
    double mx;
    double u_res;                // resampling uniform of the step
    long long n_accepted;
    unsigned int done, ticket, ticket2, pad;
    unsigned long long floor_total;
    long long step;              // evolution time the resampling draws are keyed on
    int enable;                  // resampling passes run iff non-zero
};

template <class Model>
struct GenArgs {
    double* src[Model::D];
    double* dst[Model::D];
    double* lw;
    long long n;
    AncestorMap am;
    int gather;                  // parents are read through the ancestor map (a resampling pass is pending)
    int reset_weights;           // resampled: log-weights restart at -log N (sampler.h:867)
    double lse_prev;             // normaliser of the stored log-weights (lazy: applied when they are read)
    double log_n;
    long t;
    int reps;                    // MCMC repeats (phase 2)
    typename Model::Params prm;
    unsigned long long seed;
    GenCtrl* ctrl;
    double* partial;
};

enum { GEN_INIT = 0, GEN_MOVE = 1, GEN_POST = 2 };

// Merge of the per-thread log-sum-exp accumulators of a grid; lane 0 of the last block's first warp gets the total.
template <int G>
__device__ __forceinline__ bool gen_lse_to_grid(LseAcc<G>& acc, double* partial, unsigned int* done) {
    __shared__ double s_red[8][3 + (G > 0 ? G : 1)];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    acc.warp_reduce();
    if (lane == 0) {
        s_red[warp][0] = acc.m; s_red[warp][1] = acc.s; s_red[warp][2] = acc.q;
#pragma unroll
        for (int j = 0; j < G; ++j) s_red[warp][3 + j] = acc.g[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        LseAcc<G> b; b.clear();
        for (int w = 0; w < ((int)blockDim.x >> 5); ++w) {
            LseAcc<G> t; t.m = s_red[w][0]; t.s = s_red[w][1]; t.q = s_red[w][2];
#pragma unroll
            for (int j = 0; j < G; ++j) t.g[j] = s_red[w][3 + j];
            b.merge(t);
        }
        const int nb = gridDim.x;
        partial[blockIdx.x] = b.m; partial[nb + blockIdx.x] = b.s; partial[2 * nb + blockIdx.x] = b.q;
#pragma unroll
        for (int j = 0; j < G; ++j) partial[(3 + j) * nb + blockIdx.x] = b.g[j];
        __threadfence();
        s_last = atomicAdd(done, 1u) == (unsigned)nb - 1u;
    }
    __syncthreads();
    if (!s_last || warp != 0) return false;
    __threadfence();
    const int nb = gridDim.x;
    const volatile double* P = partial;
    LseAcc<G> b; b.clear();
    for (int k = lane; k < nb; k += 32) {
        LseAcc<G> t; t.m = P[k]; t.s = P[nb + k]; t.q = P[2 * nb + k];
#pragma unroll
        for (int j = 0; j < G; ++j) t.g[j] = P[(3 + j) * nb + k];
        b.merge(t);
    }
    b.warp_reduce();
    acc = b;
    return lane == 0;
}

// One kernel for the three per-particle phases of a step (moveset.h:133-168):
//   GEN_INIT  pfInitialise, log-weight - log N (sampler.h:495)
//   GEN_MOVE  [gather parents] + pfMove
//   GEN_POST  [gather parents] + nRepeats x pfMCMC (moveset.h:140-160), acceptance count
// each followed by the log-sum-exp / ESS reduction of the log-weights it leaves behind.
template <class Model, int PHASE>
__global__ void __launch_bounds__(256) k_gen(GenArgs<Model> a) {
    constexpr int D = Model::D;
    LseAcc<0> acc; acc.clear();
    int accepted = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        double x[D], lw;
        if (PHASE == GEN_INIT) {
            Rng rng(a.seed, STREAM_INIT, 0u, (uint64_t)i);
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = 0.0;
            lw = 0.0;
            Model::init(x, lw, a.prm, rng);
            lw -= a.log_n;
        } else {
            const long long p = a.gather ? (long long)decode_ancestor(a.am, i) : i;
#pragma unroll
            for (int d = 0; d < D; ++d) x[d] = a.src[d][p];
            lw = a.reset_weights ? -a.log_n : a.lw[i] - a.lse_prev;
            if (PHASE == GEN_MOVE) {
                Rng rng(a.seed, STREAM_MOVE, (uint32_t)a.t, (uint64_t)i);
                Model::move(a.t, x, lw, a.prm, rng);
            } else {
                if constexpr (has_mcmc<Model>::value) {
                    Rng rng(a.seed, STREAM_MCMC, (uint32_t)a.t, (uint64_t)i);
                    for (int r = 0; r < a.reps; ++r) accepted += Model::mcmc(a.t, x, lw, a.prm, rng) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) a.dst[d][i] = x[d];
        a.lw[i] = lw;
        acc.add(lw, nullptr);
    }
    if (PHASE == GEN_POST) {
        accepted = __reduce_add_sync(SMCB_FULL, accepted);
        if ((threadIdx.x & 31) == 0 && accepted) atomicAdd((unsigned long long*)&a.ctrl->n_accepted, (unsigned long long)accepted);
    }
    if (gen_lse_to_grid<0>(acc, a.partial, &a.ctrl->done)) {
        a.ctrl->lse = acc.m + log(acc.s);
        a.ctrl->ess = (acc.s * acc.s) / acc.q;
        a.ctrl->mx = acc.m;
        a.ctrl->done = 0u;
    }
}

// ---- population-level reductions with user functors (what Integrate / the adaptation hooks are built from) -----------------
template <int D>
struct PopView {                 // the current population as device arrays; log-weights normalised by `lse` on the fly
    const double* x[D];
    const double* lw;
    double lse;
    long long n;
    AncestorMap am;
    int gather;                  // a resampling pass is pending: slot i holds particle decode_ancestor(i) with weight 1/N
    double log_n;
};
template <int D>
__device__ __forceinline__ void pop_load(const PopView<D>& v, long long i, double (&x)[D], double& lw) {
    const long long p = v.gather ? (long long)decode_ancestor(v.am, i) : i;
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = v.x[d][p];
    lw = v.gather ? -v.log_n : v.lw[i] - v.lse;
}

// sum_i f(x_i, logw_i)[k], k < NV  (logw normalised).  F: struct with  __device__ void operator()(const double (&x)[D], double logw, double (&out)[NV]) const
template <int D, int NV, class F>
__global__ void __launch_bounds__(256) k_pop_sum(PopView<D> v, F f, double* partial, unsigned int* done, double* out) {
    double s[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += (long long)gridDim.x * blockDim.x) {
        double x[D], lw, o[NV];
        pop_load(v, i, x, lw);
        f(x, lw, o);
#pragma unroll
        for (int k = 0; k < NV; ++k) s[k] += o[k];
    }
    __shared__ double s_red[8][NV];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s_red[warp][k] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nb = gridDim.x;
#pragma unroll
        for (int k = 0; k < NV; ++k) { double t = 0.0; for (int w = 0; w < ((int)blockDim.x >> 5); ++w) t += s_red[w][k]; partial[k * nb + blockIdx.x] = t; }
        __threadfence();
        s_last = atomicAdd(done, 1u) == (unsigned)nb - 1u;
    }
    __syncthreads();
    if (!s_last || warp != 0) return;
    __threadfence();
    const int nb = gridDim.x;
    const volatile double* P = partial;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int j = lane; j < nb; j += 32) t += P[k * nb + j];
        t = warp_sum(t);
        if (lane == 0) out[k] = t;
    }
    if (lane == 0) *done = 0u;
}

// log sum_i exp(f(x_i, logw_i)).  F: struct with  __device__ double operator()(const double (&x)[D], double logw) const
template <int D, class F>
__global__ void __launch_bounds__(256) k_pop_lse(PopView<D> v, F f, double* partial, unsigned int* done, double* out) {
    LseAcc<0> acc; acc.clear();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += (long long)gridDim.x * blockDim.x) {
        double x[D], lw;
        pop_load(v, i, x, lw);
        acc.add(f(x, lw), nullptr);
    }
    if (gen_lse_to_grid<0>(acc, partial, done)) { out[0] = acc.m + log(acc.s); out[1] = (acc.s * acc.s) / acc.q; *done = 0u; }
}

// Materialise a generation for the history: values (through the pending ancestor map, if any), normalised log-weights
template <int D>
__global__ void __launch_bounds__(256) k_pop_snapshot(PopView<D> v, double* xs, double* lws) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += (long long)gridDim.x * blockDim.x) {
        double x[D], lw;
        pop_load(v, i, x, lw);
#pragma unroll
        for (int d = 0; d < D; ++d) xs[(size_t)d * v.n + i] = x[d];
        lws[i] = lw;
    }
}
static __global__ void k_iota_u32(uint32_t* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

template <class Model> class GenericSampler;

// What an adaptation hook sees (adaptMethods.h:45-51 hands it `const population<Space>&`): the population stays on the device;
// the hook asks for reductions of its own functors or for host copies.
template <class Model>
class Population {
public:
    static constexpr int D = Model::D;
    long GetNumber() const { return (long)s_->n_; }
    // log sum_i exp(f(x_i, logw_i)) with logw the normalised log-weight; optionally the ESS of those terms
    template <class F> double LogSumExp(F f, double* ess_out = nullptr) const { return s_->pop_lse(f, ess_out); }
    // sum_i f(x_i, logw_i)[0..NV)
    template <int NV, class F> void Sum(F f, double (&out)[NV]) const { s_->template pop_sum<NV>(f, out); }
    // host copies (population.h:84-112): component d of every particle / the normalised log-weights
    std::vector<double> GetValueColumn(int d) const { return s_->host_column(d); }
    std::vector<double> GetLogWeight() const { return s_->host_logweights(); }
private:
    friend class GenericSampler<Model>;
    explicit Population(GenericSampler<Model>* s) : s_(s) {}
    GenericSampler<Model>* s_;
};

// Host callbacks of the adaptation object (adaptMethods.h:45-51), same three hooks, same call sites.
template <class Model>
struct AdaptMethods {
    using Params = typename Model::Params;
    virtual ~AdaptMethods() {}
    virtual void updateForMove(Params&, const Population<Model>&) {}
    virtual void updateForMCMC(Params&, const Population<Model>&, double /*acceptProb*/, int /*nResampled*/, int& /*nRepeats*/) {}
    virtual void updateEnd(Params&, const Population<Model>&) {}
};

struct HistoryFlags { int resampled; };   // history.h:37-48

template <class Model>
struct HistoryElement {          // history.h:63-120: one generation, kept on the device
    long number = 0;
    double* x = nullptr;         // [D][N]
    double* lw = nullptr;        // [N] normalised log-weights
    uint32_t* anc = nullptr;     // [N] uRSIndices of the generation (AL only)
    int n_accepted = 0, n_repeats = 0;
    HistoryFlags flags{0};
    double ess = 0.0;
};

template <class Model>
class GenericSampler {
public:
    static constexpr int D = Model::D;
    using Params = typename Model::Params;

    GenericSampler(long n, int history_mode, unsigned long long seed = 1, cudaStream_t stream = 0)
        : n_(n), hist_mode_(history_mode), seed_(seed), stream_(stream) {
        if (n < 1 || n > 2147483647L) throw std::invalid_argument("particle count must be in [1, 2^31)");
        if (history_mode < HISTORY_NONE || history_mode > HISTORY_AL) throw std::invalid_argument("unknown history mode");
        int dev = 0, sms = 0;
        SMCB_CUDA(cudaGetDevice(&dev));
        SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        long long need = (n + 255) / 256;
        grid_ = (int)std::min<long long>(need, (long long)sms * 4);
        ntiles_ = scan_tiles(n);
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
        size_t o_xa = take(sizeof(double) * n * D), o_xb = take(sizeof(double) * n * D), o_lw = take(sizeof(double) * n);
        size_t o_zm = take(sizeof(uint32_t) * ((n + 31) / 32 + 64)), o_zb = take(sizeof(uint32_t) * ((n + 31) / 32 + 64));
        size_t o_su = take(sizeof(uint32_t) * n), o_cs = take(sizeof(uint32_t) * resample_cslot_entries(n));
        size_t o_st = take(sizeof(unsigned long long) * 3 * ntiles_);
        size_t o_ct = take(sizeof(GenCtrl)), o_pa = take(sizeof(double) * 16 * grid_), o_out = take(sizeof(double) * 64);
        SMCB_CUDA(cudaMalloc((void**)&slab_, off));
        SMCB_CUDA(cudaMemsetAsync(slab_, 0, off, stream_));
        for (int d = 0; d < D; ++d) { xa_[d] = (double*)(slab_ + o_xa) + (size_t)d * n; xb_[d] = (double*)(slab_ + o_xb) + (size_t)d * n; }
        lw_ = (double*)(slab_ + o_lw);
        am_.zmask = (uint32_t*)(slab_ + o_zm); am_.zbase = (uint32_t*)(slab_ + o_zb); am_.surplus = (uint32_t*)(slab_ + o_su);
        cslot_ = (uint32_t*)(slab_ + o_cs);
        status_ = (unsigned long long*)(slab_ + o_st);
        ctrl_ = (GenCtrl*)(slab_ + o_ct);
        partial_ = (double*)(slab_ + o_pa);
        out_ = (double*)(slab_ + o_out);
        SMCB_CUDA(cudaMallocHost((void**)&hctrl_, sizeof(GenCtrl)));
        SMCB_CUDA(cudaMallocHost((void**)&hout_, sizeof(double) * 64));
        thr_ = 0.5 * (double)n;   // sampler.h:259-267
    }
    ~GenericSampler() {
        clear_history();
        cudaFree(slab_);
        if (cum_) { cudaFree(cum_); cudaFree(count_); cudaFree(bucket_); }
        cudaFreeHost(hctrl_); cudaFreeHost(hout_);
    }
    GenericSampler(const GenericSampler&) = delete;
    GenericSampler& operator=(const GenericSampler&) = delete;

    // ---- configuration (sampler.h:214-232)
    void SetResampleParams(int mode, double threshold) {   // sampler.h:885-893
        if (mode < MULTINOMIAL || mode > SYSTEMATIC) throw std::invalid_argument("unknown resampling mode");
        mode_ = mode;
        thr_ = threshold < 1 ? threshold * (double)n_ : threshold;
    }
    void SetAlgParam(const Params& p) { prm_ = p; }
    const Params& GetAlgParams() const { return prm_; }
    void SetAdaptMethods(AdaptMethods<Model>* a) { adapt_ = a; }     // not owned (sampler.h:270-271)
    void SetMcmcRepeats(int reps) { n_repeats_ = reps; }
    void SetSeed(unsigned long long seed) { seed_ = seed; }

    // ---- sampler.h:481-550
    void Initialise() {
        T_ = 0; lognc_path_ = 0.0; accept_prob_ = -1.0; pending_ = false; parity_ = 0; lse_ = 0.0;
        clear_history();
        launch<GEN_INIT>(/*gather*/ false, /*reset*/ false, 0L, 0);
        finish_step(0L);
    }
    // sampler.h:701-764; returns the ESS seen by the resampling decision
    double IterateEss() {
        Population<Model> pop(this);
        if (adapt_) adapt_->updateForMove(prm_, pop);
        launch<GEN_MOVE>(pending_, pending_, T_ + 1, 0);
        const double ess = finish_step(T_ + 1);
        ++T_;
        return ess;
    }
    void Iterate() { (void)IterateEss(); }
    void IterateUntil(long t_end) { while (T_ < t_end) Iterate(); }
    // sampler.h:684-699
    void IterateBack() {
        if (hist_mode_ == HISTORY_NONE) throw std::runtime_error("SMCX_MISSING_HISTORY: an attempt to undo an iteration was made, but the system history has not been stored");
        if (history_.size() < 2) throw std::runtime_error("IterateBack: no earlier generation in the history");
        free_element(history_.back());
        history_.pop_back();
        const HistoryElement<Model>& h = history_.back();
        double* const* cur = parity_ ? xb_ : xa_;
        for (int d = 0; d < D; ++d) SMCB_CUDA(cudaMemcpyAsync(cur[d], h.x + (size_t)d * n_, sizeof(double) * n_, cudaMemcpyDeviceToDevice, stream_));
        SMCB_CUDA(cudaMemcpyAsync(lw_, h.lw, sizeof(double) * n_, cudaMemcpyDeviceToDevice, stream_));
        pending_ = false; lse_ = 0.0;
        n_accepted_ = h.n_accepted; n_resampled_ = h.flags.resampled; n_repeats_ = h.n_repeats;
        --T_;
    }

    // ---- state getters (sampler.h:145-197)
    long GetNumber() const { return n_; }
    long GetTime() const { return T_; }
    double GetESS() { double e; (void)pop_lse(IdentityLogW{}, &e); return e; }
    double GetESS(long gen) const { return history_.at(gen).ess; }
    int GetAccepted() const { return n_accepted_; }
    int GetResampled() const { return n_resampled_; }
    int GetMcmcRepeats() const { return n_repeats_; }
    double GetAcceptProb() const { return accept_prob_; }
    double GetLogNCPath() const { return lognc_path_; }
    double GetLogNCStep() const { return lognc_step_; }
    double GetNCPath() const { return std::exp(lognc_path_); }
    double GetNCStep() const { return std::exp(lognc_step_); }
    long GetHistoryLength() const { return (long)history_.size(); }
    HistoryFlags GetHistoryFlags(long gen) const { return history_.at(gen).flags; }
    int GetHistorymcmcRepeats(long gen) const { return history_.at(gen).n_repeats; }
    int GetHistoryAccepted(long gen) const { return history_.at(gen).n_accepted; }
    // the current particle values, component d (population.h:84-95) / normalised log-weights / weights
    std::vector<double> GetParticleValues(int d) { return host_column(d); }
    std::vector<double> GetParticleLogWeight() { return host_logweights(); }
    std::vector<double> GetParticleWeight() { std::vector<double> w = host_logweights(); for (double& v : w) v = std::exp(v); return w; }
    // generation `gen` of the history: component d of every particle / normalised log-weights (history.h:96-108)
    std::vector<double> GetHistoryValues(long gen, int d) const {
        const HistoryElement<Model>& h = history_.at(gen);
        std::vector<double> out(n_);
        SMCB_CUDA(cudaMemcpy(out.data(), h.x + (size_t)d * n_, sizeof(double) * n_, cudaMemcpyDeviceToHost));
        return out;
    }
    std::vector<double> GetHistoryLogWeight(long gen) const {
        const HistoryElement<Model>& h = history_.at(gen);
        std::vector<double> out(n_);
        SMCB_CUDA(cudaMemcpy(out.data(), h.lw, sizeof(double) * n_, cudaMemcpyDeviceToHost));
        return out;
    }
    // uRSIndices of the most recent iteration (sampler.h:190-192): identity when it did not resample
    std::vector<unsigned int> GetuRSIndices() {
        std::vector<unsigned int> out(n_);
        if (!last_anc_valid_) { for (long i = 0; i < n_; ++i) out[i] = (unsigned)i; return out; }
        SMCB_CUDA(cudaMemcpy(out.data(), last_anc(), sizeof(uint32_t) * n_, cudaMemcpyDeviceToHost));
        return out;
    }
    std::vector<unsigned int> GetHistoryAncestors(long gen) const {
        if (hist_mode_ != HISTORY_AL) throw std::runtime_error("ancestor indices are stored in HistoryType::AL only");
        std::vector<unsigned int> out(n_);
        SMCB_CUDA(cudaMemcpy(out.data(), history_.at(gen).anc, sizeof(uint32_t) * n_, cudaMemcpyDeviceToHost));
        return out;
    }
    // sampler.h:445-459: indices of the ancestral line of particle n of the last generation, one per generation
    std::vector<unsigned int> GetALineInd(long n) const {
        if (hist_mode_ != HISTORY_AL) throw std::runtime_error("ancestral lines need HistoryType::AL");
        const long L = (long)history_.size();
        std::vector<unsigned int> line(L);
        unsigned int cur = (unsigned int)n;
        line[L - 1] = cur;
        for (long g = L - 1; g > 0; --g) {   // the ancestor row of generation g maps its slots to slots of generation g-1
            unsigned int a;
            SMCB_CUDA(cudaMemcpy(&a, history_[g].anc + cur, sizeof a, cudaMemcpyDeviceToHost));
            cur = a;
            line[g - 1] = cur;
        }
        return line;
    }
    // sampler.h:461-476: the values along that line, [generation][component]
    std::vector<std::vector<double>> GetALineSpace(long n) const {
        const std::vector<unsigned int> line = GetALineInd(n);
        std::vector<std::vector<double>> out(line.size(), std::vector<double>(D));
        for (size_t g = 0; g < line.size(); ++g)
            for (int d = 0; d < D; ++d)
                SMCB_CUDA(cudaMemcpy(&out[g][d], history_[g].x + (size_t)d * n_ + line[g], sizeof(double), cudaMemcpyDeviceToHost));
        return out;
    }

    // ---- sampler.h:559-571: sum_i w_i f(x_i) under the current normalised weights.  F: __device__ double operator()(const double (&x)[D]) const
    template <class F> double Integrate(F f) {
        double out[2];
        pop_sum<2>(IntegrateOp<F>{f}, out);
        return out[0] / out[1];
    }
    // ---- sampler.h:602-669.  F: __device__ double operator()(long t, const double (&x)[D]) const;  W: host callable double(long t)
    template <class F, class W> double IntegratePathSampling(int ps_type, F f, W width) {
        if (hist_mode_ == HISTORY_NONE) throw std::runtime_error("SMCX_MISSING_HISTORY: the path sampling integral cannot be computed as the history of the system was not stored");
        const long L = (long)history_.size();
        std::vector<double> E(L), V(L);
        for (long g = 0; g < L; ++g) history_moments(g, f, E[g], V[g]);
        long double r = 0.0L;
        for (long t = 1; t < L; ++t) {
            const long double w = (long double)width(t);
            if (ps_type == 0) r += (long double)E[t] * w;                                            // RECTANGLE
            else if (ps_type == 1) r += w / 2.0L * ((long double)E[t - 1] + (long double)E[t]);       // TRAPEZOID1
            else r += w / 2.0L * ((long double)E[t - 1] + (long double)E[t]) - w * w / 12.0L * ((long double)V[t] - (long double)V[t - 1]);   // TRAPEZOID2
        }
        return (double)r;
    }

    void Sync() { SMCB_CUDA(cudaStreamSynchronize(stream_)); }

    // (functor types that reach a kernel template must be public: nvcc restriction)
    struct IdentityLogW { __device__ double operator()(const double (&)[D], double lw) const { return lw; } };
    template <class F> struct IntegrateOp {
        F f;
        __device__ void operator()(const double (&x)[D], double lw, double (&o)[2]) const { const double w = exp(lw); o[0] = w * f(x); o[1] = w; }
    };
    template <class F> struct PathOp {
        F f; long t; double shift;
        __device__ void operator()(const double (&x)[D], double lw, double (&o)[3]) const {
            const double w = exp(lw), v = f(t, x) - shift;
            o[0] = w; o[1] = w * v; o[2] = w * v * v;
        }
    };

private:
    friend class Population<Model>;
    PopView<D> view() const {
        PopView<D> v;
        double* const* cur = parity_ ? xb_ : xa_;
        for (int d = 0; d < D; ++d) v.x[d] = cur[d];
        v.lw = lw_; v.lse = lse_; v.n = n_; v.am = am_; v.gather = pending_ ? 1 : 0; v.log_n = std::log((double)n_);
        return v;
    }
    template <class F> double pop_lse(F f, double* ess_out) {
        k_pop_lse<D, F><<<grid_, 256, 0, stream_>>>(view(), f, partial_, &ctrl_->done, out_);
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpyAsync(hout_, out_, sizeof(double) * 2, cudaMemcpyDeviceToHost, stream_));
        Sync();
        if (ess_out) *ess_out = hout_[1];
        return hout_[0];
    }
    template <int NV, class F> void pop_sum(F f, double (&out)[NV]) {
        static_assert(NV <= 16, "at most 16 sums per pass");
        k_pop_sum<D, NV, F><<<grid_, 256, 0, stream_>>>(view(), f, partial_, &ctrl_->done, out_);
        SMCB_CUDA(cudaGetLastError());
        SMCB_CUDA(cudaMemcpyAsync(hout_, out_, sizeof(double) * NV, cudaMemcpyDeviceToHost, stream_));
        Sync();
        for (int k = 0; k < NV; ++k) out[k] = hout_[k];
    }
    template <class F> void history_moments(long g, F f, double& mean, double& var) {   // history.h:158-188
        const HistoryElement<Model>& h = history_[g];
        PopView<D> v;
        for (int d = 0; d < D; ++d) v.x[d] = h.x + (size_t)d * n_;
        v.lw = h.lw; v.lse = 0.0; v.n = n_; v.am = am_; v.gather = 0; v.log_n = 0.0;
        double o[3];
        for (int pass = 0; pass < 2; ++pass) {   // second pass centred on the mean: no cancellation in the variance
            const double shift = pass ? mean : 0.0;
            k_pop_sum<D, 3, PathOp<F>><<<grid_, 256, 0, stream_>>>(v, PathOp<F>{f, g, shift}, partial_, &ctrl_->done, out_);
            SMCB_CUDA(cudaGetLastError());
            SMCB_CUDA(cudaMemcpyAsync(hout_, out_, sizeof(double) * 3, cudaMemcpyDeviceToHost, stream_));
            Sync();
            o[0] = hout_[0]; o[1] = hout_[1]; o[2] = hout_[2];
            if (!pass) mean = o[1] / o[0];
        }
        var = o[2] / o[0] - (o[1] / o[0]) * (o[1] / o[0]);
    }
    std::vector<double> host_column(int d) {
        if (d < 0 || d >= D) throw std::invalid_argument("state component out of range");
        materialise();
        std::vector<double> out(n_);
        SMCB_CUDA(cudaMemcpyAsync(out.data(), (parity_ ? xb_ : xa_)[d], sizeof(double) * n_, cudaMemcpyDeviceToHost, stream_));
        Sync();
        return out;
    }
    std::vector<double> host_logweights() {
        materialise();
        std::vector<double> out(n_);
        SMCB_CUDA(cudaMemcpyAsync(out.data(), lw_, sizeof(double) * n_, cudaMemcpyDeviceToHost, stream_));
        Sync();
        for (double& v : out) v -= lse_;
        return out;
    }

    GenArgs<Model> args(bool gather, bool reset, long t, int reps) {
        GenArgs<Model> a;
        double* const* src = parity_ ? xb_ : xa_;
        double* const* dst = parity_ ? xa_ : xb_;
        for (int d = 0; d < D; ++d) { a.src[d] = src[d]; a.dst[d] = dst[d]; }
        a.lw = lw_; a.n = n_; a.am = am_; a.gather = gather ? 1 : 0; a.reset_weights = reset ? 1 : 0; a.lse_prev = lse_;
        a.log_n = std::log((double)n_); a.t = t; a.reps = reps; a.prm = prm_; a.seed = seed_; a.ctrl = ctrl_; a.partial = partial_;
        return a;
    }
    // launches phase PHASE out of place (src -> dst), flips the buffers and fetches the step scalars
    template <int PHASE> void launch(bool gather, bool reset, long t, int reps) {
        GenArgs<Model> a = args(gather, reset, t, reps);
        if (PHASE == GEN_INIT) { for (int d = 0; d < D; ++d) a.dst[d] = xa_[d]; }
        k_gen<Model, PHASE><<<grid_, 256, 0, stream_>>>(a);
        SMCB_CUDA(cudaGetLastError());
        if (PHASE == GEN_INIT) parity_ = 0; else parity_ ^= 1;
        pending_ = false;
        SMCB_CUDA(cudaMemcpyAsync(hctrl_, ctrl_, sizeof(GenCtrl), cudaMemcpyDeviceToHost, stream_));
        Sync();
        lse_ = hctrl_->lse;
    }
    // a pending resampling pass is applied now (gather only), e.g. before a host copy of the population
    void materialise() {
        if (!pending_) return;
        const int acc = n_accepted_;
        launch<GEN_POST>(true, true, T_, 0);
        n_accepted_ = acc;
    }
    // everything of a step after the move: sampler.h:496-549 / :709-761
    double finish_step(long t) {
        lognc_step_ = hctrl_->lse;            // dlogNCIt = CalcLogNC(), log-weights normalised lazily through lse_
        lognc_path_ += lognc_step_;
        const double ess = hctrl_->ess;
        Population<Model> pop(this);
        n_resampled_ = ess < thr_ ? 1 : 0;
        if (adapt_) adapt_->updateForMCMC(prm_, pop, accept_prob_, n_resampled_, n_repeats_);
        last_anc_valid_ = false;
        if (n_resampled_) resample(t);
        // DoMCMC (moveset.h:140-160): performed iff nRepeats > 0; the resampled parents are gathered by the same kernel
        n_accepted_ = 0;
        if (n_repeats_ > 0) {
            SMCB_CUDA(cudaMemsetAsync(&ctrl_->n_accepted, 0, sizeof(long long), stream_));
            const bool was_pending = pending_;
            launch<GEN_POST>(was_pending, was_pending, t, has_mcmc<Model>::value ? n_repeats_ : 0);
            n_accepted_ = (int)hctrl_->n_accepted;
            accept_prob_ = (double)n_accepted_ / ((double)n_ * (double)n_repeats_);   // sampler.h:733-736
        }
        if (adapt_) adapt_->updateEnd(prm_, pop);
        if (hist_mode_ != HISTORY_NONE) append_history(ess);
        return ess;
    }
    void resample(long t) {
        GenCtrl h = *hctrl_;
        h.done = 0u; h.ticket = 0u; h.ticket2 = 0u; h.floor_total = 0ull; h.step = t; h.enable = 1; h.n_accepted = 0;
        h.u_res = u64_to_unit(philox_draw(seed_, STREAM_RESAMPLE, (uint32_t)t, ~0ull, 1).a);
        SMCB_CUDA(cudaMemcpyAsync(ctrl_, &h, sizeof h, cudaMemcpyHostToDevice, stream_));
        SMCB_CUDA(cudaMemsetAsync(status_, 0, sizeof(unsigned long long) * 3 * ntiles_, stream_));
        ScanArgs s{};
        s.in = lw_; s.n = n_; s.n_glob = n_; s.lse_ptr = &ctrl_->lse; s.enable_ptr = &ctrl_->enable; s.u_ptr = &ctrl_->u_res;
        s.ustrat = nullptr; s.step_ptr = &ctrl_->step; s.seed = seed_; s.ticket = &ctrl_->ticket;
        s.ws = ScanWs{status_, status_ + ntiles_, status_ + 2 * ntiles_};
        s.am = am_; s.cslot = cslot_;
        if (mode_ == SYSTEMATIC) launch_resample_passes<SRC_LOGW, SYSTEMATIC>(s, stream_);
        else if (mode_ == STRATIFIED) launch_resample_passes<SRC_LOGW, STRATIFIED>(s, stream_);
        else {
            if (!cum_) {
                SMCB_CUDA(cudaMalloc((void**)&cum_, sizeof(unsigned long long) * n_));
                SMCB_CUDA(cudaMalloc((void**)&count_, sizeof(int) * n_));
                SMCB_CUDA(cudaMalloc((void**)&bucket_, sizeof(uint32_t) * draw_bucket_entries(n_)));
            }
            DrawWs dw{cum_, count_, &ctrl_->floor_total, &ctrl_->ticket2, bucket_};
            launch_resample_draws<SRC_LOGW>(mode_, lw_, n_, &ctrl_->lse, &ctrl_->enable, nullptr, &ctrl_->step, seed_, &ctrl_->ticket, s.ws, dw, am_, nullptr, stream_);
        }
        pending_ = true;
        last_anc_valid_ = true;
        last_anc_dirty_ = true;
    }
    uint32_t* last_anc() {   // uRSIndices of the pending / most recent resampling, decoded on demand
        if (!anc_buf_) SMCB_CUDA(cudaMalloc((void**)&anc_buf_, sizeof(uint32_t) * n_));
        if (last_anc_dirty_) {
            k_decode_indices<<<(unsigned)((n_ + 255) / 256), 256, 0, stream_>>>(am_, n_, anc_buf_);
            SMCB_CUDA(cudaGetLastError());
            Sync();
            last_anc_dirty_ = false;
        }
        return anc_buf_;
    }
    void append_history(double ess) {
        HistoryElement<Model> h;
        h.number = n_;
        SMCB_CUDA(cudaMalloc((void**)&h.x, sizeof(double) * n_ * D));
        SMCB_CUDA(cudaMalloc((void**)&h.lw, sizeof(double) * n_));
        if (hist_mode_ == HISTORY_AL) {
            SMCB_CUDA(cudaMalloc((void**)&h.anc, sizeof(uint32_t) * n_));
            if (n_resampled_) {
                uint32_t* src = last_anc();   // decoded BEFORE the snapshot below can disturb anything: the map is read-only here
                SMCB_CUDA(cudaMemcpyAsync(h.anc, src, sizeof(uint32_t) * n_, cudaMemcpyDeviceToDevice, stream_));
            } else {
                k_iota_u32<<<(unsigned)((n_ + 255) / 256), 256, 0, stream_>>>(h.anc, n_);   // sampler.h:724-726
                SMCB_CUDA(cudaGetLastError());
            }
        }
        k_pop_snapshot<D><<<grid_, 256, 0, stream_>>>(view(), h.x, h.lw);
        SMCB_CUDA(cudaGetLastError());
        h.n_accepted = n_accepted_; h.n_repeats = n_repeats_; h.flags.resampled = n_resampled_;
        h.ess = n_resampled_ ? (double)n_ : ess;   // historyelement::GetESS of the stored (post-resampling) weights
        history_.push_back(h);
    }
    void free_element(HistoryElement<Model>& h) { cudaFree(h.x); cudaFree(h.lw); if (h.anc) cudaFree(h.anc); h.x = h.lw = nullptr; h.anc = nullptr; }
    void clear_history() { for (auto& h : history_) free_element(h); history_.clear(); }

    long n_, T_ = 0;
    int hist_mode_, mode_ = STRATIFIED, grid_ = 1;
    long long ntiles_ = 0;
    double thr_ = 0.0, lse_ = 0.0, lognc_path_ = 0.0, lognc_step_ = 0.0, accept_prob_ = -1.0;
    int n_accepted_ = 0, n_resampled_ = 0, n_repeats_ = 1;   // default of one MCMC repeat: sampler.h:267
    bool pending_ = false, last_anc_valid_ = false, last_anc_dirty_ = false;
    int parity_ = 0;
    unsigned long long seed_;
    cudaStream_t stream_;
    Params prm_{};
    AdaptMethods<Model>* adapt_ = nullptr;
    unsigned char* slab_ = nullptr;
    double* xa_[D];
    double* xb_[D];
    double* lw_ = nullptr;
    AncestorMap am_{};
    uint32_t* cslot_ = nullptr;
    unsigned long long* status_ = nullptr;
    GenCtrl* ctrl_ = nullptr;
    GenCtrl* hctrl_ = nullptr;
    double* partial_ = nullptr;
    double* out_ = nullptr;
    double* hout_ = nullptr;
    unsigned long long* cum_ = nullptr;
    int* count_ = nullptr;
    uint32_t* bucket_ = nullptr;
    uint32_t* anc_buf_ = nullptr;
    std::vector<HistoryElement<Model>> history_;
};

}  // namespace smcb
