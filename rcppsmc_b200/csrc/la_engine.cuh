// Likelihood-annealing SMC sampler for static Bayesian models on the device: the reference's
// smc::sampler<rad_state, smc::staticModelAdapt> loop with adaptation hooks and MCMC rejuvenation
// (src/LinReg_LA_adapt.cpp:36-184, inst/include/LinReg_LA_adapt.h:55-99, inst/include/staticModelAdapt.h:81-171,
// moveset.h:140-160 DoMCMC, sampler.h:745-759 RAM history, sampler.h:602-669 path sampling) and the fixed-schedule variant
// (src/LinReg_LA.cpp:45-187).
//
// One temperature step = a short chain of kernels that never returns to the host except for ONE scalar read
// (current temperature) that drives the `while (temp != 1)` loop:
//   k_la_cess  x K   CESS bisection as a device state machine: each launch evaluates CESSdiff at the pending temperature
//                    (two log-sum-exp reductions, last-block merge) and advances the reference's bisection by one step
//   k_la_reweight    lw += (temp_t - temp_{t-1}) loglike; online (m, s, q, s*loglike) -> logNC increment, ESS, decision
//   k_la_cov         unweighted mean + weighted scatter of theta -> 3x3 Cholesky factor, MCMC repeat count
//   k_resample_scan  (resample.cuh) systematic resampling, skipped on device when ESS >= threshold
//   k_la_mcmc        gather(ancestor) + nRepeats random-walk MH moves per particle in registers + acceptance count
//                    + RAM-history append + path-sampling moments of the finished generation
#pragma once
#include <vector>

#include "common.cuh"
#include "philox.cuh"
#include "resample.cuh"

namespace smcb {

constexpr int LA_MAX_OBS = 256;   // observations held in constant-like kernel parameters' device table

struct LaCtrl {
    long long T;                 // generation index (sampler::T)
    double temp_prev, temp_cur;
    double lse;                  // normaliser of the stored log-weights (0 once a generation is finished)
    double path;                 // dlogNCPath
    double ess;                  // ESS seen by the resampling decision
    double thr, log_n;
    double u_res;
    double chol[9];              // upper factor U (column-major), proposal = theta + U z (LinReg_LA_adapt.cpp:168)
    double ps_shift;             // centring constant for the path-sampling variance
    double accept_prob;          // sampler::acceptProb (-1 before the first MCMC step)
    int resampled, n_repeats, n_accepted, failed;
    int parity;
    unsigned int done, ticket;
    // CESS bisection state machine (staticModelAdapt.h:89-130)
    int bis_phase, bis_done;
    double bis_a, bis_b, bis_fa, bis_fb, bis_m, bis_delta, bis_curr, bis_desired;
    long long zpos, upos;        // consumption offsets into the injected draw streams
    // per-generation records (device arrays are indexed by T)
};

struct LaData {                  // regression data: y_k, x_k - mean(x)
    const double* y;
    const double* xc;
    int n_obs;
};

struct LaBuffers {
    double* a[5];                // theta0, theta1, theta2, loglike, logprior  (buffer A)
    double* b[5];                // buffer B
    double* lw;
    long long n;
};

struct LaHistory {               // RAM history: [Lcap][N] per component, plus per-generation scalars
    double* comp[6];             // theta0..2, loglike, logprior, weight
    double* ess; double* temps; double* repeats; double* ps_mean; double* ps_var; double* accept;
    long long lcap;
};

// log-likelihood / log-prior of the radiata regression (LinReg_LA_adapt.cpp:118-129).  The 42-term sum keeps
// Armadillo's two interleaved accumulators so that values track the CPU reference to the last few ulp.
__device__ __forceinline__ double la_loglik(const double th[3], const LaData& d) {
    const double sigma = sqrt(exp(th[2]));
    const double ls = log(sigma), inv = 2.0 * sigma * sigma, c = 0.5 * log(2.0 * 3.14159265358979323846);
    double acc0 = 0.0, acc1 = 0.0;
    int k = 0;
    for (; k + 1 < d.n_obs; k += 2) {
        double r0 = d.y[k] - (th[0] + th[1] * d.xc[k]);
        double r1 = d.y[k + 1] - (th[0] + th[1] * d.xc[k + 1]);
        acc0 += -ls - (r0 * r0) / inv - c;
        acc1 += -ls - (r1 * r1) / inv - c;
    }
    if (k < d.n_obs) { double r0 = d.y[k] - (th[0] + th[1] * d.xc[k]); acc0 += -ls - (r0 * r0) / inv - c; }
    return acc0 + acc1;
}
__device__ __forceinline__ double la_logprior(const double th[3]) {
    const double a_prior = 3.0, inv_b = 180000.0;  // 1 / b_prior, b_prior = (2 * 300^2)^-1
    double d0 = th[0] - 3000.0, d1 = th[1] - 185.0;
    return -log(1000.0) - (d0 * d0) / (2.0 * 1000.0 * 1000.0) - log(100.0) - (d1 * d1) / (2.0 * 100.0 * 100.0) + th[2]
           - inv_b / exp(th[2]) - th[2] * (a_prior + 1.0);
}

// Data annealing (src/LinReg.cpp:87-118): weight of observation t, and the log posterior given observations 0..t.
__device__ __forceinline__ double lr_logweight(int t, const double th[3], const LaData& d) {
    const double mean_reg = th[0] + th[1] * d.xc[t];
    const double sigma = sqrt(exp(th[2]));
    const double r = d.y[t] - mean_reg;
    return -log(sigma) - (r * r) / (2.0 * sigma * sigma) - 0.5 * log(2.0 * 3.14159265358979323846);
}
__device__ __forceinline__ double lr_logposterior(int t, const double th[3], const LaData& d) {
    LaData part = d;
    part.n_obs = t + 1;
    return la_loglik(th, part) + la_logprior(th);   // log_normpdf + log_prior (:117)
}

// Block partials -> last block: generic helper for the small reductions of this file.
template <int NV>
__device__ __forceinline__ bool la_block_to_grid(const double (&v)[NV], double* partial, unsigned int* done, double (&tot)[NV]) {
    __shared__ double s_red[8][NV];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double w[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) w[j] = warp_sum(v[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) s_red[warp][j] = w[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nb = gridDim.x;
        const int nw = (int)blockDim.x >> 5;
#pragma unroll
        for (int j = 0; j < NV; ++j) { double s = 0.0; for (int q = 0; q < nw; ++q) s += s_red[q][j]; partial[j * nb + blockIdx.x] = s; }
        __threadfence();
        s_last = atomicAdd(done, 1u) == (unsigned)nb - 1u;
    }
    __syncthreads();
    if (!s_last || warp != 0) return false;
    __threadfence();
    const int nb = gridDim.x;
    const volatile double* P = partial;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double s = 0.0;
        for (int k = lane; k < nb; k += 32) s += P[j * nb + k];
        tot[j] = warp_sum(s);
    }
    return lane == 0;
}

// Merge of LSE partials written as (m, s, extra...) per block; returns true in lane 0 of the last block's warp 0.
template <int G>
__device__ __forceinline__ bool la_lse_to_grid(LseAcc<G>& acc, double* partial, unsigned int* done) {
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

struct LaArgs {
    LaBuffers buf;
    LaData data;
    LaCtrl* ctrl;
    LaHistory hist;
    AncestorMap am;
    double* partial;             // scratch for block partials (>= 16 * grid doubles)
    unsigned long long seed;
    const double* zin;           // injected standard normals (reference consumption order) or null
    const double* uin;           // injected uniforms or null
    int adaptive;                // 1: LinRegLA_adapt, 0: LinRegLA (fixed schedule / proposal / repeats)
    int data_annealing;          // 1: LinReg (src/LinReg.cpp): one observation per step, MH target = prior x likelihood(0..t)
    double rho_n;                // target conditional ESS (tempTol * N)
};

// ---- pfInitialise (LinReg_LA_adapt.cpp:137-148) + first normalisation (sampler.h:489-506) ----------------------------
static __global__ void __launch_bounds__(256) k_la_init(LaArgs a) {
    LaCtrl* c = a.ctrl;
    const long long n = a.buf.n;
    const double temp0 = c->temp_cur, log_n = c->log_n;
    const double b_prior = 1.0 / 180000.0;
    LseAcc<1> acc; acc.clear();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double z0, z1, u[3];
        if (a.zin) {
            z0 = a.zin[2 * i]; z1 = a.zin[2 * i + 1];
            for (int k = 0; k < 3; ++k) u[k] = a.uin[3 * i + k];
        } else {
            philox_normal2(a.seed, STREAM_INIT, 0u, (uint64_t)i, 0, z0, z1);
            Draw128 d0 = philox_draw(a.seed, STREAM_INIT, 0u, (uint64_t)i, 1), d1 = philox_draw(a.seed, STREAM_INIT, 0u, (uint64_t)i, 2);
            u[0] = u64_to_open01(d0.a); u[1] = u64_to_open01(d0.b); u[2] = u64_to_open01(d1.a);
        }
        double th[3];
        th[0] = 3000.0 + 1000.0 * z0;
        th[1] = 185.0 + 100.0 * z1;
        double g = 0.0;
        for (int k = 0; k < 3; ++k) g += -log(u[k]);                 // Gamma(3, 1) as a sum of exponentials
        th[2] = log(1.0 / (b_prior * g));
        const double ll = la_loglik(th, a.data), lp = la_logprior(th);
        const double lw = (a.data_annealing ? lr_logweight(0, th, a.data) : temp0 * ll) - log_n;   // pfInitialise weight; sampler.h:495
        a.buf.a[0][i] = th[0]; a.buf.a[1][i] = th[1]; a.buf.a[2][i] = th[2]; a.buf.a[3][i] = ll; a.buf.a[4][i] = lp;
        a.buf.lw[i] = lw;
        double h[1] = {ll};
        acc.add(lw, h);
    }
    if (la_lse_to_grid<1>(acc, a.partial, &c->done)) {
        const double lse = acc.m + log(acc.s);
        c->lse = lse; c->path = lse; c->ess = (acc.s * acc.s) / acc.q;
        c->resampled = c->ess < c->thr ? 1 : 0;
        c->ps_shift = acc.g[0] / acc.s;
        c->parity = 0; c->done = 0u;
        if (a.zin) { c->zpos = 2 * n; c->upos = 3 * n; }
    }
}

// ---- one step of the CESS bisection (staticModelAdapt.h:81-130) --------------------------------------------------------
static __global__ void __launch_bounds__(256) k_la_cess(LaArgs a) {
    LaCtrl* c = a.ctrl;
    if (c->bis_done) return;
    const long long n = a.buf.n;
    const double delta = c->bis_delta, lse = c->lse;
    const double* ll = (c->parity ? a.buf.b : a.buf.a)[3];
    LseAcc<0> a1, a2; a1.clear(); a2.clear();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double lw = a.buf.lw[i] - lse, l = ll[i];
        a1.add(lw + delta * l, nullptr);
        a2.add(lw + 2 * delta * l, nullptr);
    }
    // two reductions through disjoint halves of the scratch, each with its own arrival counter (the scan ticket is free
    // outside the resampling pass).  A block bumps the second counter only after publishing its first partial, so whoever
    // closes the second reduction can re-merge the first from the scratch.
    const int nb = gridDim.x;
    (void)la_lse_to_grid<0>(a1, a.partial, &c->done);
    __syncthreads();
    if (!la_lse_to_grid<0>(a2, a.partial + 4 * nb, &c->ticket)) return;
    // the block that finished the SECOND reduction last is not necessarily the one that finished the first: re-read sum 1
    {
        const volatile double* P = a.partial;
        LseAcc<0> b; b.clear();
        for (int k = 0; k < nb; ++k) { LseAcc<0> t; t.m = P[k]; t.s = P[nb + k]; t.q = P[2 * nb + k]; b.merge(t); }
        a1 = b;
    }
    c->done = 0u; c->ticket = 0u;
    const double ls1 = a1.m + log(a1.s), ls2 = a2.m + log(a2.s);
    const double f = exp(log((double)n) + 2 * ls1 - ls2) - c->bis_desired;   // CESSdiff
    const double eps = 0.01, curr = c->bis_curr;
    switch (c->bis_phase) {
        case 0:   // ChooseTemp: can we jump straight to temperature 1?
            if (f >= -eps) { c->temp_prev = curr; c->temp_cur = 1.0; c->bis_done = 1; }
            else { c->bis_a = curr; c->bis_b = 1.0; c->bis_delta = c->bis_a - curr; c->bis_phase = 1; }
            break;
        case 1: c->bis_fa = f; c->bis_delta = c->bis_b - curr; c->bis_phase = 2; break;
        case 2:
            c->bis_fb = f;
            if (c->bis_fa * c->bis_fb > 0) { c->failed = 1; c->bis_done = 1; }   // Rcpp::stop("Bisection method ... failed")
            else { c->bis_m = (c->bis_a + c->bis_b) / 2.0; c->bis_delta = c->bis_m - curr; c->bis_phase = 3; }
            break;
        default:  // 3: first f_m (err = 10, never accepted); 4: later ones
            if (c->bis_phase == 4 && fabs(f) <= eps) { c->temp_prev = curr; c->temp_cur = c->bis_m; c->bis_done = 1; break; }
            if (f < 0.0) { c->bis_b = c->bis_m; c->bis_fb = f; } else { c->bis_a = c->bis_m; c->bis_fa = f; }
            c->bis_m = (c->bis_a + c->bis_b) / 2.0;
            c->bis_delta = c->bis_m - curr;
            c->bis_phase = 4;
            break;
    }
}

// ---- pfMove + normalising constant + ESS + resampling decision (LinReg_LA_adapt.cpp:154-157, sampler.h:707-723) -------
static __global__ void __launch_bounds__(256) k_la_reweight(LaArgs a) {
    LaCtrl* c = a.ctrl;
    const long long n = a.buf.n;
    const double dt = c->temp_cur - c->temp_prev, lse_prev = c->lse;
    double* const* cur = c->parity ? a.buf.b : a.buf.a;
    const double* ll = cur[3];
    const int tnext = (int)c->T + 1;
    LseAcc<1> acc; acc.clear();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double l = ll[i];
        double inc = dt * l;
        if (a.data_annealing) { double th[3] = {cur[0][i], cur[1][i], cur[2][i]}; inc = lr_logweight(tnext, th, a.data); }   // LinReg.cpp:139
        const double lw = (a.buf.lw[i] - lse_prev) + inc;
        a.buf.lw[i] = lw;
        double h[1] = {l};
        acc.add(lw, h);
    }
    if (la_lse_to_grid<1>(acc, a.partial, &c->done)) {
        const double lse = acc.m + log(acc.s);
        c->lse = lse; c->path += lse; c->ess = (acc.s * acc.s) / acc.q;
        c->resampled = c->ess < c->thr ? 1 : 0;
        c->ps_shift = acc.g[0] / acc.s;
        c->T += 1;
        c->done = 0u;
    }
}

// ---- rad_adapt::updateForMCMC (LinReg_LA_adapt.h:83-95): repeats + Cholesky factor of the empirical covariance ---------
// Two launches: phase 0 = unweighted column means of theta; phase 1 = weighted scatter about them (staticModelAdapt.h:137-143).
static __global__ void __launch_bounds__(256) k_la_cov(LaArgs a, int phase, double* mean3) {
    LaCtrl* c = a.ctrl;
    const long long n = a.buf.n;
    double* const* cur = c->parity ? a.buf.b : a.buf.a;
    if (phase == 0) {
        double v[3] = {0.0, 0.0, 0.0}, tot[3];
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
            v[0] += cur[0][i]; v[1] += cur[1][i]; v[2] += cur[2][i];
        }
        if (la_block_to_grid<3>(v, a.partial, &c->done, tot)) {
            for (int k = 0; k < 3; ++k) mean3[k] = tot[k] / (double)n;
            c->done = 0u;
        }
        return;
    }
    const double m0 = mean3[0], m1 = mean3[1], m2 = mean3[2], lse = c->lse;
    double v[6] = {0, 0, 0, 0, 0, 0}, tot[6];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double w = exp(a.buf.lw[i] - lse);
        const double d0 = cur[0][i] - m0, d1 = cur[1][i] - m1, d2 = cur[2][i] - m2;
        v[0] += (d0 * w) * d0; v[1] += (d0 * w) * d1; v[2] += (d0 * w) * d2;
        v[3] += (d1 * w) * d1; v[4] += (d1 * w) * d2; v[5] += (d2 * w) * d2;
    }
    if (la_block_to_grid<6>(v, a.partial, &c->done, tot)) {
        // arma::chol: upper factor U with U^T U = A
        const double A00 = tot[0], A01 = tot[1], A02 = tot[2], A11 = tot[3], A12 = tot[4], A22 = tot[5];
        const double U00 = sqrt(A00), U01 = A01 / U00, U02 = A02 / U00;
        const double U11 = sqrt(A11 - U01 * U01), U12 = (A12 - U01 * U02) / U11;
        const double U22 = sqrt(A22 - U02 * U02 - U12 * U12);
        double* U = c->chol;   // column-major
        U[0] = U00; U[1] = 0; U[2] = 0; U[3] = U01; U[4] = U11; U[5] = 0; U[6] = U02; U[7] = U12; U[8] = U22;
        // calcMcmcRepeats(acceptProb, 0.99, 10, 1000) when resampling, else 0 (LinReg_LA_adapt.h:84-88, staticModelAdapt.h:161-171)
        int reps = 0;
        if (c->resampled) {
            const double acc = c->accept_prob;
            if (acc + 1.0 <= 1e-9) reps = 10;
            else if (acc - 1.0 >= -1e-9) reps = 1;
            else if (acc <= 1e-9) reps = 1000;
            else { int r = (int)ceil(log(1.0 - 0.99) / log(1.0 - acc)); reps = r < 1000 ? r : 1000; }
        }
        c->n_repeats = reps;
        c->done = 0u;
    }
}

// resampling uniform + bookkeeping that must precede the resampling pass
static __global__ void k_la_pre_resample(LaArgs a) {
    LaCtrl* c = a.ctrl;
    if (c->resampled && !a.data_annealing) {   // LinReg resamples multinomially: its uniforms come from a separate by-step array
        if (a.uin) { c->u_res = a.uin[c->upos]; c->upos += 1; }
        else c->u_res = u64_to_unit(philox_draw(a.seed, STREAM_RESAMPLE, (uint32_t)c->T, ~0ull, 1).a);
    }
    c->n_accepted = 0;
    c->ticket = 0u;
}

// ---- replication + DoMCMC (moveset.h:140-160; pfMCMC LinReg_LA_adapt.cpp:165-179) + history append -----------------------
static __global__ void __launch_bounds__(128) k_la_mcmc(LaArgs a) {
    LaCtrl* c = a.ctrl;
    const long long n = a.buf.n;
    const int resampled = c->resampled, reps = c->n_repeats, parity = c->parity;
    const long long T = c->T;
    const double temp = c->temp_cur, lse = c->lse, log_n = c->log_n, shift = c->ps_shift;
    double* const* src = parity ? a.buf.b : a.buf.a;
    double* const* dst = parity ? a.buf.a : a.buf.b;
    double U[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) U[k] = c->chol[k];
    const long long zbase = c->zpos, ubase = c->upos;
    int accepted = 0;
    double v[3] = {0.0, 0.0, 0.0}, tot[3];   // sum w, sum w (ll - shift), sum w (ll - shift)^2
    for (long long i = (long long)blockIdx.x * 128 + threadIdx.x; i < n; i += (long long)gridDim.x * 128) {
        const long long anc = resampled ? (long long)decode_ancestor(a.am, i) : i;       // sampler.h:860-864
        double th[3] = {src[0][anc], src[1][anc], src[2][anc]};
        double ll = src[3][anc], lp = src[4][anc];
        const double lw = resampled ? -log_n : a.buf.lw[i] - lse;                          // sampler.h:867 / :714
        for (int j = 0; j < reps; ++j) {
            double z[3], u;
            if (a.zin) {
                const long long q = (long long)j * n + i;
                z[0] = a.zin[zbase + 3 * q]; z[1] = a.zin[zbase + 3 * q + 1]; z[2] = a.zin[zbase + 3 * q + 2];
                u = a.uin[ubase + q];
            } else {
                double spare;
                philox_normal2(a.seed, STREAM_MCMC, (uint32_t)T, (uint64_t)i, 4 * j, z[0], z[1]);
                philox_normal2(a.seed, STREAM_MCMC, (uint32_t)T, (uint64_t)i, 4 * j + 1, z[2], spare);
                u = u64_to_unit(philox_draw(a.seed, STREAM_MCMC, (uint32_t)T, (uint64_t)i, 4 * j + 2).a);
            }
            double pr[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {   // cholCov * z, all three terms (zeros below the diagonal) like the dense product
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += U[r + 3 * k] * z[k];
                pr[r] = th[r] + s;
            }
            if (a.data_annealing) {   // LinReg.cpp:150-166: both posteriors are recomputed from the first T+1 observations
                const double pc = lr_logposterior((int)T, th, a.data), pp = lr_logposterior((int)T, pr, a.data);
                if (exp(pp - pc) > u) { th[0] = pr[0]; th[1] = pr[1]; th[2] = pr[2]; ++accepted; }
                continue;
            }
            const double llp = la_loglik(pr, a.data), lpp = la_logprior(pr);
            const double ratio = exp(temp * (llp - ll) + lpp - lp);
            if (ratio > u) { th[0] = pr[0]; th[1] = pr[1]; th[2] = pr[2]; ll = llp; lp = lpp; ++accepted; }
        }
        dst[0][i] = th[0]; dst[1][i] = th[1]; dst[2][i] = th[2]; dst[3][i] = ll; dst[4][i] = lp;
        a.buf.lw[i] = lw;
        const double w = exp(lw);
        if (T < a.hist.lcap) {   // RAM history (sampler.h:745-759): the generation as the outputs will report it
            const size_t o = (size_t)T * (size_t)n + (size_t)i;
            a.hist.comp[0][o] = th[0]; a.hist.comp[1][o] = th[1]; a.hist.comp[2][o] = th[2];
            a.hist.comp[3][o] = ll; a.hist.comp[4][o] = lp; a.hist.comp[5][o] = w;
        }
        const double dl = ll - shift;
        v[0] += w; v[1] += w * dl; v[2] += w * dl * dl;
    }
    accepted = __reduce_add_sync(SMCB_FULL, accepted);
    if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(&c->n_accepted, accepted);
    if (la_block_to_grid<3>(v, a.partial, &c->done, tot)) {
        __threadfence();
        const int nacc = *reinterpret_cast<volatile int*>(&c->n_accepted);
        if (reps > 0) c->accept_prob = (double)nacc / ((double)n * (double)reps);   // sampler.h:733-736
        if (a.zin) { c->zpos = zbase + 3ll * reps * n; c->upos = ubase + (long long)reps * n; }
        if (T < a.hist.lcap) {
            const double e = tot[1] / tot[0];
            a.hist.ps_mean[T] = shift + e;                               // history.h:158-170
            a.hist.ps_var[T] = tot[2] / tot[0] - e * e;                  // history.h:176-188 (about the mean)
            a.hist.ess[T] = resampled ? (double)n : c->ess;              // historyelement::GetESS of the stored weights
            a.hist.temps[T] = temp;
            a.hist.repeats[T] = (double)reps;
            a.hist.accept[T] = c->accept_prob;
        }
        c->lse = 0.0; c->resampled = 0; c->parity = parity ^ 1; c->done = 0u;
        c->ticket = 0u;   // the resampling pass leaves its tile ticket advanced; k_la_cess reuses it as an arrival counter
        // arm the bisection for the next generation
        c->bis_done = 0; c->bis_phase = 0; c->bis_curr = temp; c->bis_desired = a.rho_n; c->bis_delta = 1.0 - temp;
    }
}

}  // namespace smcb
