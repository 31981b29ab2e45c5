// Model functors for the device engine.  Each is the device restatement of one reference moveset
// (user pfInitialise / pfMove virtuals, moveset.h:73-87) plus the integrands its driver feeds to
// sampler::Integrate.  Functor contract (all static, __device__):
//   D         doubles per particle state (stored as D SoA arrays)
//   NZ_INIT   standard normals consumed by init(), NZ_MOVE by move() -- handed in as z[]
//   init(x, lw, params, z)            pfInitialise: sets x and the initial log-weight
//   move(t, x, lw, params, z)         pfMove at lTime = t: updates x, adds the incremental log-weight
//   NSHIFT / shift_stat(x, h)         statistics whose pre-resample weighted means centre the fused moments
//   NOBS / observe(x, shift, f)       integrands accumulated under the current weights (Integrate)
//   finish_obs(v, shift, out)         weighted means v[] -> reported values
// Arithmetic is written in the reference's operation order with contraction disabled (-fmad=false) so that,
// on injected noise, states and log-weight increments are bit-identical to the CPU reference (only exp/log
// in the normalisation differ by ulps).
#pragma once
#include "common.cuh"

namespace smcb {

// Nonlinear state-space model of Gordon, Salmond & Smith (1993): reference src/pfnonlinbs.cpp:35-120.
struct NonlinBS {
    static constexpr int D = 1;
    static constexpr int NZ_INIT = 1;
    static constexpr int NZ_MOVE = 1;
    static constexpr int NSHIFT = 1;
    static constexpr int NOBS = 2;
    struct Params {
        const double* y;      // observations y[0..T)
        const double* cterm;  // 8 cos(1.2 t), particle-independent term of pfMove, tabulated on the host
    };
    // src/pfnonlinbs.cpp:35-42
    static constexpr double var_x0 = 10.0, var_x = 10.0, var_y = 1.0, scale_y = 1.0 / 20.0;

    __device__ static __forceinline__ double loglik(const Params& p, long long t, double x) {  // :86-88
        double d = p.y[t] - x * x * scale_y;
        return -0.5 * (d * d) / var_y;
    }
    __device__ static __forceinline__ void init(double (&x)[D], double& lw, const Params& p, const double* z) {  // :95-98
        const double std_x0 = 3.1622776601683795;  // sqrt(10.0)
        x[0] = 0.0 + std_x0 * z[0];
        lw = loglik(p, 0, x[0]);
    }
    __device__ static __forceinline__ void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) {  // :106-109
        const double std_x = 3.1622776601683795;
        double v = x[0];
        v = 0.5 * v + fm_div(25.0 * v, 1.0 + v * v) + p.cterm[t] + (0.0 + std_x * z[0]);   // 1 + v^2 >= 1: the branch-free correctly rounded quotient
        x[0] = v;
        lw += loglik(p, t, v);
    }
    __device__ static __forceinline__ void shift_stat(const double (&x)[D], double (&h)[NSHIFT]) { h[0] = x[0]; }
    // integrand_mean_x / integrand_var_x (:112-120), centred on `shift` instead of the not-yet-known mean
    __device__ static __forceinline__ void observe(const double (&x)[D], const double* shift, double (&f)[NOBS]) {
        double d = x[0] - shift[0];
        f[0] = d; f[1] = d * d;
    }
    __device__ static __forceinline__ void finish_obs(const double* v, const double* shift, double* out) {
        out[0] = shift[0] + v[0];        // mean
        out[1] = v[1] - v[0] * v[0];     // E[(x - mean)^2]; the driver reports its square root
    }
    __device__ static __forceinline__ void load_dynamic(Params&, const double*) {}   // nothing travels through the control block
};


// Constant-velocity tracking with Student-t observation noise: reference src/pflineart.cpp:32-169,
// inst/include/pflineart.h:28-60.  State (x_pos, y_pos, x_vel, y_vel) in that SoA order.
struct Lineart {
    static constexpr int D = 4;
    static constexpr int NZ_INIT = 4;
    static constexpr int NZ_MOVE = 4;
    static constexpr int NSHIFT = 2;
    static constexpr int NOBS = 4;
    struct Params {
        const double* yx;  // observed x positions y.x_pos[0..T)
        const double* yy;  // observed y positions
    };
    static constexpr double scale_y = 0.1, nu_y = 10.0, Delta = 0.1;  // src/pflineart.cpp:38-40

    __device__ static __forceinline__ double loglik(const Params& p, long long t, const double (&x)[D]) {  // :134-137
        double ax = (x[0] - p.yx[t]) / scale_y;
        double ay = (x[1] - p.yy[t]) / scale_y;
        return (-0.5 * (nu_y + 1.0)) * (log(1 + (ax * ax) / nu_y) + log(1 + (ay * ay) / nu_y));
    }
    __device__ static __forceinline__ void init(double (&x)[D], double& lw, const Params& p, const double* z) {  // :144-152
        x[0] = 0.0 + 2.0 * z[0];   // sqrt(var_s0) = 2
        x[1] = 0.0 + 2.0 * z[1];
        x[2] = 0.0 + 1.0 * z[2];   // sqrt(var_u0) = 1
        x[3] = 0.0 + 1.0 * z[3];
        lw = loglik(p, 0, x);
    }
    __device__ static __forceinline__ void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) {  // :161-169
        const double sd_s = 0.1414213562373095;    // sqrt(var_s = 0.02)
        const double sd_u = 0.03162277660168379;   // sqrt(var_u = 0.001)
        x[0] += x[2] * Delta + (0.0 + sd_s * z[0]);
        x[2] += (0.0 + sd_u * z[1]);
        x[1] += x[3] * Delta + (0.0 + sd_s * z[2]);
        x[3] += (0.0 + sd_u * z[3]);
        lw += loglik(p, t, x);
    }
    __device__ static __forceinline__ void shift_stat(const double (&x)[D], double (&h)[NSHIFT]) { h[0] = x[0]; h[1] = x[1]; }
    // integrand_mean_x / var_x / mean_y / var_y (:103-128), centred on the pre-resample weighted means
    __device__ static __forceinline__ void observe(const double (&x)[D], const double* shift, double (&f)[NOBS]) {
        double dx = x[0] - shift[0], dy = x[1] - shift[1];
        f[0] = dx; f[1] = dx * dx; f[2] = dy; f[3] = dy * dy;
    }
    __device__ static __forceinline__ void finish_obs(const double* v, const double* shift, double* out) {
        out[0] = shift[0] + v[0]; out[1] = v[1] - v[0] * v[0];   // Xm, Xv (variance, the reference takes no square root: :90-93)
        out[2] = shift[1] + v[2]; out[3] = v[3] - v[2] * v[2];   // Ym, Yv
    }
    __device__ static __forceinline__ void load_dynamic(Params&, const double*) {}   // nothing travels through the control block
};


// Inner bootstrap filter of the particle-marginal Metropolis-Hastings example: reference src/nonLinPMMH.cpp:155-190.
// Same nonlinear model as NonlinBS with (sigv, sigw) as parameters, x0 ~ N(0, 5), cos(1.2 (t+1)) and fully normalised
// Gaussian log-weights (R::dnorm(..., log = TRUE)).  Only the log normalising constant is needed by the outer chain.
struct NonlinPMMH {
    static constexpr int D = 1;
    static constexpr int NZ_INIT = 1;
    static constexpr int NZ_MOVE = 1;
    static constexpr int NSHIFT = 0;
    static constexpr int NOBS = 0;
    struct Params {
        const double* y;
        const double* cterm;   // 8 cos(1.2 (t + 1)) for t = 0..T-1, tabulated on the host
        double sigv, sigw;
        double log_sigw;       // log(sigw) from the host (NaN for a negative proposal: the chain then rejects)
    };
    __device__ static __forceinline__ double dnorm_log(double x, double mu, const Params& p) {  // R nmath dnorm4, give_log
        const double ln_sqrt_2pi = 0.918938533204672741780329736406;
        double z = (x - mu) / p.sigw;
        return -(ln_sqrt_2pi + 0.5 * z * z + p.log_sigw);
    }
    __device__ static __forceinline__ void init(double (&x)[D], double& lw, const Params& p, const double* z) {  // :168-173
        const double sd0 = 2.23606797749979;  // sqrt(5.0)
        x[0] = 0.0 + sd0 * z[0];
        lw = dnorm_log(p.y[0], (x[0] * x[0]) / 20.0, p);
    }
    __device__ static __forceinline__ void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) {  // :181-186
        double v = x[0];
        v = v / 2.0 + fm_div(25.0 * v, 1 + v * v) + p.cterm[t] + (0.0 + p.sigv * z[0]);
        x[0] = v;
        lw += dnorm_log(p.y[t], (v * v) / 20.0, p);
    }
    __device__ static __forceinline__ void shift_stat(const double (&x)[D], double (&h)[1]) { h[0] = 0.0; }
    __device__ static __forceinline__ void observe(const double (&x)[D], const double* shift, double (&f)[1]) { f[0] = 0.0; }
    __device__ static __forceinline__ void finish_obs(const double* v, const double* shift, double* out) {}
    // (sigv, sigw, log sigw) of the current proposal arrive through the control block, so one captured graph of the filter
    // serves every MH iteration
    __device__ static __forceinline__ void load_dynamic(Params& p, const double* m) { p.sigv = m[0]; p.sigw = m[1]; p.log_sigw = m[2]; }
};

// Scalar linear-Gaussian state space model of the conditional-SMC example: reference src/cSMCexamples.cpp:141-200.
// x_0 = stateInit (deterministic), x_t = phi x_{t-1} + N(0, varStateEvol), y_t ~ N(x_t, varObs); pfInitialise already moves
// x_0 to x_1 and weights it with y[0] (:165-172).
struct LGSS {
    static constexpr int D = 1;
    static constexpr int NZ_INIT = 1;
    static constexpr int NZ_MOVE = 1;
    static constexpr int NSHIFT = 0;
    static constexpr int NOBS = 0;
    struct Params {
        const double* y;
        double phi, sd_evol, x0, sd_obs, log_sd_obs;
    };
    __host__ __device__ static __forceinline__ double loglik(const Params& p, double yt, double x) {  // R nmath dnorm4, give_log (:152-156)
        const double ln_sqrt_2pi = 0.918938533204672741780329736406;
        double z = (yt - x) / p.sd_obs;
        return -(ln_sqrt_2pi + 0.5 * z * z + p.log_sd_obs);
    }
    __device__ static __forceinline__ void init(double (&x)[D], double& lw, const Params& p, const double* z) {
        x[0] = p.phi * p.x0 + (0.0 + p.sd_evol * z[0]);
        lw = loglik(p, p.y[0], x[0]);
    }
    __device__ static __forceinline__ void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) {
        x[0] = p.phi * x[0] + (0.0 + p.sd_evol * z[0]);
        lw += loglik(p, p.y[t], x[0]);
    }
    // pfWeight of the conditional move (src/cSMCexamples.cpp:203-210): the reference particle's value has just been pinned
    __device__ static __forceinline__ void weight(long long t, const double (&x)[D], double& lw, const Params& p) { lw += loglik(p, p.y[t], x[0]); }
    __device__ static __forceinline__ void shift_stat(const double (&x)[D], double (&h)[1]) { h[0] = 0.0; }
    __device__ static __forceinline__ void observe(const double (&x)[D], const double* shift, double (&f)[1]) { f[0] = 0.0; }
    __device__ static __forceinline__ void finish_obs(const double* v, const double* shift, double* out) {}
    __device__ static __forceinline__ void load_dynamic(Params&, const double*) {}   // nothing travels through the control block
};


}  // namespace smcb
