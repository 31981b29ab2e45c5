// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Builds the UNMODIFIED reference (RcppSMC 0.2.9) into oracle/_ref/libsmcref.so:
// this translation unit #includes the reference's own src/pfnonlinbs.cpp (and, through
// it, inst/include/{smctc,sampler,population,moveset,history,...}.h) straight from
// $SMC_REF_DIR (default /root/reference) against oracle/shim/RcppArmadillo.h.
// No reference source is copied into this repository.
//
// The extern "C" surface below is what tests/ and bench.py (cpu_baseline /
// --impl reference) bind with ctypes.  It is never linked into the product.
#include "RcppArmadillo.h"

#include "src/pfnonlinbs.cpp"  // reference: pfNonlinBS_impl, nonlinbs::nonlinbs_move, integrands

#include <chrono>
#include <cstdint>
#include <cstring>

namespace {
// Exposes the protected particle store so a resampling call can be driven on given log-weights.
struct OpenSampler : smc::sampler<double, smc::nullParams> {
    using smc::sampler<double, smc::nullParams>::sampler;
    void set_population(const double* val, const double* logw, long n) {
        std::vector<double> v(val, val + n);
        arma::vec lw(n);
        for (long i = 0; i < n; ++i) lw(i) = logw[i];
        pPopulation = smc::population<double>(v, lw);
    }
    const smc::population<double>& pop() const { return pPopulation; }
};
}  // namespace

extern "C" {

// ---- randomness injection ---------------------------------------------------
void smcref_seed(uint64_t seed) { shim::st().eng.seed(seed); shim::st().nd.reset(); }
void smcref_clear_queues(void) {
    shim::State& s = shim::st();
    s.uq.clear(); s.zq.clear(); s.mq.clear(); s.exp_override = nullptr; s.exp_override_n = 0;
    s.n_unif = s.n_norm = s.n_multinom = 0;
}
void smcref_push_uniforms(const double* u, long n) { for (long i = 0; i < n; ++i) shim::st().uq.push_back(u[i]); }
void smcref_push_normals(const double* z, long n) { for (long i = 0; i < n; ++i) shim::st().zq.push_back(z[i]); }
void smcref_push_multinomial(const int* c, long k) { shim::st().mq.emplace_back(c, c + k); }
long smcref_pending_uniforms(void) { return (long)shim::st().uq.size(); }
long smcref_pending_normals(void) { return (long)shim::st().zq.size(); }
long smcref_consumed_uniforms(void) { return shim::st().n_unif; }
long smcref_consumed_normals(void) { return shim::st().n_norm; }
// After `skip` further arma::exp() calls over length-n vectors, the next one returns w[] verbatim.  Used to feed
// the reference's Resample identical normalised weights: in `cumsum(exp(lw - stableLogSumWeights(lw)))`
// (sampler.h:811,833) the first exp belongs to stableLogSumWeights (skip = 1).  The pointer must stay valid.
void smcref_override_exp(const double* w, long n, int skip) {
    shim::st().exp_override = w; shim::st().exp_override_n = n; shim::st().exp_override_skip = skip;
}

// ---- engine primitives (reference population.h:46-50, sampler.h:413-417, :779-868) ------------
double smcref_lse(const double* logw, long n) {
    arma::vec lw(n);
    for (long i = 0; i < n; ++i) lw(i) = logw[i];
    return smc::stableLogSumWeights(lw);
}

// ESS exactly as sampler::GetESS on the given (not necessarily normalised) log-weights.
double smcref_ess(const double* logw, long n) {
    OpenSampler s(n, HistoryType::NONE);
    std::vector<double> v(n, 0.0);
    s.set_population(v.data(), logw, n);
    return s.GetESS();
}

// Runs sampler::Resample(mode) on a population with values 0..n-1 and the given log-weights.
// idx[] receives uRSIndices, val[] (optional) the replicated values, returns 0.
int smcref_resample(int mode, const double* logw, long n, unsigned* idx, double* val) {
    OpenSampler s(n, HistoryType::NONE);
    std::vector<double> v(n);
    for (long i = 0; i < n; ++i) v[i] = (double)i;
    s.set_population(v.data(), logw, n);
    s.Resample((ResampleType::Enum)mode);
    for (long i = 0; i < n; ++i) idx[i] = s.GetuRSIndex(i);
    if (val) for (long i = 0; i < n; ++i) val[i] = s.pop().GetValueN(i);
    return 0;
}

// ---- pfNonlinBS ---------------------------------------------------------------------------------
// The reference driver verbatim (MULTINOMIAL, threshold 1.01 N; src/pfnonlinbs.cpp:52-79).
int smcref_pfnonlinbs_default(const double* y, long T, long N, double* mean, double* sd) {
    arma::vec data(T);
    for (long t = 0; t < T; ++t) data(t) = y[t];
    Rcpp::List r = pfNonlinBS_impl(data, N);
    const Rcpp::Column* m = r.find("mean");
    const Rcpp::Column* s = r.find("sd");
    for (long t = 0; t < T; ++t) { mean[t] = m->val[t]; sd[t] = s->val[t]; }
    return 0;
}

// Same loop as the reference driver (src/pfnonlinbs.cpp:56-75) with the resampling mode / threshold
// as arguments and the per-step scalars the reference computes but does not return.
// Outputs (each length T, any may be NULL): mean, sd, logNC path, ESS (pre-resample; for t=0 it is
// recomputed from the weights since Initialise() does not return it), resampled flag.
int smcref_pfnonlinbs(const double* yobs, long T, long N, int mode, double threshold,
                      double* mean, double* sd, double* lognc, double* ess, int* resampled) {
    nonlinbs::y = arma::vec(T);
    for (long t = 0; t < T; ++t) nonlinbs::y(t) = yobs[t];
    nonlinbs::myMove = new nonlinbs::nonlinbs_move;
    {
        smc::sampler<double, smc::nullParams> Sampler(N, HistoryType::NONE, nonlinbs::myMove);
        Sampler.SetResampleParams((ResampleType::Enum)mode, threshold);
        Sampler.Initialise();
        for (long n = 0; n < T; ++n) {
            double e = -1.0;
            if (n > 0) e = Sampler.IterateEss();
            double m = Sampler.Integrate(nonlinbs::integrand_mean_x, NULL);
            double v = Sampler.Integrate(nonlinbs::integrand_var_x, (void*)&m);
            if (mean) mean[n] = m;
            if (sd) sd[n] = std::sqrt(v);
            if (lognc) lognc[n] = Sampler.GetLogNCPath();
            if (ess) ess[n] = e;
            if (resampled) resampled[n] = Sampler.GetResampled();
        }
    }
    delete nonlinbs::myMove;
    return 0;
}

// Wall-clock of the reference filter (used as the CPU baseline): returns seconds.
double smcref_time_pfnonlinbs(const double* yobs, long T, long N, int mode, double threshold) {
    std::vector<double> mean(T), sd(T);
    auto t0 = std::chrono::steady_clock::now();
    smcref_pfnonlinbs(yobs, T, N, mode, threshold, mean.data(), sd.data(), NULL, NULL, NULL);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
