// C ABI, model level: nonLinPMMH (reference src/nonLinPMMH.cpp:37-153): random-walk Metropolis-Hastings over (sigv, sigw)
// with a bootstrap particle filter estimate of the likelihood per iteration.  The outer chain is host code (strictly
// sequential, one scalar read per iteration); the filter is the device engine.  Independent chains = independent calls,
// one per GPU / process: no communication.
#include <cmath>
#include <memory>
#include <vector>

#include "cabi_common.cuh"
#include "models.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {
double log_prior(double sigv, double sigw) {   // nonLinPMMH.cpp:160-163, IG(0.01, 0.01)-form prior on the s.d.s themselves
    const double a = 0.01, b = 0.01;
    return 2 * a * std::log(b) - 2 * std::lgamma(a) - (a + 1) * std::log(sigv) - b / sigv - (a + 1) * std::log(sigw) - b / sigw;
}
}  // namespace

extern "C" int smcb200_nonLinPMMH(const double* data_host, long T, long particles, long iterations, uint64_t seed,
                                  int resample_mode, double threshold, const double* innov_v_host, const double* innov_w_host,
                                  const double* unif_host, const double* filter_noise_host, const double* filter_unif_host,
                                  smcb200_pmmh_progress progress, long msg_freq, void* user, double* samples_sigv_host,
                                  double* samples_sigw_host, double* loglike_host, double* logprior_host, double* accept_rate_host) {
    return guarded([&]() -> int {
        if (!data_host || T < 1 || particles < 1 || iterations < 0) throw std::invalid_argument("smcb200_nonLinPMMH: data, T >= 1, particles >= 1, iterations >= 0 required");
        if (resample_mode < 0 || resample_mode > 3) throw std::invalid_argument("smcb200_nonLinPMMH: unknown resampling mode");
        if (!samples_sigv_host || !samples_sigw_host || !loglike_host || !logprior_host) throw std::invalid_argument("smcb200_nonLinPMMH: output arrays required");
        require_device();
        const long M = iterations;
        Sampler<NonlinPMMH> s(particles, T, seed, 0);
        DevBuf y(sizeof(double) * T), cterm(sizeof(double) * T), noise, unif;
        std::vector<double> c(T);
        for (long t = 0; t < T; ++t) c[t] = 8 * std::cos(1.2 * (double)(t + 1));   // nonLinPMMH.cpp:183 with lTime = t
        SMCB_CUDA(cudaMemcpy(y.p, data_host, sizeof(double) * T, cudaMemcpyHostToDevice));
        SMCB_CUDA(cudaMemcpy(cterm.p, c.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
        const size_t per_noise = (size_t)particles * (size_t)T;
        const size_t per_unif = resample_mode == SMCB200_SYSTEMATIC ? (size_t)T : resample_mode == SMCB200_STRATIFIED ? (size_t)T * (size_t)(particles + 1) : per_noise;
        if (filter_noise_host) noise = DevBuf(sizeof(double) * per_noise);
        if (filter_unif_host) unif = DevBuf(sizeof(double) * per_unif);
        s.SetResampleParams(resample_mode, threshold);

        auto filter_loglik = [&](long it, double sigv, double sigw) -> double {   // Initialise(); IterateUntil(T-1); GetLogNCPath()
            NonlinPMMH::Params p{y.as<double>(), cterm.as<double>(), sigv, sigw, std::log(sigw)};
            s.SetParams(p);
            const double dyn[3] = {sigv, sigw, std::log(sigw)};   // the same three values, as seen by a replayed graph
            s.SetDynamicParams(dyn, 3);
            s.SetSeed(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(it + 1));
            const double* zn = nullptr; const double* un = nullptr;
            if (filter_noise_host) { SMCB_CUDA(cudaMemcpy(noise.p, filter_noise_host + (size_t)it * per_noise, sizeof(double) * per_noise, cudaMemcpyHostToDevice)); zn = noise.as<double>(); }
            if (filter_unif_host) { SMCB_CUDA(cudaMemcpy(unif.p, filter_unif_host + (size_t)it * per_unif, sizeof(double) * per_unif, cudaMemcpyHostToDevice)); un = unif.as<double>(); }
            s.SetInjectedNoise(zn, resample_mode == SMCB200_SYSTEMATIC ? un : nullptr, resample_mode == SMCB200_STRATIFIED ? un : nullptr,
                               (resample_mode == SMCB200_MULTINOMIAL || resample_mode == SMCB200_RESIDUAL) ? un : nullptr);
            // graph replay pays off after ~60 filters (capture + instantiation of a 4 T-node graph); injected draws are re-uploaded per filter: eager
            s.RunFilter(T, /*use_graph=*/!filter_noise_host && !filter_unif_host && M >= 64);
            return s.ReadCtrl().lognc_path;
        };
        // outer-chain randomness: Rcpp::rnorm(M, 0, 0.15), Rcpp::rnorm(M, 0, 0.08), Rcpp::runif(M) (nonLinPMMH.cpp:72-74)
        std::vector<double> iv(M), iw(M), uu(M);
        for (long i = 0; i < M; ++i) {
            if (innov_v_host) { iv[i] = innov_v_host[i]; iw[i] = innov_w_host[i]; uu[i] = unif_host[i]; }
            else {
                Draw128 d = philox_draw(seed, STREAM_USER, 0u, (uint64_t)i, 0), e = philox_draw(seed, STREAM_USER, 0u, (uint64_t)i, 1);
                double u1 = u64_to_open01(d.a), u2 = u64_to_unit(d.b);
                double r = std::sqrt(-2.0 * std::log(u1)), ang = 6.283185307179586 * u2;
                iv[i] = 0 + 0.15 * (r * std::cos(ang)); iw[i] = 0 + 0.08 * (r * std::sin(ang));
                uu[i] = u64_to_unit(e.a);
            }
        }
        double sigv = 10.0, sigw = 10.0;                       // :64-68
        samples_sigv_host[0] = sigv; samples_sigw_host[0] = sigw;
        loglike_host[0] = filter_loglik(0, sigv, sigw);
        logprior_host[0] = log_prior(sigv, sigw);
        if (accept_rate_host) accept_rate_host[0] = 0.0;
        long accepted = 0;
        double sum_v = sigv, sum_w = sigw;
        for (long i = 1; i <= M; ++i) {
            const double pv = samples_sigv_host[i - 1] + iv[i - 1], pw = samples_sigw_host[i - 1] + iw[i - 1];   // :88-89
            const double ll_prop = filter_loglik(i, pv, pw);
            const double lp_prop = log_prior(pv, pw);
            const double ratio = std::exp(ll_prop - loglike_host[i - 1] + lp_prop - logprior_host[i - 1]);
            if (ratio > uu[i - 1]) {   // NaN (negative proposals) compares false: reject
                samples_sigv_host[i] = pv; samples_sigw_host[i] = pw; loglike_host[i] = ll_prop; logprior_host[i] = lp_prop;
                ++accepted;
            } else {
                samples_sigv_host[i] = samples_sigv_host[i - 1]; samples_sigw_host[i] = samples_sigw_host[i - 1];
                loglike_host[i] = loglike_host[i - 1]; logprior_host[i] = logprior_host[i - 1];
            }
            if (accept_rate_host) accept_rate_host[i] = (double)accepted / (double)i;
            if (progress && msg_freq > 0 && i % msg_freq == 0)   // the reference prints to Rcpp::Rcout every msg_freq iterations (:116-138)
                progress(i, M, sum_v / (double)i, sum_w / (double)i, loglike_host[i], logprior_host[i], (double)accepted / (double)i, user);
            sum_v += samples_sigv_host[i]; sum_w += samples_sigw_host[i];
        }
        return SMCB200_OK;
    });
}
