// C ABI, model level: pfLineartBS (reference src/pflineart.cpp:53-99).
#include <cmath>
#include <memory>
#include <vector>

#include "cabi_common.cuh"
#include "models.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {
struct LineartFilter {
    long long n = 0, tcap = 0;
    std::unique_ptr<Sampler<Lineart>> s;
    DevBuf yx, yy, noise, unif;
};
}  // namespace

extern "C" int smcb200_pfLineartBS(const double* data_host, long T, long particles, int resample_mode, double threshold,
                                   uint64_t seed, const double* noise_host, const double* uniforms_host,
                                   smcb200_step_callback callback, void* user, double* Xm_host, double* Xv_host,
                                   double* Ym_host, double* Yv_host, double* lognc_host, double* ess_host, int* resampled_host) {
    return guarded([&]() -> int {
        if (!data_host || T < 1 || particles < 1) throw std::invalid_argument("smcb200_pfLineartBS: data, T >= 1 and particles >= 1 required");
        if (resample_mode < 0 || resample_mode > 3) throw std::invalid_argument("smcb200_pfLineartBS: unknown resampling mode");
        require_device();
        thread_local std::unique_ptr<LineartFilter> cache;   // engine reused while the shape matches
        if (cache && (cache->n != particles || cache->tcap < T)) cache.reset();
        if (!cache) {
            cache.reset(new LineartFilter);
            cache->n = particles; cache->tcap = T;
            cache->s.reset(new Sampler<Lineart>(particles, T, seed, 0));
            cache->yx = DevBuf(sizeof(double) * T);
            cache->yy = DevBuf(sizeof(double) * T);
        }
        LineartFilter& f = *cache;
        // data is the reference's arma::mat [T x 2], column-major: x column then y column (src/pflineart.cpp:61-62)
        SMCB_CUDA(cudaMemcpy(f.yx.p, data_host, sizeof(double) * T, cudaMemcpyHostToDevice));
        SMCB_CUDA(cudaMemcpy(f.yy.p, data_host + T, sizeof(double) * T, cudaMemcpyHostToDevice));
        const double* zn = nullptr; const double* un = nullptr;
        if (noise_host) {
            size_t bytes = sizeof(double) * 4 * (size_t)particles * (size_t)T;
            f.noise = DevBuf(bytes);
            SMCB_CUDA(cudaMemcpy(f.noise.p, noise_host, bytes, cudaMemcpyHostToDevice));
            zn = f.noise.as<double>();
        }
        if (uniforms_host) {
            size_t cnt = resample_mode == SMCB200_SYSTEMATIC ? (size_t)T
                       : resample_mode == SMCB200_STRATIFIED ? (size_t)T * (size_t)(particles + 1) : (size_t)T * (size_t)particles;
            f.unif = DevBuf(sizeof(double) * cnt);
            SMCB_CUDA(cudaMemcpy(f.unif.p, uniforms_host, sizeof(double) * cnt, cudaMemcpyHostToDevice));
            un = f.unif.as<double>();
        }
        Sampler<Lineart>& s = *f.s;
        s.SetParams(Lineart::Params{f.yx.as<double>(), f.yy.as<double>()});
        s.SetSeed(seed);
        s.SetResampleParams(resample_mode, threshold);
        s.SetInjectedNoise(zn, resample_mode == SMCB200_SYSTEMATIC ? un : nullptr, resample_mode == SMCB200_STRATIFIED ? un : nullptr,
                           (resample_mode == SMCB200_MULTINOMIAL || resample_mode == SMCB200_RESIDUAL) ? un : nullptr);
        std::vector<double> obs(4 * (size_t)T, 0.0), xm(T, 0.0), ym(T, 0.0);
        s.Initialise();
        for (long t = 1; t < T; ++t) {
            s.Iterate();
            if (callback) {  // the reference calls fun(Xm, Ym) after every Iterate (src/pflineart.cpp:85): costs one extra moment pass
                s.ObserveFinal();
                s.ReadTraces(t + 1, nullptr, nullptr, nullptr, obs.data());
                for (long k = 0; k <= t; ++k) { xm[k] = obs[4 * k]; ym[k] = obs[4 * k + 2]; }
                callback(xm.data(), ym.data(), T, user);
            }
        }
        s.ObserveFinal();
        s.ReadTraces(T, lognc_host, ess_host, resampled_host, obs.data());
        for (long t = 0; t < T; ++t) {
            if (Xm_host) Xm_host[t] = obs[4 * t];
            if (Xv_host) Xv_host[t] = obs[4 * t + 1];
            if (Ym_host) Ym_host[t] = obs[4 * t + 2];
            if (Yv_host) Yv_host[t] = obs[4 * t + 3];
        }
        return SMCB200_OK;
    });
}
