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
#include "csmc.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {

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
    CsmcArgs<LGSS> a{};
    a.thr = 0.5 * (double)n; a.run0 = 0; a.keep_refidx = 0; a.t_cap = T;   // SetResampleParams(mode, 0.5), a fresh sampler object per chain
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
    k_csmc_chain<LGSS><<<(unsigned)n_chains, CS_THREADS, 0, cs>>>(a);
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
