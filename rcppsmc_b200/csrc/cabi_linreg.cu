// C ABI, model level: LinRegLA_adapt / LinRegLA (reference src/LinReg_LA_adapt.cpp:36-106, src/LinReg_LA.cpp:45-111).
#include <cmath>
#include <cstring>
#include <vector>

#include "cabi_common.cuh"
#include "la_engine.cuh"

using namespace smcb;
using namespace smcb_abi;

namespace {

// sampler.h:602-669 on the per-generation means / variances of the path-sampling integrand
void path_sampling(const std::vector<double>& E, const std::vector<double>& V, const std::vector<double>& temps, long L, double* out3) {
    long double rect = 0, trap = 0, trap2 = 0;
    for (long t = 1; t < L; ++t) {
        long double w = (long double)(temps[t] - temps[t - 1]);
        rect += (long double)E[t] * w;
        trap += w / 2.0 * ((long double)E[t - 1] + (long double)E[t]);
        trap2 += w / 2.0 * ((long double)E[t - 1] + (long double)E[t]) - powl(w, 2.0) / 12.0 * ((long double)V[t] - (long double)V[t - 1]);
    }
    out3[0] = (double)rect; out3[1] = (double)trap; out3[2] = (double)trap2;
}

long run_la(bool adaptive, const double* data_host, long n_obs, const double* temps_in, long L_in, long particles, double resampTol,
            double tempTol, uint64_t seed, const double* noise_host, long n_noise, const double* unif_host, long n_unif, long max_temps,
            double* theta, double* loglike, double* logprior, double* weights, double* ess, double* temps_out, double* repeats,
            double* lognc4, double* accept) {
    if (!data_host || n_obs < 1 || particles < 1) throw std::invalid_argument("LinRegLA: data, n_obs >= 1 and particles >= 1 required");
    if (particles > 2147483647L) throw std::invalid_argument("LinRegLA: particle count must be below 2^31");
    if (!adaptive && (!temps_in || L_in < 1)) throw std::invalid_argument("LinRegLA: a temperature schedule is required");
    if ((noise_host == nullptr) != (unif_host == nullptr)) throw std::invalid_argument("LinRegLA: inject both noise and uniforms or neither");
    require_device();
    const long long n = particles;
    const long lcap = adaptive ? max_temps : L_in;
    if (lcap < 1) throw std::invalid_argument("LinRegLA: max_temps must be positive");
    cudaStream_t st = 0;
    // ---- data: y and x - mean(x) (mean through Armadillo's two-accumulator sum, LinReg_LA_adapt.cpp:46)
    std::vector<double> y(data_host, data_host + n_obs), xc(n_obs);
    {
        double s0 = 0, s1 = 0; long i, j;
        const double* x = data_host + n_obs;
        for (i = 0, j = 1; j < n_obs; i += 2, j += 2) { s0 += x[i]; s1 += x[j]; }
        if (i < n_obs) s0 += x[i];
        const double mean_x = (s0 + s1) / (double)n_obs;
        for (long k = 0; k < n_obs; ++k) xc[k] = x[k] - mean_x;
    }
    DevBuf d_y(sizeof(double) * n_obs), d_xc(sizeof(double) * n_obs);
    SMCB_CUDA(cudaMemcpy(d_y.p, y.data(), sizeof(double) * n_obs, cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(d_xc.p, xc.data(), sizeof(double) * n_obs, cudaMemcpyHostToDevice));
    // ---- buffers
    DevBuf state(sizeof(double) * 10 * n), lw(sizeof(double) * n);
    DevBuf hist(sizeof(double) * 6 * (size_t)lcap * n), hs(sizeof(double) * 6 * lcap);
    const long long ntiles = scan_tiles(n);
    DevBuf zmask(sizeof(uint32_t) * ((n + 31) / 32 + 64)), zbase(sizeof(uint32_t) * ((n + 31) / 32 + 64)), surplus(sizeof(uint32_t) * n);
    DevBuf status(sizeof(unsigned long long) * 3 * ntiles), ctrl(sizeof(LaCtrl)), mean3(sizeof(double) * 4);
    int dev = 0, sms = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (n + 255) / 256;
    const int grid = (int)(want < 2LL * sms ? want : 2LL * sms);
    const int grid_mcmc = (int)((n + 127) / 128 < 8LL * sms ? (n + 127) / 128 : 8LL * sms);
    DevBuf partial(sizeof(double) * 16 * (size_t)(grid_mcmc > grid ? grid_mcmc : grid));
    DevBuf zin, uin;
    if (noise_host) {
        zin = DevBuf(sizeof(double) * n_noise); uin = DevBuf(sizeof(double) * n_unif);
        SMCB_CUDA(cudaMemcpy(zin.p, noise_host, sizeof(double) * n_noise, cudaMemcpyHostToDevice));
        SMCB_CUDA(cudaMemcpy(uin.p, unif_host, sizeof(double) * n_unif, cudaMemcpyHostToDevice));
    }
    LaArgs a{};
    for (int k = 0; k < 5; ++k) { a.buf.a[k] = state.as<double>() + (size_t)k * n; a.buf.b[k] = state.as<double>() + (size_t)(5 + k) * n; }
    a.buf.lw = lw.as<double>(); a.buf.n = n;
    a.data = LaData{d_y.as<double>(), d_xc.as<double>(), (int)n_obs};
    a.ctrl = ctrl.as<LaCtrl>();
    for (int k = 0; k < 6; ++k) a.hist.comp[k] = hist.as<double>() + (size_t)k * lcap * n;
    a.hist.ess = hs.as<double>(); a.hist.temps = hs.as<double>() + lcap; a.hist.repeats = hs.as<double>() + 2 * lcap;
    a.hist.ps_mean = hs.as<double>() + 3 * lcap; a.hist.ps_var = hs.as<double>() + 4 * lcap; a.hist.accept = hs.as<double>() + 5 * lcap;
    a.hist.lcap = lcap;
    a.am = AncestorMap{zmask.as<uint32_t>(), zbase.as<uint32_t>(), surplus.as<uint32_t>()};
    a.partial = partial.as<double>(); a.seed = seed; a.zin = zin.as<double>(); a.uin = uin.as<double>();
    a.adaptive = adaptive ? 1 : 0; a.data_annealing = 0; a.rho_n = tempTol * (double)n;

    LaCtrl h{};
    h.T = 0; h.temp_prev = 0.0; h.temp_cur = adaptive ? 0.0 : temps_in[0];
    h.thr = adaptive ? (resampTol < 1 ? resampTol * (double)n : resampTol) : 0.5 * (double)n;   // LinReg_LA.cpp:64 hard-codes 0.5
    h.log_n = std::log((double)n); h.accept_prob = -1.0;
    h.n_repeats = adaptive ? 0 : 10;                                                           // LinReg_LA.cpp:65 SetMcmcRepeats(10)
    if (!adaptive) {   // chol(covRW), LinReg_LA.cpp:35-36
        const double A[9] = {2500, -2.5, 0.03, -2.5, 130.0, 0.0, 0.03, 0.0, 0.04};
        double* U = h.chol;
        for (int k = 0; k < 9; ++k) U[k] = 0.0;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i <= j; ++i) {
                double s = A[i + 3 * j];
                for (int k = 0; k < i; ++k) s -= U[k + 3 * i] * U[k + 3 * j];
                U[i + 3 * j] = (i == j) ? std::sqrt(s) : s / U[i + 3 * i];
            }
    }
    SMCB_CUDA(cudaMemcpy(ctrl.p, &h, sizeof h, cudaMemcpyHostToDevice));

    ScanArgs sc{};
    sc.in = a.buf.lw; sc.n = n; sc.n_glob = n; sc.lse_ptr = &a.ctrl->lse; sc.enable_ptr = &a.ctrl->resampled; sc.u_ptr = &a.ctrl->u_res;
    sc.ticket = &a.ctrl->ticket;
    sc.ws = ScanWs{status.as<unsigned long long>(), status.as<unsigned long long>() + ntiles, status.as<unsigned long long>() + 2 * ntiles};
    sc.am = a.am;
    DevBuf cslot(sizeof(uint32_t) * resample_cslot_entries(n));
    sc.cslot = cslot.as<uint32_t>();

    long L = 0;
    for (long g = 0;; ++g) {
        if (g >= lcap) throw std::runtime_error("LinRegLA_adapt: more temperatures than max_temps");
        if (g == 0) {
            k_la_init<<<grid, 256, 0, st>>>(a);
        } else {
            if (adaptive) {
                for (int batch = 0;; ++batch) {   // CESS bisection: device state machine, a batch of launches per host check
                    for (int r = 0; r < 64; ++r) k_la_cess<<<grid, 256, 0, st>>>(a);
                    SMCB_CUDA(cudaGetLastError());
                    SMCB_CUDA(cudaMemcpy(&h, ctrl.p, sizeof h, cudaMemcpyDeviceToHost));
                    if (h.failed) throw std::runtime_error("Bisection method to choose the next temperature failed");   // staticModelAdapt.h:95
                    if (h.bis_done) break;
                    if (batch >= 7) {
                        char msg[512];
                        snprintf(msg, sizeof msg, "CESS bisection did not terminate after 512 evaluations (generation %ld, phase %d, a=%.17g b=%.17g m=%.17g "
                                 "fa=%g fb=%g delta=%g curr=%g desired=%g lse=%g ess=%g)", g, h.bis_phase, h.bis_a, h.bis_b, h.bis_m, h.bis_fa, h.bis_fb,
                                 h.bis_delta, h.bis_curr, h.bis_desired, h.lse, h.ess);
                        throw std::runtime_error(msg);
                    }
                }
            } else {
                double tp[2] = {temps_in[g - 1], temps_in[g]};
                SMCB_CUDA(cudaMemcpyAsync(&a.ctrl->temp_prev, tp, sizeof tp, cudaMemcpyHostToDevice, st));
            }
            k_la_reweight<<<grid, 256, 0, st>>>(a);
        }
        SMCB_CUDA(cudaGetLastError());
        if (adaptive) {
            k_la_cov<<<grid, 256, 0, st>>>(a, 0, mean3.as<double>());
            k_la_cov<<<grid, 256, 0, st>>>(a, 1, mean3.as<double>());
        }
        k_la_pre_resample<<<1, 1, 0, st>>>(a);
        SMCB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(unsigned long long) * 3 * ntiles, st));
        launch_resample_passes<SRC_LOGW, SYSTEMATIC>(sc, st);
        k_la_mcmc<<<grid_mcmc, 128, 0, st>>>(a);
        SMCB_CUDA(cudaGetLastError());
        L = g + 1;
        if (adaptive) { if (g > 0 && h.temp_cur == 1.0) break; }
        else if (L == L_in) break;
    }
    SMCB_CUDA(cudaStreamSynchronize(st));
    SMCB_CUDA(cudaMemcpy(&h, ctrl.p, sizeof h, cudaMemcpyDeviceToHost));
    // ---- outputs in the reference's shapes: theta cube [N x 3 x L], matrices [N x L] (column-major), vectors [L]
    for (long l = 0; l < L; ++l) {
        for (int cidx = 0; cidx < 3; ++cidx)
            if (theta) SMCB_CUDA(cudaMemcpy(theta + ((size_t)l * 3 + cidx) * n, a.hist.comp[cidx] + (size_t)l * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (loglike) SMCB_CUDA(cudaMemcpy(loglike + (size_t)l * n, a.hist.comp[3] + (size_t)l * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (logprior) SMCB_CUDA(cudaMemcpy(logprior + (size_t)l * n, a.hist.comp[4] + (size_t)l * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (weights) SMCB_CUDA(cudaMemcpy(weights + (size_t)l * n, a.hist.comp[5] + (size_t)l * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    }
    std::vector<double> sc6(6 * (size_t)lcap);
    SMCB_CUDA(cudaMemcpy(sc6.data(), hs.p, sizeof(double) * 6 * lcap, cudaMemcpyDeviceToHost));
    std::vector<double> E(L), V(L), tv(L);
    for (long l = 0; l < L; ++l) {
        if (ess) ess[l] = sc6[l];
        tv[l] = sc6[lcap + l];
        if (temps_out) temps_out[l] = tv[l];
        if (repeats) repeats[l] = sc6[2 * lcap + l];
        E[l] = sc6[3 * lcap + l]; V[l] = sc6[4 * lcap + l];
        if (accept) accept[l] = sc6[5 * lcap + l];
    }
    if (lognc4) { lognc4[0] = h.path; path_sampling(E, V, tv, L, lognc4 + 1); }
    return L;
}

// LinReg_impl (src/LinReg.cpp:46-84): data annealing, MULTINOMIAL / 0.5, 10 random-walk MH repeats per step with chol(covRW).
void run_linreg(const double* data_host, long n_obs, long particles, uint64_t seed, const double* noise_host, long n_noise,
                const double* unif_host, long n_unif, const double* draw_unif_host, double* theta, double* weights, double* lognc) {
    if (!data_host || n_obs < 1 || particles < 1 || particles > 2147483647L) throw std::invalid_argument("LinReg: data, n_obs >= 1 and 1 <= particles < 2^31 required");
    if ((noise_host == nullptr) != (unif_host == nullptr)) throw std::invalid_argument("LinReg: inject both noise and uniforms or neither");
    require_device();
    const long long n = particles;
    cudaStream_t st = 0;
    std::vector<double> y(data_host, data_host + n_obs), xc(n_obs);
    {
        double s0 = 0, s1 = 0; long i, j;
        const double* x = data_host + n_obs;
        for (i = 0, j = 1; j < n_obs; i += 2, j += 2) { s0 += x[i]; s1 += x[j]; }
        if (i < n_obs) s0 += x[i];
        const double mean_x = (s0 + s1) / (double)n_obs;
        for (long k = 0; k < n_obs; ++k) xc[k] = x[k] - mean_x;
    }
    DevBuf d_y(sizeof(double) * n_obs), d_xc(sizeof(double) * n_obs);
    SMCB_CUDA(cudaMemcpy(d_y.p, y.data(), sizeof(double) * n_obs, cudaMemcpyHostToDevice));
    SMCB_CUDA(cudaMemcpy(d_xc.p, xc.data(), sizeof(double) * n_obs, cudaMemcpyHostToDevice));
    DevBuf state(sizeof(double) * 10 * n), lw(sizeof(double) * n), hs(sizeof(double) * 8);
    const long long ntiles = scan_tiles(n);
    DevBuf zmask(sizeof(uint32_t) * ((n + 31) / 32 + 64)), zbase(sizeof(uint32_t) * ((n + 31) / 32 + 64)), surplus(sizeof(uint32_t) * n);
    DevBuf status(sizeof(unsigned long long) * 3 * ntiles), ctrl(sizeof(LaCtrl)), cum(sizeof(unsigned long long) * n), count(sizeof(int) * n);
    DevBuf extra(sizeof(unsigned long long) * 2), step(sizeof(long long));
    int dev = 0, sms = 0;
    SMCB_CUDA(cudaGetDevice(&dev));
    SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (n + 255) / 256;
    const int grid = (int)(want < 2LL * sms ? want : 2LL * sms);
    const int grid_mcmc = (int)((n + 127) / 128 < 8LL * sms ? (n + 127) / 128 : 8LL * sms);
    DevBuf partial(sizeof(double) * 16 * (size_t)(grid_mcmc > grid ? grid_mcmc : grid));
    DevBuf zin, uin, udraw;
    if (noise_host) {
        zin = DevBuf(sizeof(double) * n_noise); uin = DevBuf(sizeof(double) * n_unif);
        SMCB_CUDA(cudaMemcpy(zin.p, noise_host, sizeof(double) * n_noise, cudaMemcpyHostToDevice));
        SMCB_CUDA(cudaMemcpy(uin.p, unif_host, sizeof(double) * n_unif, cudaMemcpyHostToDevice));
    }
    if (draw_unif_host) {   // resampling uniforms, indexed by step: U[t * N + j]
        udraw = DevBuf(sizeof(double) * (size_t)n_obs * n);
        SMCB_CUDA(cudaMemcpy(udraw.p, draw_unif_host, sizeof(double) * (size_t)n_obs * n, cudaMemcpyHostToDevice));
    }
    LaArgs a{};
    for (int k = 0; k < 5; ++k) { a.buf.a[k] = state.as<double>() + (size_t)k * n; a.buf.b[k] = state.as<double>() + (size_t)(5 + k) * n; }
    a.buf.lw = lw.as<double>(); a.buf.n = n;
    a.data = LaData{d_y.as<double>(), d_xc.as<double>(), (int)n_obs};
    a.ctrl = ctrl.as<LaCtrl>();
    a.hist.lcap = 0;
    a.am = AncestorMap{zmask.as<uint32_t>(), zbase.as<uint32_t>(), surplus.as<uint32_t>()};
    a.partial = partial.as<double>(); a.seed = seed; a.zin = zin.as<double>(); a.uin = uin.as<double>();
    a.adaptive = 0; a.data_annealing = 1; a.rho_n = 0.0;
    LaCtrl h{};
    h.thr = 0.5 * (double)n; h.log_n = std::log((double)n); h.accept_prob = -1.0; h.n_repeats = 10;   // LinReg.cpp:62-63
    {
        const double A[9] = {2500, -2.5, 0.03, -2.5, 130.0, 0.0, 0.03, 0.0, 0.04};   // covRW, LinReg.cpp:34-35
        double* U = h.chol;
        for (int k = 0; k < 9; ++k) U[k] = 0.0;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i <= j; ++i) {
                double s = A[i + 3 * j];
                for (int k = 0; k < i; ++k) s -= U[k + 3 * i] * U[k + 3 * j];
                U[i + 3 * j] = (i == j) ? std::sqrt(s) : s / U[i + 3 * i];
            }
    }
    SMCB_CUDA(cudaMemcpy(ctrl.p, &h, sizeof h, cudaMemcpyHostToDevice));
    ScanWs ws{status.as<unsigned long long>(), status.as<unsigned long long>() + ntiles, status.as<unsigned long long>() + 2 * ntiles};
    DevBuf bucket(sizeof(uint32_t) * draw_bucket_entries(n));
    DrawWs dw{cum.as<unsigned long long>(), count.as<int>(), extra.as<unsigned long long>(), (unsigned int*)(extra.as<unsigned long long>() + 1), bucket.as<uint32_t>()};
    for (long g = 0; g < n_obs; ++g) {
        if (g == 0) k_la_init<<<grid, 256, 0, st>>>(a);
        else k_la_reweight<<<grid, 256, 0, st>>>(a);
        k_la_pre_resample<<<1, 1, 0, st>>>(a);
        SMCB_CUDA(cudaMemsetAsync(status.p, 0, sizeof(unsigned long long) * 3 * ntiles, st));
        SMCB_CUDA(cudaMemsetAsync(extra.p, 0, sizeof(unsigned long long) * 2, st));
        launch_resample_draws<SRC_LOGW>(MULTINOMIAL, a.buf.lw, n, &a.ctrl->lse, &a.ctrl->resampled, udraw.as<double>(), &a.ctrl->T, seed,
                                        &a.ctrl->ticket, ws, dw, a.am, nullptr, st);
        k_la_mcmc<<<grid_mcmc, 128, 0, st>>>(a);
        SMCB_CUDA(cudaGetLastError());
    }
    SMCB_CUDA(cudaStreamSynchronize(st));
    SMCB_CUDA(cudaMemcpy(&h, ctrl.p, sizeof h, cudaMemcpyDeviceToHost));
    double* const* cur = h.parity ? a.buf.b : a.buf.a;
    for (int cidx = 0; cidx < 3; ++cidx)
        if (theta) SMCB_CUDA(cudaMemcpy(theta + (size_t)cidx * n, cur[cidx], sizeof(double) * n, cudaMemcpyDeviceToHost));   // arma::mat [N x 3]
    if (weights) {
        std::vector<double> lwh(n);
        SMCB_CUDA(cudaMemcpy(lwh.data(), a.buf.lw, sizeof(double) * n, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < n; ++i) weights[i] = std::exp(lwh[i] - h.lse);   // GetParticleWeight()
    }
    if (lognc) *lognc = h.path;
}

}  // namespace

extern "C" int smcb200_LinReg(const double* data_host, long n_obs, long particles, uint64_t seed, const double* noise_host, long n_noise,
                              const double* uniforms_host, long n_uniforms, const double* resample_uniforms_host, double* theta_host,
                              double* weights_host, double* lognc_host) {
    return guarded([&]() -> int {
        run_linreg(data_host, n_obs, particles, seed, noise_host, n_noise, uniforms_host, n_uniforms, resample_uniforms_host, theta_host,
                   weights_host, lognc_host);
        return SMCB200_OK;
    });
}

extern "C" long smcb200_LinRegLA_adapt(const double* data_host, long n_obs, long particles, double resampTol, double tempTol,
                                       uint64_t seed, const double* noise_host, long n_noise, const double* uniforms_host, long n_uniforms,
                                       long max_temps, double* theta_host, double* loglike_host, double* logprior_host,
                                       double* weights_host, double* ess_host, double* temps_host, double* repeats_host,
                                       double* lognc4_host, double* accept_host) {
    long L = -1;
    int rc = guarded([&]() -> int {
        L = run_la(true, data_host, n_obs, nullptr, 0, particles, resampTol, tempTol, seed, noise_host, n_noise, uniforms_host, n_uniforms,
                   max_temps, theta_host, loglike_host, logprior_host, weights_host, ess_host, temps_host, repeats_host, lognc4_host, accept_host);
        return SMCB200_OK;
    });
    return rc == SMCB200_OK ? L : -(long)rc;
}

extern "C" int smcb200_LinRegLA(const double* data_host, long n_obs, const double* temps_host, long n_temps, long particles, uint64_t seed,
                                const double* noise_host, long n_noise, const double* uniforms_host, long n_uniforms, double* theta_host,
                                double* loglike_host, double* logprior_host, double* weights_host, double* ess_host, double* lognc4_host) {
    return guarded([&]() -> int {
        run_la(false, data_host, n_obs, temps_host, n_temps, particles, 0.5, 0.0, seed, noise_host, n_noise, uniforms_host, n_uniforms, n_temps,
               theta_host, loglike_host, logprior_host, weights_host, ess_host, nullptr, nullptr, lognc4_host, nullptr);
        return SMCB200_OK;
    });
}
