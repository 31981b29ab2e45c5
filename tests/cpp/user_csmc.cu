// Out-of-tree user of smc::conditionalSampler<Model> (include/smcb200/sampler.cuh): the linear-Gaussian state-space model of the
// reference's conditional-SMC example (src/cSMCexamples.cpp:141-218) written as a USER model against the public header only, run as a
// chain of conditional bootstrap filters with all four resampling schemes, plus a two-dimensional model that exercises D > 1.
// Prints, per run: logNC, the carried reference-slot indices, a checksum of the AL history (ancestors and values) and one ancestral
// line; tests/test_user_csmc.py compares the LGSS lines with smcb200_conditionalBPF (which is pinned to the reference's fixtures).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <smcb200/sampler.cuh>

struct UserLgss {   // x_t = phi x_{t-1} + N(0, sd_evol^2), x_0 ~ phi x0 + N(0, sd_evol^2); y_t ~ N(x_t, sd_obs^2)
    static constexpr int D = 1, NZ_INIT = 1, NZ_MOVE = 1;
    struct Params { const double* y; double phi, sd_evol, x0, sd_obs, log_sd_obs; };
    __device__ static double loglik(const Params& p, double yt, double x) {
        const double z = (yt - x) / p.sd_obs;
        return -(0.918938533204672741780329736406 + 0.5 * z * z + p.log_sd_obs);
    }
    __device__ static void init(double (&x)[D], double& lw, const Params& p, const double* z) { x[0] = p.phi * p.x0 + (0.0 + p.sd_evol * z[0]); lw = loglik(p, p.y[0], x[0]); }
    __device__ static void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) { x[0] = p.phi * x[0] + (0.0 + p.sd_evol * z[0]); lw += loglik(p, p.y[t], x[0]); }
    __device__ static void weight(long long t, const double (&x)[D], double& lw, const Params& p) { lw += loglik(p, p.y[t], x[0]); }
};

struct UserTwoD {   // position + velocity random walk, the position is observed
    static constexpr int D = 2, NZ_INIT = 2, NZ_MOVE = 2;
    struct Params { const double* y; };
    __device__ static double loglik(const Params& p, long long t, const double (&x)[D]) { const double d = p.y[t] - x[0]; return -0.5 * d * d; }
    __device__ static void init(double (&x)[D], double& lw, const Params& p, const double* z) { x[0] = z[0]; x[1] = 0.3 * z[1]; lw = loglik(p, 0, x); }
    __device__ static void move(long long t, double (&x)[D], double& lw, const Params& p, const double* z) {
        x[0] = x[0] + x[1] + 0.5 * z[0]; x[1] = 0.9 * x[1] + 0.3 * z[1]; lw += loglik(p, t, x);
    }
    __device__ static void weight(long long t, const double (&x)[D], double& lw, const Params& p) { lw += loglik(p, t, x); }
};

int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : 1000, T = argc > 2 ? atol(argv[2]) : 40;
    const unsigned long long seed = argc > 3 ? strtoull(argv[3], nullptr, 10) : 7ull;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "user_csmc: no CUDA device\n"); return 12; }
    // data and reference trajectories: deterministic, reproduced by the test
    std::vector<double> y((size_t)T);
    for (long t = 0; t < T; ++t) y[(size_t)t] = std::sin(0.3 * (double)t) + 0.1 * (double)((t * 7) % 5);
    double* d_y = nullptr;
    cudaMalloc(&d_y, sizeof(double) * (size_t)T);
    cudaMemcpy(d_y, y.data(), sizeof(double) * (size_t)T, cudaMemcpyHostToDevice);
    try {
        {
            std::vector<std::array<double, 1>> ref((size_t)T);
            for (long t = 0; t < T; ++t) ref[(size_t)t] = {0.8 * std::sin(0.3 * (double)t)};
            smc::conditionalSampler<UserLgss> s(N, HistoryType::AL, ref, seed);
            s.SetAlgParam(UserLgss::Params{d_y, 0.9, 1.0, 0.0, 1.0, 0.0});
            for (int run = 0; run < 8; ++run) {   // two passes over the four schemes on ONE object: the reference-slot indices carry over
                s.SetResampleParams((ResampleType::Enum)(run % 4), 0.5);
                for (long t = 0; t < T; ++t) ref[(size_t)t][0] = 0.8 * std::sin(0.3 * (double)t) + 0.05 * (double)run;
                s.Initialise(ref);
                s.IterateUntil(T - 1);
                unsigned long long ck = 1469598103934665603ull;
                double vs = 0.0;
                for (long t = 0; t < T; ++t) {
                    for (uint32_t a : s.GetHistoryAncestors(t)) ck = (ck ^ a) * 1099511628211ull;
                    for (double v : s.GetHistoryValues(t)) vs += v;
                }
                printf("lgss run %d mode %d logNC %.17g refidx", run, run % 4, s.GetLogNCPath());
                for (long t = 0; t < T; ++t) printf(" %ld", s.GetReferenceValueIndex(t));
                printf(" anc_fnv %llu val_sum %.17g line", ck, vs);
                for (uint32_t a : s.GetALineInd(N / 3)) printf(" %u", a);
                printf("\n");
            }
            bool threw = false;
            try { s.SetMcmcRepeats(3); } catch (const std::logic_error&) { threw = true; }
            if (!threw) { fprintf(stderr, "SetMcmcRepeats must throw\n"); return 1; }
        }
        {
            std::vector<std::array<double, 2>> ref((size_t)T);
            for (long t = 0; t < T; ++t) ref[(size_t)t] = {std::sin(0.3 * (double)t), 0.1};
            smc::conditionalSampler<UserTwoD> s(N, HistoryType::AL, ref, seed);
            s.SetAlgParam(UserTwoD::Params{d_y});
            for (int mode = 0; mode < 4; ++mode) {
                s.SetResampleParams((ResampleType::Enum)mode, 0.5);
                s.Initialise();
                s.IterateUntil(T - 1);
                // the reference particle's line must be the reference trajectory itself: follow the slot indices through the history
                int bad = 0;
                for (long t = 0; t < T; ++t) {
                    const std::vector<double> v = s.GetHistoryValues(t);
                    // generation t holds the post-resampling population: the reference value of time t sits in the slot the NEXT step pins
                    const long k = t + 1 < T ? s.GetReferenceValueIndex(t + 1) : -1;
                    // (checked for MULTINOMIAL only: the other schemes reproduce the reference as it is -- RESIDUAL connects the drawn
                    // slot to the reference particle only when that particle has no deterministic offspring, conditionalSampler.h:629-633,
                    // and STRATIFIED / SYSTEMATIC never update referenceTrajectoryIndices, so a later step pins whatever slot an earlier
                    // run left there)
                    if (mode == 0 && k >= 0 && (v[(size_t)k] != ref[(size_t)t][0] || v[(size_t)N + (size_t)k] != ref[(size_t)t][1])) ++bad;
                }
                double ess_last = s.GetESS(T - 1);
                printf("twod mode %d logNC %.12g ess_last %.6g pinned_mismatches %d\n", mode, s.GetLogNCPath(), ess_last, bad);
                if (bad || !(ess_last >= 1.0 && ess_last <= (double)N) || !std::isfinite(s.GetLogNCPath())) { fprintf(stderr, "twod check failed\n"); return 1; }
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "user_csmc: %s\n", e.what());
        return 1;
    }
    printf("user_csmc OK\n");
    return 0;
}
