// A user model written OUTSIDE the engine, against the public plug-in header only: Bayesian logistic regression (two
// coefficients, Gaussian prior), sampled by likelihood tempering with adaptive temperatures (conditional ESS bisection),
// adaptive random-walk Metropolis-Hastings rejuvenation (empirical covariance -> Cholesky factor, adaptive repeat count)
// and path-sampling evidence estimates -- the same ingredients as the reference's LinReg_LA_adapt example
// (src/LinReg_LA_adapt.cpp, inst/include/LinReg_LA_adapt.h:55-99) on a different likelihood.
//
// The program checks the device run three ways and exits non-zero on any mismatch:
//   1. generation by generation against a CPU loop of the SAME __host__ __device__ functors (same Philox streams; the ancestor
//      indices of each resampling are taken from the AL history, the resampling itself has its own bit-exact tests);
//   2. logNC increments / ESS recomputed on the host from the log-weights the functors produce;
//   3. the evidence against a brute-force quadrature of the two-dimensional posterior (known answer).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tests/cpp/user_model.cu -o user_model
#include <smcb200/sampler.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NOBS = 40;

struct Logistic {
    static constexpr int D = 4;   // theta0, theta1, log-likelihood, log-prior
    struct Params {
        double temp_prev, temp_cur;
        double chol[4];           // upper factor of the proposal covariance, row-major 2x2
        double x[NOBS], y[NOBS];
    };
    SMCB_HD static double loglike(const double th0, const double th1, const Params& p) {
        double s = 0.0;
        for (int k = 0; k < NOBS; ++k) {
            const double eta = th0 + th1 * p.x[k];
            // log sigma(eta) for y = 1, log(1 - sigma(eta)) for y = 0, written without overflow
            const double a = p.y[k] > 0.5 ? -eta : eta;
            s -= a > 0.0 ? a + log1p(exp(-a)) : log1p(exp(a));
        }
        return s;
    }
    SMCB_HD static double logprior(const double th0, const double th1) {   // N(0, 5^2) x N(0, 5^2)
        return -(th0 * th0 + th1 * th1) / 50.0 - log(2.0 * 3.14159265358979323846 * 25.0);
    }
    SMCB_HD static void init(double (&x)[D], double& logw, const Params& p, smc::rng& rng) {
        x[0] = 5.0 * rng.normal();
        x[1] = 5.0 * rng.normal();
        x[2] = loglike(x[0], x[1], p);
        x[3] = logprior(x[0], x[1]);
        logw = p.temp_cur * x[2];
    }
    SMCB_HD static void move(long, double (&x)[D], double& logw, const Params& p, smc::rng&) {
        logw += (p.temp_cur - p.temp_prev) * x[2];
    }
    SMCB_HD static bool mcmc(long, double (&x)[D], double&, const Params& p, smc::rng& rng) {
        const double z0 = rng.normal(), z1 = rng.normal(), u = rng.uniform();
        const double c0 = x[0] + p.chol[0] * z0, c1 = x[1] + p.chol[1] * z0 + p.chol[3] * z1;   // theta + U^T z
        const double ll = loglike(c0, c1, p), lp = logprior(c0, c1);
        if (exp(p.temp_cur * (ll - x[2]) + lp - x[3]) > u) { x[0] = c0; x[1] = c1; x[2] = ll; x[3] = lp; return true; }
        return false;
    }
};

struct LogisticAdapt : smc::adaptMethods<Logistic> {
    smc::staticModelAdapt<Logistic, 2, 2> sma;
    double rho;
    std::vector<Logistic::Params> at_move, at_mcmc;   // what the functors of each generation saw (for the CPU replay)
    std::vector<int> repeats;
    explicit LogisticAdapt(double rho_) : rho(rho_) {}
    void updateForMove(Logistic::Params& p, const smc::population<Logistic>& pop) override {
        sma.ChooseTemp(pop, rho * pop.GetNumber());
        p.temp_prev = sma.GetTempPrevious(); p.temp_cur = sma.GetTempCurr();
        at_move.push_back(p);
    }
    void updateForMCMC(Logistic::Params& p, const smc::population<Logistic>& pop, double acceptProb, int nResampled, int& nRepeats) override {
        if (nResampled) {
            nRepeats = sma.calcMcmcRepeats(acceptProb, 0.99, 5, 50);
            sma.calcCholCov(pop);
            p.chol[0] = sma.GetCholCov(0, 0); p.chol[1] = sma.GetCholCov(0, 1); p.chol[2] = 0.0; p.chol[3] = sma.GetCholCov(1, 1);
        } else {
            nRepeats = 0;
        }
        at_mcmc.push_back(p);
        repeats.push_back(nRepeats);
    }
};

struct LoglikeIntegrand { __device__ double operator()(long, const double (&x)[4]) const { return x[2]; } };
struct Theta0 { __device__ double operator()(const double (&x)[4]) const { return x[0]; } };
struct Theta1 { __device__ double operator()(const double (&x)[4]) const { return x[1]; } };

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++g_fail; fprintf(stderr, "CHECK FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static double host_lse(const std::vector<double>& v, double* ess = nullptr) {
    double m = -INFINITY;
    for (double x : v) m = std::max(m, x);
    long double s = 0, q = 0;
    for (double x : v) { const long double e = std::exp((long double)(x - m)); s += e; q += e * e; }
    if (ess) *ess = (double)(s * s / q);
    return m + (double)std::log(s);
}

int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : 16384;
    const unsigned long long seed = 20261020ull;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { fprintf(stderr, "user_model: no CUDA device (the engine has no CPU fallback)\n"); return 12; }

    Logistic::Params prm{};
    prm.temp_prev = 0.0; prm.temp_cur = 0.0;
    prm.chol[0] = prm.chol[3] = 1.0;
    for (int k = 0; k < NOBS; ++k) {
        prm.x[k] = -2.0 + 4.0 * k / (NOBS - 1.0);
        const double pr = 1.0 / (1.0 + std::exp(-(0.5 + 1.5 * prm.x[k])));
        prm.y[k] = (std::fmod(0.37 + 0.618033988749895 * (k + 1), 1.0) < pr) ? 1.0 : 0.0;   // fixed pseudo-random labels
    }

    smc::sampler<Logistic> s(N, HistoryType::AL, seed);
    LogisticAdapt adapt(0.9);
    s.SetResampleParams(ResampleType::SYSTEMATIC, 0.5);
    s.SetAdaptMethods(&adapt);
    s.SetAlgParam(prm);
    s.SetMcmcRepeats(0);
    s.Initialise();
    std::vector<double> ess_seen{-1.0}, lognc_step{s.GetLogNCStep()};
    while (adapt.sma.GetTempCurr() != 1.0) {
        ess_seen.push_back(s.IterateEss());
        lognc_step.push_back(s.GetLogNCStep());
        if (s.GetTime() > 200) { fprintf(stderr, "temperature schedule did not reach 1\n"); return 1; }
    }
    const long L = s.GetHistoryLength();
    printf("generations %ld, temperatures:", L);
    for (double t : adapt.sma.GetTemps()) printf(" %.4f", t);
    printf("\nlogNC %.6f\n", s.GetLogNCPath());
    CHECK(L == s.GetTime() + 1 && L == (long)adapt.sma.GetTemps().size(), "history length %ld, time %ld", L, s.GetTime());

    // ---- 1/2: CPU replay of every generation with the same functors
    std::vector<double> x[4], lw(N);
    for (int d = 0; d < 4; ++d) x[d].resize(N);
    double max_dx = 0.0, max_dlw = 0.0;
    long accepted_total_mismatch = 0;
    for (long g = 0; g < L; ++g) {
        const Logistic::Params& pm = g == 0 ? prm : adapt.at_move[g - 1];
        // pfInitialise / pfMove
        if (g == 0) {
            for (long i = 0; i < N; ++i) {
                smc::rng r(seed, smcb::STREAM_INIT, 0u, (uint64_t)i);
                double xi[4] = {0, 0, 0, 0}, w = 0.0;
                Logistic::init(xi, w, pm, r);
                for (int d = 0; d < 4; ++d) x[d][i] = xi[d];
                lw[i] = w - std::log((double)N);
            }
        } else {
            for (long i = 0; i < N; ++i) {
                smc::rng r(seed, smcb::STREAM_MOVE, (uint32_t)g, (uint64_t)i);
                double xi[4] = {x[0][i], x[1][i], x[2][i], x[3][i]};
                Logistic::move(g, xi, lw[i], pm, r);
                for (int d = 0; d < 4; ++d) x[d][i] = xi[d];
            }
        }
        double ess;
        const double inc = host_lse(lw, &ess);
        CHECK(std::fabs(inc - lognc_step[g]) <= 1e-9 * std::max(1.0, std::fabs(inc)), "generation %ld: logNC increment host %.12g device %.12g", g, inc, lognc_step[g]);
        if (g > 0) CHECK(std::fabs(ess - ess_seen[g]) <= 1e-6 * ess, "generation %ld: ESS host %.9g device %.9g", g, ess, ess_seen[g]);
        for (double& v : lw) v -= inc;
        const int resampled = s.GetHistoryFlags(g).resampled;
        CHECK(resampled == (ess < 0.5 * N ? 1 : 0), "generation %ld: resampling decision", g);
        // Resample: ancestors from the AL history (checked for the defining property of systematic resampling: every parent's
        // children count is within one of N w)
        if (resampled) {
            const std::vector<unsigned int> anc = s.GetHistoryAncestors(g);
            std::vector<int> cnt(N, 0);
            for (long i = 0; i < N; ++i) { CHECK(anc[i] < (unsigned)N, "ancestor out of range"); ++cnt[anc[i]]; }
            long bad = 0;
            for (long i = 0; i < N; ++i) if (std::fabs(cnt[i] - N * std::exp(lw[i])) > 1.0 + 1e-6) ++bad;
            CHECK(bad == 0, "generation %ld: %ld parents with a children count more than one away from N w", g, bad);
            std::vector<double> y[4];
            for (int d = 0; d < 4; ++d) { y[d].resize(N); for (long i = 0; i < N; ++i) y[d][i] = x[d][anc[i]]; x[d].swap(y[d]); }
            for (double& v : lw) v = -std::log((double)N);
        }
        // DoMCMC
        const int reps = adapt.repeats[g];
        CHECK(reps == s.GetHistorymcmcRepeats(g), "generation %ld: repeats %d vs history %d", g, reps, s.GetHistorymcmcRepeats(g));
        long acc = 0;
        if (reps > 0) {
            const Logistic::Params& pc = adapt.at_mcmc[g];
            for (long i = 0; i < N; ++i) {
                smc::rng r(seed, smcb::STREAM_MCMC, (uint32_t)g, (uint64_t)i);
                double xi[4] = {x[0][i], x[1][i], x[2][i], x[3][i]};
                for (int j = 0; j < reps; ++j) acc += Logistic::mcmc(g, xi, lw[i], pc, r) ? 1 : 0;
                for (int d = 0; d < 4; ++d) x[d][i] = xi[d];
            }
        }
        // compare with the generation the device stored; a host/device difference in the last ulps of a proposal can flip an
        // accept/reject decision of a particle sitting exactly on the boundary, so a handful of outliers are tolerated and counted
        long outliers = 0;
        for (int d = 0; d < 4; ++d) {
            const std::vector<double> dv = s.GetHistoryValues(g, d);
            for (long i = 0; i < N; ++i) {
                const double e = std::fabs(dv[i] - x[d][i]) / std::max(1.0, std::fabs(x[d][i]));
                if (e > 1e-9) { ++outliers; x[d][i] = dv[i]; } else max_dx = std::max(max_dx, e);
            }
        }
        CHECK(outliers <= 8, "generation %ld: %ld state components differ from the CPU replay", g, outliers);
        const std::vector<double> dl = s.GetHistoryLogWeight(g);
        for (long i = 0; i < N; ++i) max_dlw = std::max(max_dlw, std::fabs(dl[i] - lw[i]));
        accepted_total_mismatch += std::labs(acc - (long)s.GetHistoryAccepted(g));
    }
    CHECK(max_dlw <= 1e-9, "log-weights differ from the CPU replay by %.3g", max_dlw);
    CHECK(accepted_total_mismatch <= 8, "acceptance counts differ by %ld in total", accepted_total_mismatch);
    printf("CPU replay: max state error %.3g, max log-weight error %.3g, acceptance-count mismatch %ld\n", max_dx, max_dlw, accepted_total_mismatch);

    // ---- 3: evidence by quadrature (trapezoid on a fine grid; the posterior is concentrated well inside [-12, 12]^2)
    double logZ;
    {
        const int G = 1201; const double lo = -12.0, h = 24.0 / (G - 1);
        std::vector<double> terms; terms.reserve((size_t)G * G);
        for (int a = 0; a < G; ++a) for (int b = 0; b < G; ++b) {
            const double t0 = lo + a * h, t1 = lo + b * h;
            const double wq = ((a == 0 || a == G - 1) ? 0.5 : 1.0) * ((b == 0 || b == G - 1) ? 0.5 : 1.0);
            terms.push_back(Logistic::loglike(t0, t1, prm) + Logistic::logprior(t0, t1) + std::log(wq * h * h));
        }
        logZ = host_lse(terms);
    }
    const double ps_rect = s.IntegratePathSampling(PathSamplingType::RECTANGLE, LoglikeIntegrand{}, [&](long t) { return adapt.sma.GetTemp((int)t) - adapt.sma.GetTemp((int)t - 1); });
    const double ps_trap = s.IntegratePathSampling(PathSamplingType::TRAPEZOID1, LoglikeIntegrand{}, [&](long t) { return adapt.sma.GetTemp((int)t) - adapt.sma.GetTemp((int)t - 1); });
    const double ps_trap2 = s.IntegratePathSampling(LoglikeIntegrand{}, [&](long t) { return adapt.sma.GetTemp((int)t) - adapt.sma.GetTemp((int)t - 1); });
    printf("log evidence: quadrature %.5f, SMC %.5f, path sampling rect %.5f trap %.5f trap2 %.5f\n", logZ, s.GetLogNCPath(), ps_rect, ps_trap, ps_trap2);
    const double tol = 6.0 / std::sqrt((double)N) + 0.02;
    CHECK(std::fabs(s.GetLogNCPath() - logZ) < tol, "SMC evidence %.5f vs quadrature %.5f", s.GetLogNCPath(), logZ);
    CHECK(std::fabs(ps_trap2 - logZ) < 0.1, "path-sampling (trapezoid2) evidence %.5f vs quadrature %.5f", ps_trap2, logZ);
    CHECK(std::fabs(ps_trap - logZ) < 0.3 && std::fabs(ps_rect - logZ) < 1.5, "path-sampling rect/trap estimates off: %.4f %.4f", ps_rect, ps_trap);
    printf("posterior mean theta: %.4f %.4f (Integrate), ESS now %.1f\n", s.Integrate(Theta0{}), s.Integrate(Theta1{}), s.GetESS());

    // ---- IterateBack (sampler.h:684-699) restores the previous generation
    {
        const long T = s.GetTime();
        const std::vector<double> before = s.GetHistoryValues(L - 2, 0);
        s.IterateBack();
        CHECK(s.GetTime() == T - 1 && s.GetHistoryLength() == L - 1, "IterateBack bookkeeping");
        const std::vector<double> now = s.GetParticleValues(0);
        long diff = 0;
        for (long i = 0; i < N; ++i) diff += now[i] != before[i];
        CHECK(diff == 0, "IterateBack: %ld particle values differ from the stored generation", diff);
        const std::vector<unsigned int> line = s.GetALineInd(N / 2);
        CHECK((long)line.size() == L - 1 && line.back() == (unsigned)(N / 2), "GetALineInd shape");
        const auto sp = s.GetALineSpace(N / 2);
        CHECK(sp.size() == line.size() && sp[0].size() == 4, "GetALineSpace shape");
    }
    if (g_fail) { fprintf(stderr, "user_model: %d check(s) failed\n", g_fail); return 1; }
    printf("user_model OK\n");
    return 0;
}
