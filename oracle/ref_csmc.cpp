// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/cSMCexamples.cpp (compareNCestimates_imp:
// bootstrap filter vs conditional bootstrap filter on a scalar linear-Gaussian model, inst/include/conditionalSampler.h)
// included from $SMC_REF_DIR against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/cSMCexamples.cpp"

extern "C" {
// data[T]; refTraj [T x simNum] column-major; outputs smc/csmc [simNum x 4] column-major (multinomial, residual, stratified, systematic)
int smcref_compare_nc(const double* data, long T, long N, int simNum, double phi, double varEvol, double stateInit0, double varObs0,
                      const double* refTraj, double* smcOut, double* csmcOut) {
    arma::vec d(T);
    for (long t = 0; t < T; ++t) d(t) = data[t];
    Rcpp::List par;
    par["phi"] = phi; par["varStateEvol"] = varEvol; par["stateInit"] = stateInit0; par["varObs"] = varObs0;
    arma::mat ref(T, simNum);
    for (long i = 0; i < T * simNum; ++i) ref.d[i] = refTraj[i];
    Rcpp::List r = compareNCestimates_imp(d, N, simNum, par, ref);
    if (r.is_nil) return -1;
    const Rcpp::Column* a = r.find("smcOut"); const Rcpp::Column* b = r.find("csmcOut");
    for (long i = 0; i < 4L * simNum; ++i) { smcOut[i] = a->val[i]; csmcOut[i] = b->val[i]; }
    return 0;
}

// ---- one smc::conditionalSampler object driven run by run (the driver above keeps ONE object for every simulation and mode, so
// referenceTrajectoryIndices carries over between runs; this surface makes that state observable).
namespace {
struct Chain {
    smc::moveset<States, smc::nullParams>* mv = nullptr;
    smc::conditionalSampler<States, smc::nullParams>* cs = nullptr;
    smc::sampler<States, smc::nullParams>* bpf = nullptr;
    long T = 0, N = 0;
} g_chain;
}
void smcref_csmc_close(void) {
    delete g_chain.cs; delete g_chain.bpf; delete g_chain.mv;
    g_chain = Chain();
}
int smcref_csmc_open(const double* data, long T, long N, double phi, double varEvol, double stateInit0, double varObs0) {
    smcref_csmc_close();
    arma::vec d(T);
    for (long t = 0; t < T; ++t) d(t) = data[t];
    y.yObs = d; params.statePhi = phi; params.stateVarEvol = varEvol; stateInit = stateInit0; varObs = varObs0;
    g_chain.mv = new cSMCexamples_move;
    std::vector<States> ref0(T);
    for (long t = 0; t < T; ++t) ref0[t].xState = 0.0;
    g_chain.cs = new smc::conditionalSampler<States, smc::nullParams>(N, HistoryType::AL, g_chain.mv, ref0);
    g_chain.bpf = new smc::sampler<States, smc::nullParams>(N, HistoryType::NONE, g_chain.mv);
    g_chain.T = T; g_chain.N = N;
    return 0;
}
// One conditional run, as one block of the driver (:101-105): SetResampleParams(mode, 0.5); SetReferenceTrajectory; Initialise;
// IterateUntil(T-1).  Outputs: logNC path, referenceTrajectoryIndices [T] after the run, and the AL history: ancestors [T][N],
// values [T][N], log-weights [T][N], resampled flags [T], plus GetALineInd(n) for every n as lines [T][N] and GetALineSpace as
// line_values [T][N].
int smcref_csmc_run(int mode, const double* ref, double* lognc, unsigned* refidx, unsigned* anc, double* values, double* logw,
                    int* resampled, unsigned* lines, double* line_values) {
    try {
        const long T = g_chain.T, N = g_chain.N;
        std::vector<States> r(T);
        for (long t = 0; t < T; ++t) r[t].xState = ref[t];
        auto& cs = *g_chain.cs;
        cs.SetResampleParams((ResampleType::Enum)mode, 0.5);
        cs.SetReferenceTrajectory(r);
        cs.Initialise();
        cs.IterateUntil(T - 1);
        if (lognc) *lognc = cs.GetLogNCPath();
        if (refidx) for (long t = 0; t < T; ++t) refidx[t] = (unsigned)cs.GetReferenceValueIndex(t);
        const auto& H = cs.GetHistory();
        if ((long)H.size() != T) return -2;
        for (long t = 0; t < T; ++t) {
            arma::Col<unsigned int> a = H[t].GetAIndices();
            const smc::population<States> pop = cs.GetHistoryPopulation(t);   // the Refs getter returns a dangling reference
            for (long i = 0; i < N; ++i) {
                if (anc) anc[t * N + i] = a(i);
                if (values) values[t * N + i] = pop.GetValueN(i).xState;
                if (logw) logw[t * N + i] = pop.GetLogWeightN(i);
            }
            if (resampled) { smc::historyelement<States> he = H[t]; resampled[t] = he.WasResampled() ? 1 : 0; }
        }
        if (lines || line_values)
            for (long n = 0; n < N; ++n) {
                arma::Col<unsigned int> L = cs.GetALineInd(n);
                std::vector<States> S = cs.GetALineSpace(n);
                for (long t = 0; t < T; ++t) {
                    if (lines) lines[t * N + n] = L(t);
                    if (line_values) line_values[t * N + n] = S[t].xState;
                }
            }
        return 0;
    } catch (smc::exception e) { return -1; }
}
// One plain bootstrap-filter run as in the driver (:82-85).
int smcref_bpf_lgss_run(int mode, double* lognc) {
    try {
        auto& s = *g_chain.bpf;
        s.SetResampleParams((ResampleType::Enum)mode, 0.5);
        s.Initialise();
        s.IterateUntil(g_chain.T - 1);
        *lognc = s.GetLogNCPath();
        return 0;
    } catch (smc::exception e) { return -1; }
}
}
