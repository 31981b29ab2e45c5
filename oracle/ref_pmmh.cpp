// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/nonLinPMMH.cpp included from $SMC_REF_DIR
// against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/nonLinPMMH.cpp"

extern "C" {
// The reference driver verbatim (MULTINOMIAL, 0.5).  Outputs have M+1 entries.
int smcref_nonlin_pmmh(const double* y, long T, long N, long M, double* sigv, double* sigw, double* loglike, double* logprior, double* acc) {
    arma::vec data(T);
    for (long t = 0; t < T; ++t) data(t) = y[t];
    std::ostringstream sink; std::streambuf* old = std::cout.rdbuf(sink.rdbuf());   // progress prints go to Rcout
    Rcpp::DataFrame r = nonLinPMMH_impl(data, (unsigned long)N, (unsigned long)M, false, 100);
    std::cout.rdbuf(old);
    if (r.is_nil) return -1;
    const char* names[5] = {"samples_sigv", "samples_sigw", "loglike", "logprior", "MH_acceptance_rate"};
    double* outs[5] = {sigv, sigw, loglike, logprior, acc};
    for (int k = 0; k < 5; ++k) { const Rcpp::Column* c = r.find(names[k]); for (long i = 0; i <= M; ++i) outs[k][i] = c->val[i]; }
    return 0;
}
// Inner filter through the reference's own sampler + nonLinPMMH_move with a chosen resampling mode (the driver hard-codes
// MULTINOMIAL, whose rmultinom is not uniform-reproducible): Initialise(); IterateUntil(T-1); GetLogNCPath().
double smcref_pmmh_filter(const double* yobs, long T, long N, double sigv, double sigw, int mode, double threshold) {
    nonLinPMMH::y = arma::vec(T);
    for (long t = 0; t < T; ++t) nonLinPMMH::y(t) = yobs[t];
    nonLinPMMH::theta_prop.sigv = sigv; nonLinPMMH::theta_prop.sigw = sigw;
    nonLinPMMH::myMove = new nonLinPMMH::nonLinPMMH_move;
    double out;
    {
        smc::sampler<double, smc::nullParams> Sampler(N, HistoryType::NONE, nonLinPMMH::myMove);
        Sampler.SetResampleParams((ResampleType::Enum)mode, threshold);
        Sampler.Initialise();
        Sampler.IterateUntil(T - 1);
        out = Sampler.GetLogNCPath();
    }
    delete nonLinPMMH::myMove;
    return out;
}
}
