// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/LinReg_LA.cpp (LinRegLA_impl: fixed temperature
// schedule, 10 random-walk MH repeats per step) included from $SMC_REF_DIR against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/LinReg_LA.cpp"

#include <cstring>

extern "C" {
// data: [n_obs x 2] column-major (y, x); temps[L].  Outputs as in smcref_linreg_la_adapt (no Temps / mcmcRepeats).
int smcref_linreg_la(const double* data, long n_obs, const double* intemps, long L, long N, double* theta, double* loglike,
                     double* logprior, double* weights, double* ess, double* lognc) {
    arma::mat D(n_obs, 2);
    for (long i = 0; i < 2 * n_obs; ++i) D.d[i] = data[i];
    arma::vec tv(L);
    for (long i = 0; i < L; ++i) tv(i) = intemps[i];
    Rcpp::List r = LinRegLA_impl(D, tv, (unsigned long)N);
    if (r.is_nil) return -1;
    auto put = [&](const char* name, double* dst) { const Rcpp::Column* c = r.find(name); if (dst && c) std::memcpy(dst, c->val.data(), sizeof(double) * c->val.size()); };
    put("theta", theta); put("loglike", loglike); put("logprior", logprior); put("Weights", weights); put("ESS", ess);
    const char* nc[4] = {"logNC_standard", "logNC_ps_rect", "logNC_ps_trap", "logNC_ps_trap2"};
    for (int k = 0; k < 4; ++k) lognc[k] = r.find(nc[k])->scalar;
    return 0;
}
}
