// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/LinReg_LA_adapt.cpp (LinRegLA_adapt_impl,
// rad_move, rad_adapt on top of smc::staticModelAdapt) included from $SMC_REF_DIR against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/LinReg_LA_adapt.cpp"

#include <cstring>

extern "C" {
// data: [n_obs x 2] column-major (y column, x column).  Outputs are sized by the caller for at most max_temps
// temperatures; returns the number of temperatures L (or -1 on failure / overflow).
// theta [L][3][N] (slice-major, column-major inside a slice), loglike/logprior/weights [L][N], ess/temps/repeats [L],
// lognc[4] = {standard, ps_rect, ps_trap, ps_trap2}.
long smcref_linreg_la_adapt(const double* data, long n_obs, long N, double resampTol, double tempTol, long max_temps,
                            double* theta, double* loglike, double* logprior, double* weights, double* ess,
                            double* temps, double* repeats, double* lognc) {
    arma::mat D(n_obs, 2);
    for (long i = 0; i < 2 * n_obs; ++i) D.d[i] = data[i];
    Rcpp::List r = LinRegLA_adapt_impl(D, (unsigned long)N, resampTol, tempTol);
    if (r.is_nil) return -1;
    const Rcpp::Column* t = r.find("Temps");
    long L = (long)t->val.size();
    if (L > max_temps) return -1;
    auto put = [&](const char* name, double* dst) { const Rcpp::Column* c = r.find(name); if (dst && c) std::memcpy(dst, c->val.data(), sizeof(double) * c->val.size()); };
    put("theta", theta); put("loglike", loglike); put("logprior", logprior); put("Weights", weights);
    put("ESS", ess); put("Temps", temps); put("mcmcRepeats", repeats);
    const char* nc[4] = {"logNC_standard", "logNC_ps_rect", "logNC_ps_trap", "logNC_ps_trap2"};
    for (int k = 0; k < 4; ++k) lognc[k] = r.find(nc[k])->scalar;
    return L;
}
}
