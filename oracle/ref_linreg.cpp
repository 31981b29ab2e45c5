// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/LinReg.cpp (LinReg_impl: data annealing)
// included from $SMC_REF_DIR against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/LinReg.cpp"

#include <cstring>

extern "C" int smcref_linreg(const double* data, long n_obs, long N, double* theta, double* weights, double* lognc) {
    arma::mat D(n_obs, 2);
    for (long i = 0; i < 2 * n_obs; ++i) D.d[i] = data[i];
    Rcpp::List r = LinReg_impl(D, (unsigned long)N);
    if (r.is_nil) return -1;
    const Rcpp::Column* t = r.find("theta"); const Rcpp::Column* w = r.find("weights");
    if (theta) std::memcpy(theta, t->val.data(), sizeof(double) * t->val.size());
    if (weights) std::memcpy(weights, w->val.data(), sizeof(double) * w->val.size());
    *lognc = r.find("logNC")->scalar;
    return 0;
}
