// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/blockpfgaussianopt.cpp (block-sampling particle
// filter with path-valued particles, systematic resampling at ESS < N/2) included from $SMC_REF_DIR against
// oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/blockpfgaussianopt.cpp"

extern "C" {
// data[T]; outputs weight[N], values [N x T] column-major, logNC[1]
int smcref_blockpf(const double* data, long T, long N, long lag, double* weight, double* values, double* lognc) {
    arma::vec d(T);
    for (long t = 0; t < T; ++t) d(t) = data[t];
    Rcpp::List r = blockpfGaussianOpt_impl(d, N, lag);
    if (r.is_nil) return -1;
    const Rcpp::Column* w = r.find("weight"); const Rcpp::Column* v = r.find("values"); const Rcpp::Column* c = r.find("logNC");
    for (long i = 0; i < N; ++i) weight[i] = w->val[i];
    for (long i = 0; i < N * T; ++i) values[i] = v->val[i];
    *lognc = c->val[0];
    return 0;
}
}
