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
}
