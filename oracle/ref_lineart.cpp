// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// Second translation unit of oracle/_ref/libsmcref.so: the UNMODIFIED reference src/pflineart.cpp (pfLineartBS_impl,
// pflineart_move, integrands) included from $SMC_REF_DIR against oracle/shim/RcppArmadillo.h.
#include "RcppArmadillo.h"

#include "src/pflineart.cpp"

extern "C" {

// The reference driver verbatim (RESIDUAL, 0.5; src/pflineart.cpp:53-99); data = [T x 2] column-major.
int smcref_pflineart_default(const double* data, long T, long N, double* Xm, double* Xv, double* Ym, double* Yv) {
    arma::mat d(T, 2);
    for (long i = 0; i < 2 * T; ++i) d.d[i] = data[i];
    Rcpp::DataFrame r = pfLineartBS_impl(d, (unsigned long)N, false, Rcpp::Function());
    const char* names[4] = {"Xm", "Xv", "Ym", "Yv"};
    double* outs[4] = {Xm, Xv, Ym, Yv};
    for (int k = 0; k < 4; ++k) { const Rcpp::Column* c = r.find(names[k]); for (long t = 0; t < T; ++t) outs[k][t] = c->val[t]; }
    return 0;
}

// Same loop as the reference driver (:56-88) with mode / threshold as arguments and the per-step sampler scalars.
int smcref_pflineart(const double* data, long T, long N, int mode, double threshold, double* Xm, double* Xv, double* Ym,
                     double* Yv, double* lognc, double* ess, int* resampled) {
    pflineart::y.x_pos = arma::vec(T); pflineart::y.y_pos = arma::vec(T);
    for (long t = 0; t < T; ++t) { pflineart::y.x_pos(t) = data[t]; pflineart::y.y_pos(t) = data[T + t]; }
    pflineart::myMove = new pflineart::pflineart_move;
    {
        smc::sampler<pflineart::cv_state, smc::nullParams> Sampler(N, HistoryType::NONE, pflineart::myMove);
        Sampler.SetResampleParams((ResampleType::Enum)mode, threshold);
        Sampler.Initialise();
        for (long n = 0; n < T; ++n) {
            double e = -1.0;
            if (n > 0) e = Sampler.IterateEss();
            double xm = Sampler.Integrate(pflineart::integrand_mean_x, NULL);
            double xv = Sampler.Integrate(pflineart::integrand_var_x, (void*)&xm);
            double ym = Sampler.Integrate(pflineart::integrand_mean_y, NULL);
            double yv = Sampler.Integrate(pflineart::integrand_var_y, (void*)&ym);
            if (Xm) Xm[n] = xm; if (Xv) Xv[n] = xv; if (Ym) Ym[n] = ym; if (Yv) Yv[n] = yv;
            if (lognc) lognc[n] = Sampler.GetLogNCPath();
            if (ess) ess[n] = e;
            if (resampled) resampled[n] = Sampler.GetResampled();
        }
    }
    delete pflineart::myMove;
    return 0;
}

}  // extern "C"
