// Rcpp glue for the RcppSMC package: the eight [[Rcpp::export]] bodies with the numeric work forwarded to libsmcb200.so
// through include/smcb200.hpp.  A maintainer drops this file into src/ in place of the *_impl bodies of
//   src/pfnonlinbs.cpp:52, src/pflineart.cpp:53, src/LinReg.cpp:46, src/LinReg_LA.cpp:45, src/LinReg_LA_adapt.cpp:36,
//   src/blockpfgaussianopt.cpp:41, src/nonLinPMMH.cpp:37, src/cSMCexamples.cpp:36
// and adds to src/Makevars:  PKG_CPPFLAGS += -I<repo>/include   PKG_LIBS += -L<repo>/rcppsmc_b200 -lsmcb200 -Wl,-rpath,<repo>/rcppsmc_b200
// The signatures, return names and shapes are the reference's own, so R/*.R and src/RcppExports.cpp stay untouched.
// R, Rcpp and Armadillo are not installed in this repository's build image: the file is syntax-checked there against the
// stand-in headers of oracle/shim (tests/test_cpp_host.py), not compiled against real Rcpp.
#include <RcppArmadillo.h>

#include "smcb200.hpp"

namespace {
// R's RNG stream cannot drive a device generator: derive the Philox key from it once per call, so set.seed() still makes a
// call reproducible.
inline uint64_t seed_from_R() { return (uint64_t)(unif_rand() * 9007199254740992.0); }
inline arma::vec to_arma(const smcb200::vec& v) { arma::vec a(v.size()); for (size_t i = 0; i < v.size(); ++i) a(i) = v[i]; return a; }
inline arma::mat to_arma(const smcb200::vec& v, size_t rows, size_t cols) {
    arma::mat m(rows, cols);
    for (size_t i = 0; i < rows * cols; ++i) m.memptr()[i] = v[i];
    return m;
}
inline smcb200::vec from_arma(const arma::vec& a) { return smcb200::vec(a.memptr(), a.memptr() + a.n_elem); }
inline smcb200::vec from_arma(const arma::mat& a) { return smcb200::vec(a.memptr(), a.memptr() + a.n_elem); }
}  // namespace

// [[Rcpp::export]]
Rcpp::List pfNonlinBS_impl(arma::vec data, long part) {
    try {
        smcb200::pfNonlinBS_result r = smcb200::pfNonlinBS_impl(from_arma(data), part, seed_from_R());
        return Rcpp::List::create(Rcpp::_["mean"] = to_arma(r.mean), Rcpp::_["sd"] = to_arma(r.sd));
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::DataFrame pfLineartBS_impl(arma::mat data, unsigned long part, bool usef, Rcpp::Function fun) {
    struct Hook {
        static void call(const double* Xm, const double* Ym, long T, void* user) {
            const Rcpp::Function& f = *static_cast<const Rcpp::Function*>(user);
            f(Rcpp::NumericVector(Xm, Xm + T), Rcpp::NumericVector(Ym, Ym + T));   // fun(Xm, Ym) after every iterate (src/pflineart.cpp:85)
        }
    };
    try {
        smcb200::pfLineartBS_result r = smcb200::pfLineartBS_impl(from_arma(data), part, usef ? &Hook::call : nullptr, usef ? &fun : nullptr, seed_from_R());
        return Rcpp::DataFrame::create(Rcpp::Named("Xm") = to_arma(r.Xm), Rcpp::Named("Xv") = to_arma(r.Xv), Rcpp::Named("Ym") = to_arma(r.Ym),
                                       Rcpp::Named("Yv") = to_arma(r.Yv));
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::List LinReg_impl(arma::mat Data, unsigned long lNumber) {
    try {
        smcb200::LinReg_result r = smcb200::LinReg_impl(from_arma(Data), lNumber, seed_from_R());
        return Rcpp::List::create(Rcpp::Named("theta") = to_arma(r.theta, lNumber, 3), Rcpp::Named("weights") = to_arma(r.weights),
                                  Rcpp::Named("logNC") = r.logNC);
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

namespace {
Rcpp::List la_list(const smcb200::LinRegLA_result& r, unsigned long N, bool adaptive) {
    arma::cube theta(N, 3, r.L);
    for (long l = 0; l < r.L; ++l)
        for (unsigned long c = 0; c < 3; ++c)
            for (unsigned long i = 0; i < N; ++i) theta.slice(l)(i, c) = r.theta[((size_t)l * 3 + c) * N + i];
    Rcpp::List out = Rcpp::List::create(Rcpp::Named("theta") = theta, Rcpp::Named("loglike") = to_arma(r.loglike, N, r.L),
                                        Rcpp::Named("logprior") = to_arma(r.logprior, N, r.L), Rcpp::Named("Weights") = to_arma(r.Weights, N, r.L),
                                        Rcpp::Named("ESS") = to_arma(r.ESS), Rcpp::Named("logNC_standard") = r.logNC_standard,
                                        Rcpp::Named("logNC_ps_rect") = r.logNC_ps_rect, Rcpp::Named("logNC_ps_trap") = r.logNC_ps_trap,
                                        Rcpp::Named("logNC_ps_trap2") = r.logNC_ps_trap2);
    if (adaptive) { out["Temps"] = to_arma(r.Temps, r.L, 1); out["mcmcRepeats"] = to_arma(r.mcmcRepeats, r.L, 1); }
    return out;
}
}  // namespace

// [[Rcpp::export]]
Rcpp::List LinRegLA_impl(arma::mat Data, arma::vec intemps, unsigned long lNumber) {
    try {
        return la_list(smcb200::LinRegLA_impl(from_arma(Data), from_arma(intemps), lNumber, seed_from_R()), lNumber, false);
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::List LinRegLA_adapt_impl(arma::mat Data, unsigned long lNumber, double resampTol, double tempTol) {
    try {
        return la_list(smcb200::LinRegLA_adapt_impl(from_arma(Data), lNumber, resampTol, tempTol, seed_from_R()), lNumber, true);
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::List blockpfGaussianOpt_impl(arma::vec data, long part, long lag) {
    try {
        smcb200::blockpfGaussianOpt_result r = smcb200::blockpfGaussianOpt_impl(from_arma(data), part, lag, seed_from_R());
        return Rcpp::List::create(Rcpp::_["weight"] = to_arma(r.weight), Rcpp::_["values"] = to_arma(r.values, part, data.n_elem),
                                  Rcpp::_["logNC"] = r.logNC);
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::DataFrame nonLinPMMH_impl(arma::vec data, unsigned long lNumber, unsigned long lMCMCits, bool verbose = false, int msg_freq = 100) {
    struct Hook {
        static void call(long i, long m, double mv, double mw, double ll, double lp, double ar, void*) {   // src/nonLinPMMH.cpp:116-138
            Rcpp::Rcout << "Percentage completed: " << (100.0 * (double)i / (double)m) << "%; mean sigv " << mv << " mean sigw " << mw
                        << " loglike " << ll << " logprior " << lp << " MH acceptance " << ar << std::endl;
        }
    };
    try {
        smcb200::nonLinPMMH_result r = smcb200::nonLinPMMH_impl(from_arma(data), lNumber, lMCMCits, verbose ? &Hook::call : nullptr, msg_freq, nullptr,
                                                                seed_from_R());
        return Rcpp::DataFrame::create(Rcpp::Named("samples_sigv") = to_arma(r.samples_sigv), Rcpp::Named("samples_sigw") = to_arma(r.samples_sigw),
                                       Rcpp::Named("loglike") = to_arma(r.loglike), Rcpp::Named("logprior") = to_arma(r.logprior),
                                       Rcpp::Named("MH_acceptance_rate") = to_arma(r.MH_acceptance_rate));
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}

// [[Rcpp::export]]
Rcpp::List compareNCestimates_imp(arma::vec data, long lParticleNum, int simNum, Rcpp::List parInits, arma::mat referenceTraj) {
    try {
        smcb200::parInits p{parInits["phi"], parInits["varStateEvol"], parInits["stateInit"], parInits["varObs"]};
        smcb200::compareNCestimates_result r = smcb200::compareNCestimates_imp(from_arma(data), lParticleNum, simNum, p, from_arma(referenceTraj),
                                                                               seed_from_R());
        Rcpp::List out;
        out["smcOut"] = to_arma(r.smcOut, simNum, 4);
        out["csmcOut"] = to_arma(r.csmcOut, simNum, 4);
        return out;
    } catch (const smcb200::error& e) { Rcpp::stop(e.what()); }
    return R_NilValue;
}
