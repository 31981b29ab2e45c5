// smcb200.hpp -- C++ host side above the C ABI (include/smcb200.h): the reference's eight exported functions with their own
// names, argument order and result fields (reference src/RcppExports.cpp:15-139 and the *_impl bodies cited per function),
// on std::vector instead of Rcpp / Armadillo types.  Header-only; link against libsmcb200.so.  This is what the bodies of
// src/*.cpp reduce to once the numeric work is behind the C ABI: an Rcpp build wraps these structs into List / DataFrame
// (INTEGRATION.md).  Errors surface as smcb200::error carrying smcb200_last_error() -- the counterpart of smc::exception /
// Rcpp::stop in the reference drivers.  Matrices are column-major, as arma::mat.
#ifndef SMCB200_HPP
#define SMCB200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "smcb200.h"

namespace smcb200 {

struct error : std::runtime_error {
    int status;
    error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(int status) {
    if (status != SMCB200_OK) throw error(status, smcb200_last_error());
}

typedef std::vector<double> vec;

// pfNonlinBS_impl(arma::vec data, long part) -> List(mean, sd)                                   src/pfnonlinbs.cpp:52-79
struct pfNonlinBS_result { vec mean, sd, logNC, ESS; std::vector<int> resampled; };
inline pfNonlinBS_result pfNonlinBS_impl(const vec& data, long part, uint64_t seed = 1, int resample_mode = SMCB200_MULTINOMIAL,
                                         double threshold = -1.0) {
    const long T = (long)data.size();
    pfNonlinBS_result r{vec(T), vec(T), vec(T), vec(T), std::vector<int>(T)};
    check(smcb200_pfNonlinBS(data.data(), T, part, resample_mode, threshold < 0 ? 1.01 * (double)part : threshold, seed, nullptr, nullptr,
                             r.mean.data(), r.sd.data(), r.logNC.data(), r.ESS.data(), r.resampled.data()));   // reference: MULTINOMIAL, 1.01 N (:62)
    return r;
}

// pfLineartBS_impl(arma::mat data [T x 2], unsigned long part, bool usef, Rcpp::Function fun) -> DataFrame(Xm, Xv, Ym, Yv)   src/pflineart.cpp:53-99
struct pfLineartBS_result { vec Xm, Xv, Ym, Yv; };
inline pfLineartBS_result pfLineartBS_impl(const vec& data_Tx2, unsigned long part, smcb200_step_callback fun = nullptr, void* user = nullptr,
                                           uint64_t seed = 1) {
    const long T = (long)data_Tx2.size() / 2;
    pfLineartBS_result r{vec(T), vec(T), vec(T), vec(T)};
    check(smcb200_pfLineartBS(data_Tx2.data(), T, (long)part, SMCB200_RESIDUAL, 0.5, seed, nullptr, nullptr, fun, user, r.Xm.data(), r.Xv.data(),
                              r.Ym.data(), r.Yv.data(), nullptr, nullptr, nullptr));                            // reference: RESIDUAL, 0.5 (:67)
    return r;
}

// LinReg_impl(arma::mat Data [n x 2], unsigned long lNumber) -> List(theta [N x 3], weights [N], logNC)       src/LinReg.cpp:46-84
struct LinReg_result { vec theta, weights; double logNC; };
inline LinReg_result LinReg_impl(const vec& Data_nx2, unsigned long lNumber, uint64_t seed = 1) {
    LinReg_result r{vec(3 * lNumber), vec(lNumber), 0.0};
    check(smcb200_LinReg(Data_nx2.data(), (long)Data_nx2.size() / 2, (long)lNumber, seed, nullptr, 0, nullptr, 0, nullptr, r.theta.data(),
                         r.weights.data(), &r.logNC));
    return r;
}

// LinRegLA_impl / LinRegLA_adapt_impl -> List(theta [N x 3 x L], loglike, logprior, Weights [N x L], ESS [L], logNC_standard,
// logNC_ps_rect, logNC_ps_trap, logNC_ps_trap2 [, Temps [L], mcmcRepeats [L]])            src/LinReg_LA.cpp:45-115, src/LinReg_LA_adapt.cpp:36-107
struct LinRegLA_result {
    long L;
    vec theta, loglike, logprior, Weights, ESS, Temps, mcmcRepeats;
    double logNC_standard, logNC_ps_rect, logNC_ps_trap, logNC_ps_trap2;
};
inline LinRegLA_result LinRegLA_impl(const vec& Data_nx2, const vec& intemps, unsigned long lNumber, uint64_t seed = 1) {
    const long N = (long)lNumber, L = (long)intemps.size();
    LinRegLA_result r{L, vec(3 * N * L), vec(N * L), vec(N * L), vec(N * L), vec(L), intemps, vec(), 0, 0, 0, 0};
    double nc[4];
    check(smcb200_LinRegLA(Data_nx2.data(), (long)Data_nx2.size() / 2, intemps.data(), L, N, seed, nullptr, 0, nullptr, 0, r.theta.data(),
                           r.loglike.data(), r.logprior.data(), r.Weights.data(), r.ESS.data(), nc));
    r.logNC_standard = nc[0]; r.logNC_ps_rect = nc[1]; r.logNC_ps_trap = nc[2]; r.logNC_ps_trap2 = nc[3];
    return r;
}
inline LinRegLA_result LinRegLA_adapt_impl(const vec& Data_nx2, unsigned long lNumber, double resampTol, double tempTol, uint64_t seed = 1,
                                           long max_temps = 512) {
    // The history is sized by the temperature capacity (6 doubles per particle and temperature, on the device and here), while
    // a run typically needs 20-40 temperatures: a small capacity first, the caller's only if the schedule turns out longer (the
    // draws are counter-based, so the repeat reproduces the same run).
    const long N = (long)lNumber;
    LinRegLA_result r{};
    double nc[4];
    long L = 0;
    for (long C : {max_temps > 64 ? 64L : max_temps, max_temps}) {
        r = LinRegLA_result{0, vec(3 * N * C), vec(N * C), vec(N * C), vec(N * C), vec(C), vec(C), vec(C), 0, 0, 0, 0};
        L = smcb200_LinRegLA_adapt(Data_nx2.data(), (long)Data_nx2.size() / 2, N, resampTol, tempTol, seed, nullptr, 0, nullptr, 0, C,
                                   r.theta.data(), r.loglike.data(), r.logprior.data(), r.Weights.data(), r.ESS.data(), r.Temps.data(),
                                   r.mcmcRepeats.data(), nc, nullptr);
        if (L > 0 || C == max_temps || std::string(smcb200_last_error()).find("more temperatures than max_temps") == std::string::npos) break;
    }
    if (L <= 0) throw error((int)-L, smcb200_last_error());
    r.L = L;
    r.theta.resize(3 * N * L); r.loglike.resize(N * L); r.logprior.resize(N * L); r.Weights.resize(N * L);
    r.ESS.resize(L); r.Temps.resize(L); r.mcmcRepeats.resize(L);
    r.logNC_standard = nc[0]; r.logNC_ps_rect = nc[1]; r.logNC_ps_trap = nc[2]; r.logNC_ps_trap2 = nc[3];
    return r;
}

// blockpfGaussianOpt_impl(arma::vec data, long part, long lag) -> List(weight [N], values [N x T], logNC)     src/blockpfgaussianopt.cpp:41-71
struct blockpfGaussianOpt_result { vec weight, values; double logNC; };
inline blockpfGaussianOpt_result blockpfGaussianOpt_impl(const vec& data, long part, long lag, uint64_t seed = 1) {
    blockpfGaussianOpt_result r{vec(part), vec((size_t)part * data.size()), 0.0};
    check(smcb200_blockpfGaussianOpt(data.data(), (long)data.size(), part, lag, seed, nullptr, nullptr, r.weight.data(), r.values.data(), &r.logNC,
                                     nullptr, nullptr, nullptr));
    return r;
}

// nonLinPMMH_impl(arma::vec data, unsigned long lNumber, unsigned long lMCMCits, bool verbose, int msg_freq)
//   -> DataFrame(samples_sigv, samples_sigw, loglike, logprior, MH_acceptance_rate)                           src/nonLinPMMH.cpp:37-153
struct nonLinPMMH_result { vec samples_sigv, samples_sigw, loglike, logprior, MH_acceptance_rate; };
inline nonLinPMMH_result nonLinPMMH_impl(const vec& data, unsigned long lNumber, unsigned long lMCMCits, smcb200_pmmh_progress progress = nullptr,
                                         int msg_freq = 100, void* user = nullptr, uint64_t seed = 1) {
    const size_t M = lMCMCits + 1;
    nonLinPMMH_result r{vec(M), vec(M), vec(M), vec(M), vec(M)};
    check(smcb200_nonLinPMMH(data.data(), (long)data.size(), (long)lNumber, (long)lMCMCits, seed, SMCB200_MULTINOMIAL, 0.5, nullptr, nullptr, nullptr,
                             nullptr, nullptr, progress, msg_freq, user, r.samples_sigv.data(), r.samples_sigw.data(), r.loglike.data(),
                             r.logprior.data(), r.MH_acceptance_rate.data()));                                   // reference: MULTINOMIAL, 0.5 (:77)
    return r;
}

// compareNCestimates_imp(arma::vec data, long lParticleNum, int simNum, Rcpp::List parInits, arma::mat referenceTraj [T x simNum])
//   -> List(smcOut [simNum x 4], csmcOut [simNum x 4])                                                          src/cSMCexamples.cpp:36-139
struct parInits { double phi, varStateEvol, stateInit, varObs; };
struct compareNCestimates_result { vec smcOut, csmcOut; };
inline compareNCestimates_result compareNCestimates_imp(const vec& data, long lParticleNum, int simNum, const parInits& p,
                                                        const vec& referenceTraj_TxS, uint64_t seed = 1, bool independent_sims = false) {
    compareNCestimates_result r{vec(4 * (size_t)simNum), vec(4 * (size_t)simNum)};
    check(smcb200_compareNCestimates(data.data(), (long)data.size(), lParticleNum, simNum, p.phi, p.varStateEvol, p.stateInit, p.varObs,
                                     referenceTraj_TxS.data(), seed, independent_sims ? 1 : 0, r.smcOut.data(), r.csmcOut.data()));
    return r;
}

}  // namespace smcb200
#endif
