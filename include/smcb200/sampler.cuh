// Public plug-in header of the B200 engine: the reference's C++ template surface under its own names, for user models compiled
// with nvcc (-gencode arch=compute_100a,code=sm_100a).  It replaces, for a model written as __host__ __device__ functors,
//   smc::sampler<Space,Params>          inst/include/sampler.h:133-237
//   smc::moveset<Space,Params>          inst/include/moveset.h:73-96        (the four per-particle functions = static members of Model)
//   smc::conditionalSampler<Space,Params>  inst/include/conditionalSampler.h:39-155
//   smc::adaptMethods<Space,Params>     inst/include/adaptMethods.h:45-51
//   smc::staticModelAdapt               inst/include/staticModelAdapt.h:37-171
//   ResampleType / HistoryType / PathSamplingType                          inst/include/sampler.h:45-68
//
//   #include <smcb200/sampler.cuh>
//   struct MyModel {
//       static constexpr int D = 3;                                         // doubles per particle
//       struct Params { double temp_prev, temp_cur; double chol[4]; };      // algParams
//       SMCB_HD static void init(double (&x)[D], double& logw, const Params&, smc::rng&);
//       SMCB_HD static void move(long t, double (&x)[D], double& logw, const Params&, smc::rng&);
//       SMCB_HD static bool mcmc(long t, double (&x)[D], double& logw, const Params&, smc::rng&);   // optional
//   };
//   smc::sampler<MyModel> s(N, HistoryType::RAM);
//   s.SetResampleParams(ResampleType::SYSTEMATIC, 0.5); s.SetAdaptMethods(&myAdapt); s.Initialise(); s.IterateUntil(T);
//
// Differences a port has to know about: Space is a fixed-size array of doubles (stored as structure-of-arrays columns on the
// device); the per-particle functions are static members instead of virtuals (there is no virtual dispatch on the device);
// randomness comes from the smc::rng handed to each call (counter-based Philox4x32-10; the reference calls R's global RNG);
// integrands are functor objects instead of function pointers + void* (the functor's members are the auxiliary data).
#pragma once
#include "../../rcppsmc_b200/csrc/generic.cuh"
#include "../../rcppsmc_b200/csrc/csmc.cuh"

namespace ResampleType { enum Enum { MULTINOMIAL = 0, RESIDUAL, STRATIFIED, SYSTEMATIC }; }   // sampler.h:45-51
namespace HistoryType { enum Enum { NONE = 0, RAM, AL }; }                                    // sampler.h:61-66
namespace PathSamplingType { enum Enum { RECTANGLE = 0, TRAPEZOID1, TRAPEZOID2 }; }           // sampler.h:53-58

namespace smc {

using rng = smcb::Rng;
template <class Model> using population = smcb::Population<Model>;
template <class Model> using adaptMethods = smcb::AdaptMethods<Model>;
using historyflags = smcb::HistoryFlags;

template <class Model>
class sampler : public smcb::GenericSampler<Model> {
    using Base = smcb::GenericSampler<Model>;
public:
    sampler(long lSize, HistoryType::Enum htHistoryMode, unsigned long long seed = 1, cudaStream_t stream = 0)
        : Base(lSize, (int)htHistoryMode, seed, stream) {}
    void SetResampleParams(ResampleType::Enum rtMode, double dThreshold) { Base::SetResampleParams((int)rtMode, dThreshold); }
    using Base::IntegratePathSampling;
    template <class F, class W> double IntegratePathSampling(PathSamplingType::Enum type, F f, W width) { return Base::IntegratePathSampling((int)type, f, width); }
    template <class F, class W> double IntegratePathSampling(F f, W width) { return Base::IntegratePathSampling((int)PathSamplingType::TRAPEZOID2, f, width); }   // sampler.h:201
};

// smc::conditionalSampler<Space,Params> (inst/include/conditionalSampler.h:39-155) for a model with the device functors
//   init(x, logw, params, z) / move(t, x, logw, params, z) / weight(t, x, logw, params)     (z = the call's standard normals)
// and constexpr D, NZ_INIT, NZ_MOVE: conditional resampling with all four ResampleType schemes (conditionalSampler.h:377-655),
// the reference particle pinned by DoConditionalMove (moveset.h:170-180), HistoryType::AL ancestry.  One object is one chain whose
// reference-slot indices carry over between runs; a whole run executes as ONE kernel launch of one thread block.
template <class Model>
class conditionalSampler : public smcb::ConditionalSampler<Model> {
    using Base = smcb::ConditionalSampler<Model>;
public:
    using Space = typename Base::Space;
    conditionalSampler(long lSize, HistoryType::Enum htHistoryMode, const std::vector<Space>& referenceTrajectoryInit, unsigned long long seed = 1,
                       cudaStream_t stream = 0)
        : Base(lSize, (int)htHistoryMode, referenceTrajectoryInit, seed, stream) {}
    void SetResampleParams(ResampleType::Enum rtMode, double dThreshold) { Base::SetResampleParams((int)rtMode, dThreshold); }
};

// Adaptation helpers for likelihood-tempered static models (staticModelAdapt.h:37-171), evaluated on the device through the
// population view an adaptation hook receives.  LL = state component that holds the particle's log-likelihood; the first
// DT components are the parameter vector theta whose covariance drives the MCMC proposal.
template <class Model, int LL, int DT>
class staticModelAdapt {
    static constexpr int D = Model::D;
    static_assert(LL >= 0 && LL < D && DT >= 1 && DT <= D, "component indices out of range");
    static_assert(DT * (DT + 1) / 2 <= 16, "covariance of at most 5 components per pass");
public:   // (functor types that reach a kernel template must be public: nvcc restriction)
    struct CessTerm {   // logweight + k * tempDiff * loglike (staticModelAdapt.h:84-85)
        double scale;
        __device__ double operator()(const double (&x)[D], double lw) const { return lw + scale * x[LL]; }
    };
    struct ColSum {
        __device__ void operator()(const double (&x)[D], double, double (&o)[DT]) const {
#pragma unroll
            for (int k = 0; k < DT; ++k) o[k] = x[k];
        }
    };
    struct Scatter {
        double mean[DT];
        __device__ void operator()(const double (&x)[D], double lw, double (&o)[DT * (DT + 1) / 2]) const {
            const double w = exp(lw);
            int q = 0;
#pragma unroll
            for (int r = 0; r < DT; ++r)
#pragma unroll
                for (int c = r; c < DT; ++c) o[q++] = ((x[r] - mean[r]) * w) * (x[c] - mean[c]);
        }
    };
private:
    std::vector<double> temp;
    double empCov[DT][DT] = {};
    double cholCov[DT][DT] = {};   // upper factor U, U^T U = empCov (arma::chol)

    double CESSdiff(const population<Model>& pop, double tempDiff, double desiredCESS) const {   // :82-88
        const double ls1 = pop.LogSumExp(CessTerm{tempDiff}), ls2 = pop.LogSumExp(CessTerm{2 * tempDiff});
        return std::exp(std::log((double)pop.GetNumber()) + 2 * ls1 - ls2) - desiredCESS;
    }
    double bisection(double curr, const population<Model>& pop, double desiredCESS, double epsilon) const {   // :91-119
        double a = curr, b = 1.0;
        double f_a = CESSdiff(pop, a - curr, desiredCESS), f_b = CESSdiff(pop, b - curr, desiredCESS);
        if (f_a * f_b > 0) throw std::runtime_error("Bisection method to choose the next temperature failed");
        double m = (a + b) / 2.0, f_m = CESSdiff(pop, m - curr, desiredCESS), err = 10.0;
        while (err > epsilon) {
            if (f_m < 0.0) { b = m; f_b = f_m; } else { a = m; f_a = f_m; }
            m = (a + b) / 2.0;
            f_m = CESSdiff(pop, m - curr, desiredCESS);
            err = std::abs(f_m);
        }
        return m;
    }

public:
    staticModelAdapt() { temp.push_back(0.0); }
    void ChooseTemp(const population<Model>& pop, double desiredCESS, double epsilon = 0.01) {   // :127-134
        const double temp_curr = temp.back();
        if (CESSdiff(pop, 1.0 - temp_curr, desiredCESS) >= -epsilon) temp.push_back(1.0);
        else temp.push_back(bisection(temp_curr, pop, desiredCESS, epsilon));
    }
    void calcEmpCov(const population<Model>& pop) {   // :137-143: weighted scatter about the UNWEIGHTED column means
        double cs[DT];
        pop.template Sum<DT>(ColSum{}, cs);
        Scatter sc;
        for (int k = 0; k < DT; ++k) sc.mean[k] = cs[k] / (double)pop.GetNumber();
        double tri[DT * (DT + 1) / 2];
        pop.template Sum<DT * (DT + 1) / 2>(sc, tri);
        int q = 0;
        for (int r = 0; r < DT; ++r) for (int c = r; c < DT; ++c) { empCov[r][c] = tri[q]; empCov[c][r] = tri[q]; ++q; }
    }
    void calcCholCov(const population<Model>& pop) {   // :150-153
        calcEmpCov(pop);
        for (int r = 0; r < DT; ++r) for (int c = 0; c < DT; ++c) cholCov[r][c] = 0.0;
        for (int j = 0; j < DT; ++j) {
            double d = empCov[j][j];
            for (int k = 0; k < j; ++k) d -= cholCov[k][j] * cholCov[k][j];
            if (!(d > 0.0)) throw std::runtime_error("chol(): decomposition failed");
            cholCov[j][j] = std::sqrt(d);
            for (int c = j + 1; c < DT; ++c) {
                double v = empCov[j][c];
                for (int k = 0; k < j; ++k) v -= cholCov[k][j] * cholCov[k][c];
                cholCov[j][c] = v / cholCov[j][j];
            }
        }
    }
    int calcMcmcRepeats(double acceptProb, double desiredAcceptProb, int initialN, int maxReps) const {   // :161-171
        if (acceptProb + 1.0 <= 1e-9) return initialN;
        if (acceptProb - 1.0 >= -1e-9) return 1;
        if (acceptProb <= 1e-9) return maxReps;
        return std::min(maxReps, (int)std::ceil(std::log(1.0 - desiredAcceptProb) / std::log(1.0 - acceptProb)));
    }
    double GetTemp(int t) const { return temp[t]; }
    double GetTempCurr() const { return temp.back(); }
    double GetTempPrevious() const { return temp.rbegin()[1]; }
    std::vector<double> GetTemps() const { return temp; }
    double GetCholCov(int r, int c) const { return cholCov[r][c]; }
    double GetEmpCov(int r, int c) const { return empCov[r][c]; }
};

}  // namespace smc
