/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 * Plain-C restatement of the RcppSMC 0.2.9 SMC hot path (see smc_oracle.c for the file:line map).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (rcppsmc_b200/) never does. */
#ifndef SMC_ORACLE_H
#define SMC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { SMCO_MULTINOMIAL = 0, SMCO_RESIDUAL = 1, SMCO_STRATIFIED = 2, SMCO_SYSTEMATIC = 3 };

double smco_lse(const double* logw, long n);
double smco_ess(const double* logw, long n);
void smco_normalised_weights(const double* logw, long n, double* w);
void smco_quantise_weights(const double* w, long n, double* wq);
void smco_cumsum(const double* w, long n, double* W);

long smco_counts_systematic(const double* w, long n, double U, int* count);
long smco_counts_stratified(const double* w, long n, const double* U, int* count);
long smco_counts_multinomial_iid(const double* w, long n, long ndraws, const double* U, int* count);
long smco_residual_floor(const double* w, long n, int* floor_count, double* resid_w);
void smco_counts_to_indices(int* count, long n, unsigned* idx);
int smco_resample_from_weights(int mode, const double* w, long n, const double* U, long nU,
                               const int* counts_in, unsigned* idx, int* count_out);
int smco_resample(int mode, const double* logw, long n, const double* U, long nU,
                  const int* counts_in, unsigned* idx, int* count_out);
double smco_integrate(const double* logw, const double* f, long n);

void smco_set_uniforms_by_step(int on);
void smco_set_grid_weights(int on);
void smco_sim_nonlin(long T, const double* z_state, const double* z_obs, double var_init,
                     double var_evol, double var_obs, double cos_offset, double* x, double* y);
int smco_pfnonlinbs(const double* y, long T, long N, int mode, double threshold, const double* z,
                    const double* U, unsigned long long seed, double* mean, double* sd,
                    double* lognc, double* ess, int* resampled, unsigned* idx_trace);
int smco_pflineart(const double* data, long T, long N, int mode, double threshold, const double* z, const double* U,
                   unsigned long long seed, double* Xm, double* Xv, double* Ym, double* Yv, double* lognc, double* ess,
                   int* resampled);
long smco_linreg_la(const double* data, long n_obs, long N, int adaptive, const double* temps_in, long L_in, double resampTol,
                    double tempTol, const double* z, const double* U, unsigned long long seed, long max_temps, double* theta_out,
                    double* ll_out, double* lp_out, double* w_out, double* ess_out, double* temps_out, double* reps_out,
                    double* lognc_out, double* accept_out);
double smco_pmmh_filter(const double* y, long T, long N, double sigv, double sigw, int mode, double threshold, const double* z,
                        const double* U, unsigned long long seed);
int smco_nonlin_pmmh(const double* y, long T, long N, long M, int mode, double threshold, const double* innov_v, const double* innov_w,
                     const double* unif, const double* filter_noise, const double* filter_unif, unsigned long long seed, double* sigv,
                     double* sigw, double* loglike, double* logprior, double* accept_rate);
int smco_linreg(const double* data, long n_obs, long N, const double* z, const double* U, const double* Udraw,
                unsigned long long seed, double* theta_out, double* w_out, double* lognc_out);
int smco_blockpf(const double* y, long T, long N, long lag_max, const double* z, const double* U, unsigned long long seed,
                 double* weight, double* values, double* lognc, double* ess_trace, int* resampled, unsigned* anc);
int smco_csmc_lgss_run(const double* y, long T, long N, int mode, double phi, double var_evol, double x0, double var_obs,
                       const double* ref, unsigned* refidx, const double* z, const double* U, int addressed,
                       unsigned long long seed, double* lognc, unsigned* anc, double* values, double* logw, int* resampled,
                       long* kt_out, long* nstoch);
double smco_bpf_lgss(const double* y, long T, long N, int mode, double phi, double var_evol, double x0, double var_obs,
                     const double* z, const double* U, unsigned long long seed, int* resampled);
void smco_ancestor_lines(const unsigned* anc, const double* values, long T, long N, unsigned* lines, double* line_values);
double smco_time_pfnonlinbs(const double* y, long T, long N, int mode, double threshold,
                            unsigned long long seed);

#ifdef __cplusplus
}
#endif
#endif
