/* smcb200 -- C ABI of the B200-native SMC engine (libsmcb200.so).
 *
 * This is the drop-in boundary for the SMC hot path of RcppSMC 0.2.9: the Rcpp glue
 * (reference src/RcppExports.cpp:15-139) keeps its eight .Call routines and forwards the numeric work to
 * these entry points instead of instantiating smc::sampler<Space,Params> on the CPU (INTEGRATION.md shows
 * the stub).  Plain pointers and sizes only; all buffers named `host` are caller-owned host memory.
 * No C++ exception crosses this boundary: every function returns a status and smcb200_last_error() gives
 * the message (the analogue of the reference's smc::exception / Rcpp::stop, smc-exception.h:52-75).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with SMCB200_ERR_CUDA.
 * One handle = one caller thread at a time (the reference is not re-entrant at all: file-scope model
 * globals, inst/include/pfnonlinbs.h:50).
 */
#ifndef SMCB200_H
#define SMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB200_OK 0
#define SMCB200_ERR_INVALID 1 /* bad argument (message in smcb200_last_error) */
#define SMCB200_ERR_CUDA 2    /* CUDA runtime failure or no device */
#define SMCB200_ERR_STATE 3   /* call order violation (e.g. read before run) */

/* ResampleType::Enum, reference inst/include/sampler.h:45-51 (same numeric values). */
#define SMCB200_MULTINOMIAL 0
#define SMCB200_RESIDUAL 1
#define SMCB200_STRATIFIED 2
#define SMCB200_SYSTEMATIC 3

const char* smcb200_last_error(void);
int smcb200_version(void);
/* Number of visible CUDA devices (0 when none / no driver); never fails. */
int smcb200_device_count(void);

/* ---- unit level ------------------------------------------------------------------------------------- */

/* Philox4x32-10 block function (known-answer surface for the RNG that replaces R's; host computation). */
int smcb200_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The same function evaluated by a kernel (the device build of the round function): n blocks of (ctr[4], key[2]) in,
 * n blocks of 4 words out.  Known-answer tests of the device path. */
int smcb200_philox4x32_device(const uint32_t* ctr_key_host, long n, uint32_t* out_host);

/* n standard normals exactly as the engine draws them for (seed, stream, step): particle i gets out[i].
 * Runs on the device. */
int smcb200_philox_normals(uint64_t seed, uint32_t stream, uint32_t step, long n, double* out_host);

/* The engine's own device implementations of the elementary functions it uses in place of libm's
 * (fn: 0 exp(x) for x <= 0, 1 -2 ln(u) for u in (0,1], 2 sin(2 pi u), 3 cos(2 pi u), 4 sqrt(x)); accuracy test surface. */
int smcb200_fastmath_eval(int fn, const double* in_host, long n, double* out_host);

/* Replaces smc::stableLogSumWeights (population.h:46-50) and sampler::GetESS (sampler.h:413-417).
 * out3 = { log sum exp(logw), ESS = (sum w)^2 / sum w^2, max(logw) }. */
int smcb200_lse(const double* logw_host, long n, double* out3_host);

/* Replaces sampler::Resample's count + index stages (sampler.h:786-857) on NORMALISED weights.
 *   weights      n normalised weights (quantised internally to the 2^-52 grid; weights already on the grid
 *                give ancestor indices bit-identical to the reference for SYSTEMATIC / STRATIFIED)
 *   uniforms     U in [0,1): SYSTEMATIC 1 value; STRATIFIED n+1 values (reference draw order, :810-819);
 *                MULTINOMIAL n values, RESIDUAL n values (native iid inverse-CDF definition)
 *   counts_in    optional: children counts to map directly (replaces rmultinom's output, :790/:801);
 *                when given, weights / uniforms may be NULL
 *   idx_out      n ancestor indices = uRSIndices (:845-857)
 *   counts_out   optional: n children counts */
int smcb200_resample(int mode, const double* weights_host, long n, const double* uniforms_host, long n_uniforms,
                     const int* counts_in_host, unsigned int* idx_out_host, int* counts_out_host);

/* ---- model level: pfNonlinBS (reference src/pfnonlinbs.cpp:52-79, R/pfNonlinBS.R) ------------------- */

/* One-shot filter, the call the Rcpp wrapper makes.  `data` = y[0..T), `particles` = N.
 * Reference behaviour is resample_mode = SMCB200_MULTINOMIAL, threshold = 1.01*N (src/pfnonlinbs.cpp:62).
 *   seed      Philox key (the reference uses R's global RNG instead)
 *   noise     optional N*T injected standard normals in the reference's consumption order
 *             (N for Initialise, then N per Iterate), else NULL -> Philox
 *   uniforms  optional injected resampling uniforms, indexed by step: SYSTEMATIC uniforms[t];
 *             STRATIFIED uniforms[t*(N+1) + j]; MULTINOMIAL / RESIDUAL uniforms[t*N + j]; else NULL -> Philox
 * Outputs (length T each; any may be NULL): mean, sd as returned by pfNonlinBS_impl (:77-78); lognc =
 * GetLogNCPath() after each step; ess = the ESS each step's resampling decision saw; resampled flags. */
int smcb200_pfNonlinBS(const double* data_host, long T, long particles, int resample_mode, double threshold,
                       uint64_t seed, const double* noise_host, const double* uniforms_host, double* mean_host,
                       double* sd_host, double* lognc_host, double* ess_host, int* resampled_host);

/* Handle-based variant: buffers live across runs so a benchmark can separate allocation / transfers from the
 * filter itself.  `stream` is a cudaStream_t (NULL = default stream); all work of the handle is enqueued there. */
typedef struct smcb200_pf smcb200_pf;
int smcb200_pfNonlinBS_create(long particles, long tcap, uint64_t seed, void* stream, smcb200_pf** out);
/* Island mode (one process per GPU on one NVSwitch box): rank `rank` of `world` owns particles_local particles of ONE
 * population of world*particles_local particles with GLOBAL resampling.  Each rank exports one 64-byte CUDA IPC handle
 * (ipc_export), the host all-gathers them (any transport) and hands all `world` handles to ipc_import; after that the
 * kernels exchange per-step scalars and remote parents directly over NVLink peer memory.  Every rank must use the same
 * seed, data and run() arguments; read() returns identical traces on every rank.  SYSTEMATIC / STRATIFIED only. */
int smcb200_pfNonlinBS_create_island(long particles_local, long tcap, uint64_t seed, void* stream, int rank, int world, smcb200_pf** out);
int smcb200_pfNonlinBS_ipc_export(smcb200_pf* h, void* handle64_out);
int smcb200_pfNonlinBS_ipc_import(smcb200_pf* h, const void* handles_world_x64);
int smcb200_pfNonlinBS_set_data(smcb200_pf* h, const double* data_host, long T);
int smcb200_pfNonlinBS_set_noise(smcb200_pf* h, const double* noise_host, const double* uniforms_host, long n_uniforms);
/* Enqueues Initialise + (T-1) Iterate + the final moment pass; asynchronous. */
int smcb200_pfNonlinBS_run(smcb200_pf* h, int resample_mode, double threshold);
int smcb200_pfNonlinBS_sync(smcb200_pf* h);
/* Device time of the last run in milliseconds (CUDA events on the handle's stream). */
int smcb200_pfNonlinBS_elapsed_ms(smcb200_pf* h, float* ms);
int smcb200_pfNonlinBS_read(smcb200_pf* h, double* mean_host, double* sd_host, double* lognc_host, double* ess_host,
                            int* resampled_host);
/* Ancestor indices (uRSIndices) of the last step's resampling; identity if it did not resample. */
int smcb200_pfNonlinBS_read_ancestors(smcb200_pf* h, unsigned int* idx_host);
/* Kernel launches issued by the last run (for benchmark bookkeeping). */
long smcb200_pfNonlinBS_launches(smcb200_pf* h);
/* Per-kernel device timing: with profiling on, run() records one CUDA event after every launch; kernel_ms()
 * then returns the summed duration and launch count of the propagate (k_step) and resampling (k_resample_scan)
 * kernels of the last run (Initialise and the closing moment pass excluded). */
int smcb200_pfNonlinBS_set_profile(smcb200_pf* h, int on);
int smcb200_pfNonlinBS_kernel_ms(smcb200_pf* h, float* move_ms, long* move_launches, float* scan_ms, long* scan_launches);
int smcb200_pfNonlinBS_destroy(smcb200_pf* h);

/* ---- model level: pfLineartBS (reference src/pflineart.cpp:53-99, R/pfLineartBS.R) ------------------- */

/* Per-step callback: the reference calls the R closure fun(Xm, Ym) after every Iterate when usef is TRUE
 * (src/pflineart.cpp:85); entries beyond the current step are 0.  NULL = usef FALSE. */
typedef void (*smcb200_step_callback)(const double* Xm, const double* Ym, long T, void* user);

/* `data` is the reference's arma::mat [T x 2] in column-major order (x column, then y column).
 * Reference behaviour: resample_mode = SMCB200_RESIDUAL, threshold = 0.5 (src/pflineart.cpp:67).
 * noise: optional 4*N*T injected standard normals in the reference's consumption order (per particle x_pos, y_pos,
 * x_vel, y_vel at Initialise; x_pos, x_vel, y_pos, y_vel per move, :147-150, :163-166); uniforms as in
 * smcb200_pfNonlinBS.  Outputs: Xm, Xv, Ym, Yv as returned by pfLineartBS_impl (:90-93; Xv/Yv are variances) plus the
 * logNC path, ESS and resampled traces. */
int smcb200_pfLineartBS(const double* data_host, long T, long particles, int resample_mode, double threshold,
                        uint64_t seed, const double* noise_host, const double* uniforms_host,
                        smcb200_step_callback callback, void* user, double* Xm_host, double* Xv_host, double* Ym_host,
                        double* Yv_host, double* lognc_host, double* ess_host, int* resampled_host);

/* ---- model level: LinRegLA_adapt / LinRegLA (reference src/LinReg_LA_adapt.cpp:36-106, src/LinReg_LA.cpp:45-111) ---- */

/* Likelihood-annealing SMC sampler with adaptive temperatures (CESS bisection, rho = tempTol), adaptive MCMC repeats and
 * an empirical-covariance random-walk proposal.  `data` = arma::mat [n_obs x 2] column-major (y column, x column).
 * Outputs use the reference's shapes with L = number of temperatures (the return value, <= max_temps):
 *   theta [N x 3 x L] (cube, column-major), loglike / logprior / weights [N x L], ess / temps / repeats [L],
 *   lognc4 = { logNC_standard, logNC_ps_rect, logNC_ps_trap, logNC_ps_trap2 }, accept (optional) = acceptProb per step.
 * noise / uniforms: optional injected N(0,1) / U(0,1) streams in the reference's consumption order (Initialise: 2 normals
 * and 3 uniforms per particle [Gamma(3) as a sum of exponentials]; per step: 1 uniform when resampling, then per MCMC
 * repeat and particle 3 normals + 1 uniform).  Returns L > 0, or -status on failure (smcb200_last_error()). */
long smcb200_LinRegLA_adapt(const double* data_host, long n_obs, long particles, double resampTol, double tempTol, uint64_t seed,
                            const double* noise_host, long n_noise, const double* uniforms_host, long n_uniforms, long max_temps,
                            double* theta_host, double* loglike_host, double* logprior_host, double* weights_host, double* ess_host,
                            double* temps_host, double* repeats_host, double* lognc4_host, double* accept_host);

/* Fixed temperature schedule temps[n_temps], proposal chol(covRW), 10 MH repeats every step, SYSTEMATIC / 0.5. */
int smcb200_LinRegLA(const double* data_host, long n_obs, const double* temps_host, long n_temps, long particles, uint64_t seed,
                     const double* noise_host, long n_noise, const double* uniforms_host, long n_uniforms, double* theta_host,
                     double* loglike_host, double* logprior_host, double* weights_host, double* ess_host, double* lognc4_host);

/* ---- model level: nonLinPMMH (reference src/nonLinPMMH.cpp:37-153, R/nonLinPMMH.R) ------------------------------------ */

/* Progress hook: the reference prints to Rcpp::Rcout every msg_freq iterations (:116-138); same quantities. */
typedef void (*smcb200_pmmh_progress)(long iteration, long iterations, double mean_sigv, double mean_sigw, double loglike,
                                      double logprior, double acceptance_rate, void* user);

/* Particle-marginal Metropolis-Hastings: `iterations` random-walk proposals on (sigv, sigw) starting at (10, 10); each
 * evaluates the likelihood with a `particles`-particle bootstrap filter over data[0..T).  Reference behaviour:
 * resample_mode = SMCB200_MULTINOMIAL, threshold = 0.5 (:77).  Outputs have iterations+1 entries, as the reference's
 * DataFrame(samples_sigv, samples_sigw, loglike, logprior, MH_acceptance_rate).
 * Optional injected draws: innov_v/innov_w/unif [iterations] (the pre-drawn outer-chain randomness, :72-74, already scaled:
 * 0.15 z, 0.08 z); filter_noise [(iterations+1) * particles * T] and filter_unif [(iterations+1) * per-filter count] for the
 * inner filters (layouts as in smcb200_pfNonlinBS).  Independent chains are independent calls (one per GPU). */
int smcb200_nonLinPMMH(const double* data_host, long T, long particles, long iterations, uint64_t seed, int resample_mode,
                       double threshold, const double* innov_v_host, const double* innov_w_host, const double* unif_host,
                       const double* filter_noise_host, const double* filter_unif_host, smcb200_pmmh_progress progress,
                       long msg_freq, void* user, double* samples_sigv_host, double* samples_sigw_host, double* loglike_host,
                       double* logprior_host, double* accept_rate_host);

/* LinReg_impl (reference src/LinReg.cpp:46-84): data-annealing SMC sampler, one observation per step, MULTINOMIAL / 0.5,
 * 10 random-walk MH repeats per step.  Outputs: theta [N x 3] column-major, weights [N] (normalised), logNC.
 * Injected draws: noise / uniforms as in smcb200_LinRegLA (Initialise 2 normals + 3 uniforms per particle, then per
 * repeat and particle 3 normals + 1 uniform); resample_uniforms [n_obs * N] indexed by step for the multinomial draws. */
int smcb200_LinReg(const double* data_host, long n_obs, long particles, uint64_t seed, const double* noise_host, long n_noise,
                   const double* uniforms_host, long n_uniforms, const double* resample_uniforms_host, double* theta_host,
                   double* weights_host, double* lognc_host);

/* ---- model level: blockpfGaussianOpt (reference src/blockpfgaussianopt.cpp:41-71, R/blockpfGaussianOpt.R) ------------- */

/* Block-sampling particle filter with the optimal Gaussian block proposal; each particle is a whole path.
 * SYSTEMATIC resampling at ESS < particles/2 (:53).  Outputs mirror List(weight, values, logNC) (:70):
 * weight_host [particles] = GetParticleWeight() (normalised), values_host [particles x n_data] column-major (row i is the
 * path of particle i), lognc_host [1].  Optional diagnostics: ess_host [n_data] (ESS seen by each step's resampling
 * decision), resampled_host [n_data], ancestors_host [n_data][particles] (uRSIndices of each step, identity rows when the
 * step did not resample).
 * Optional injected draws: noise_host [n_data][lag][particles] standard normals, entry (t, k, i) is the k-th draw of
 * particle i's move at time t (k = 0 samples value[t], k >= 1 the backward draws in the reference's order, :121-127);
 * uniforms_host [n_data] systematic base uniforms by step.  NULL -> Philox(seed). */
int smcb200_blockpfGaussianOpt(const double* data_host, long n_data, long particles, long lag, uint64_t seed,
                               const double* noise_host, const double* uniforms_host, double* weight_host, double* values_host,
                               double* lognc_host, double* ess_host, int* resampled_host, uint32_t* ancestors_host);

/* Same filter, path matrix left on the device (replaces nothing in the reference, whose List(weight, values, logNC) always
 * returns N x T doubles to R, src/blockpfgaussianopt.cpp:70): values_dev is a DEVICE pointer to particles x n_data doubles,
 * column-major like values_host.  The other outputs are host buffers as above (any of them may be NULL except values_dev). */
int smcb200_blockpfGaussianOpt_device(const double* data_host, long n_data, long particles, long lag, uint64_t seed,
                                      double* weight_host, double* values_dev, double* lognc_host, double* ess_host,
                                      int* resampled_host);

/* ---- conditional SMC (reference inst/include/conditionalSampler.h, src/cSMCexamples.cpp, sampler.h:445-476) ------------ */

/* Conditional bootstrap particle filters on the scalar linear-Gaussian model of the reference example
 * (x_t = phi x_{t-1} + N(0, varStateEvol), x_0 = stateInit, y_t ~ N(x_t, varObs); cSMCexamples.cpp:141-218), HistoryType::AL,
 * resampling at ESS < particles/2 with the conditional multinomial / residual / stratified / systematic schemes of
 * conditionalSampler.h:377-655.
 * Work is organised in CHAINS: a chain is what the reference does with one smc::conditionalSampler object -- runs executed in
 * order, with the sampler's reference-index state carried from run to run (the reference driver uses a single object for every
 * simulation and scheme, :76-125).  Chains are independent and run concurrently (one thread block each).
 * Run r of chain c is entry c * runs_per_chain + r of modes_host (SMCB200_* resampling modes) and reference_host [runs][T].
 * Outputs per run: lognc_host [runs] = GetLogNCPath(); optional refidx_host [runs][T] = GetReferenceValueIndex(t) after the run;
 * optional AL history ancestors_host / values_host / logw_host [runs][T][particles] (History[t].GetAIndices(), post-resampling
 * particle values, normalised log-weights) and resampled_host [runs][T].
 * Optional injected draws: noise_host [runs][T][particles] standard normals; uniforms_host [runs][T][particles + 2]: entry 0 of a
 * step draws K_t, entry 1 the systematic offset, entry 2 + i belongs to offspring slot i.  NULL -> Philox(seed). */
int smcb200_conditionalBPF(const double* data_host, long T, long particles, double phi, double varStateEvol, double stateInit,
                           double varObs, long n_chains, long runs_per_chain, const int* modes_host, const double* reference_host,
                           uint64_t seed, const double* noise_host, const double* uniforms_host, double* lognc_host,
                           uint32_t* refidx_host, uint32_t* ancestors_host, double* values_host, double* logw_host,
                           int* resampled_host);

/* Plain bootstrap filter on the same model (SamplerBPF of cSMCexamples.cpp:66-69,82-99): log normalising constant of one run.
 * Injected draws as in smcb200_pfNonlinBS (noise [T][particles]; uniforms by step for the chosen mode). */
int smcb200_lgssBPF(const double* data_host, long T, long particles, double phi, double varStateEvol, double stateInit, double varObs,
                    int resample_mode, uint64_t seed, const double* noise_host, const double* uniforms_host, double* lognc_host);

/* compareNCestimates_imp(data, lParticleNum, simNum, parInits{phi, varStateEvol, stateInit, varObs}, referenceTraj [T x simNum])
 * (cSMCexamples.cpp:36-139): smcOut_host / csmcOut_host [simNum x 4] column-major, columns multinomial, residual, stratified,
 * systematic.  independent_sims = 0 reproduces the reference's single conditional sampler object (one chain of 4 simNum runs,
 * sequential by construction); independent_sims = 1 gives every simulation a fresh sampler (simNum concurrent chains of 4). */
int smcb200_compareNCestimates(const double* data_host, long T, long particles, int simNum, double phi, double varStateEvol,
                               double stateInit, double varObs, const double* referenceTraj_host, uint64_t seed, int independent_sims,
                               double* smcOut_host, double* csmcOut_host);

/* Ancestral lines of an AL history: lines_host [T][particles], column n = sampler::GetALineInd(n) (sampler.h:445-459);
 * with values_host [T][particles] (scalar particle values per step) also line_values_host = GetALineSpace(n) (:461-476). */
int smcb200_ancestor_lines(const uint32_t* ancestors_host, long T, long particles, const double* values_host, uint32_t* lines_host,
                           double* line_values_host);

#ifdef __cplusplus
}
#endif
#endif /* SMCB200_H */
