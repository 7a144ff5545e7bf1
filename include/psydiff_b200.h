/* psydiff_b200.h — C ABI of libpsydiff_b200.so: the B200-native drop-in for psydiff's SR-UKF likelihood and
 * finite-difference gradient path.  Plain pointers and sizes only; matrices cross the boundary in R's layout
 * (column-major doubles), exactly as they reach the reference's Rcpp glue (src/RcppExports.cpp:15-321).
 *
 * All citations are file:line relative to the reference tree (jhorzek/psydiff).
 *
 * Conventions
 *   - every function returning int returns 0 on success, >0 on error; psy_last_error() then holds the message
 *     (the reference raises R errors through Rcpp::stop / BEGIN_RCPP, src/RcppExports.cpp:18-26);
 *   - numeric failure is NOT an error: the affected -2LL / gradient entry is NaN (callers test anyNA,
 *     R/optimWrapper.R:153, R/GIST.R:46-47);
 *   - single host thread per model handle, like the R API (R/modelTemplate.R:323);
 *   - the library has no CPU fallback: every compute entry point fails with PSY_ERR_CUDA when no sm_100 device
 *     (or the compiled model module) is available.
 */
#ifndef PSYDIFF_B200_H
#define PSYDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSY_OK 0
#define PSY_ERR_ARG 1
#define PSY_ERR_CUDA 2
#define PSY_ERR_COMPILE 3
#define PSY_ERR_STATE 4

#define PSY_STEPPER_RK4 0     /* integrateFunction = "rk4"                 (R/modelTemplate.R:326-329) */
#define PSY_STEPPER_DOPRI5 1  /* integrateFunction = "runge_kutta_dopri5"  (R/modelTemplate.R:330-333) */
#define PSY_STEPPER_DEFAULT 2 /* "default": rk4, dopri5 only on a non-finite state (R/modelTemplate.R:334-344) */

#define PSY_DIR_CENTRAL 0 /* direction = "central" (R/modelTemplate.R:868) */
#define PSY_DIR_LEFT 1
#define PSY_DIR_RIGHT 2

typedef struct psy_model psy_model;

/* thread-local message of the last failing call on this thread */
const char* psy_last_error(void);
int psy_version(void);
/* steps Boost.odeint's integrate_const(runge_kutta4, sys, x, 0, t1, h) takes (R/modelTemplate.R:327-329): the count the
 * kernel uses per observation interval; the remainder of the interval is not integrated. */
int psy_rk4_step_count(double t1, double h);

/* ---- newPsydiff's source assembly (R/prepareModel.R:479-512, 737-770; template R/modelTemplate.R:5-1011) ----------
 * The reference pastes prepareEquations(latentEquations) and prepareEquations(manifestEquations) into modelTemplate()
 * and stores the text in model$cppCode.  psy_emit_model_source does the same for the CUDA translation unit: every entry
 * of model$pars$parameterList the statements name is declared as a fixed-size container whose free elements read the
 * instance's slot array, and m0 / A0LogChol / LLogChol / RLogChol become the initial-state functions of the model struct.
 * Pure host string work, callable from any language; the result goes to psy_model_compile. */
typedef struct psy_container {
  const char* name;     /* name in model$pars$parameterList; "m0", "A0LogChol", "LLogChol", "RLogChol" must be present */
  int is_scalar;        /* targetClass "double" (1) or "matrix" (0) (R/prepareModel.R:646-690) */
  int rows, cols;
  const double* values; /* column-major start values */
  const int* slots;     /* column-major: 0-based row of parameterTableInit for a free element, -1 for a fixed one */
} psy_container;
/* out/cap: destination buffer; *needed (may be NULL) receives the size including the terminator; out == NULL with
 * needed != NULL only queries the size. */
int psy_emit_model_source(int n, int m, int n_slots, const char* latent_equations, const char* manifest_equations,
                          const psy_container* containers, int n_containers, int sr_update, int stepper, char* out,
                          size_t cap, size_t* needed);
/* modelTemplate() (R/modelTemplate.R:5): the template text with its upper-case placeholders */
int psy_model_template(int sr_update, char* out, size_t cap, size_t* needed);
/* Jacobian sparsity of the drift as read off the statements: bit r of mask[i] set = dx_i may depend on x_r (conservative:
 * all ones for anything but plain element statements).  The kernel skips sigma-point columns that cannot move a component. */
int psy_drift_dependency(int n, const char* latent_equations, const psy_container* containers, int n_containers,
                         unsigned* mask);

/* ---- compileModel (R/prepareModel.R:553-562) ---------------------------------------------------------------
 * The reference writes model$cppCode to a temp file and hands it to Rcpp::sourceCpp (host g++).  Here the
 * generated CUDA translation unit (user drift + measurement statements spliced into the filter kernel template,
 * cf. R/prepareModel.R:479-512) is compiled for sm_100a and cached by content hash: with an `nvcc -cubin` subprocess by
 * default, or in-process through NVRTC (dlopen'ed libnvrtc, no nvcc needed) when the environment has PSY_COMPILER=nvrtc.
 * No GPU is needed for this call.  cubin_path receives the path of the compiled module. */
int psy_model_compile(const char* cuda_source, const char* include_dir, const char* cache_dir,
                      const char* extra_flags, char* cubin_path, size_t cubin_path_cap);

/* Create a model handle for n latent / m manifest variables and nfree free parameter slots (rows of the
 * per-person parameterTable, R/prepareModel.R:467-472), and load the compiled module onto the current device. */
psy_model* psy_model_create(int n, int m, int nfree, const char* cubin_path);
void psy_model_destroy(psy_model* mdl);
/* Replace the handle's kernel module (same n, m, slots) keeping the resident dataset and slot map — used when
 * model$integrateFunction or srUpdate change after compileModel: the stepper and the update variant are compile-time
 * properties of the generated kernel. */
int psy_model_load_module(psy_model* mdl, const char* cubin_path);

/* ---- devices: getGradientsParallel / compileParallelGradients (R/prepareModel.R:572-616) inside the library ----------
 * The reference's only parallel path starts nCores PSOCK workers and ships them the model on every call.  Here ONE handle and
 * ONE host thread drive several GPUs: persons are independent units (R/modelTemplate.R:244), so psy_upload cuts the dataset into
 * contiguous person blocks balanced by filter work (Σ_t 1 + dt_t/h), every device filters its block on its own stream, and the
 * per-parameter partial sums (5P+2 doubles per device) are added on the first device after peer copies over NVLink, in
 * device order (deterministic).  A handle starts on the device that is current at psy_model_create.
 *   n_devices = 0: all visible devices; device_ids = NULL: the creating device plus the next n_devices-1 visible ones.
 * May be called at any time: a resident dataset and slot map are re-distributed from the library's pinned host copy. */
int psy_device_count(void);
int psy_model_set_devices(psy_model* mdl, int n_devices, const int* device_ids);
/* number of devices in use (<= the number asked for when there are fewer persons); ids into device_ids[0..cap) */
int psy_model_devices(const psy_model* mdl, int* device_ids, int cap);
/* one device's share: its id, first person and person count, filter instances of a full sweep, kernel ms of the last launch */
int psy_shard_info(const psy_model* mdl, int shard, int* device, int* first_person, int* n_persons, int64_t* n_instances,
                   double* last_kernel_ms);

/* Settings that fitModel reads back from the model list on every call (R/modelTemplate.R:169-176):
 * alpha, beta, kappa, timeStep[], integrateFunction; sr = srUpdate of newPsydiff (R/prepareModel.R:403). */
int psy_model_settings(psy_model* mdl, double alpha, double beta, double kappa, int stepper,
                       const double* time_step, int n_time_step, int sr);

/* model$data (R/prepareModel.R:413-426, R/modelTemplate.R:201-206): observations rows x m COLUMN-major with
 * NaN/NA = missing, dt[rows], persons as N contiguous row blocks (person_off[N+1]) with their id values.
 * The data stays resident in HBM across all later calls.  Calling it again with the same number of persons replaces the rows
 * and keeps the slot map; with one device the instance / pair / chunk lists of psy_set_slots are kept too (they depend on the
 * slot map, not on the rows), with several devices they are rebuilt because the work-balanced person cuts may move. */
int psy_upload(psy_model* mdl, int64_t rows, int n_persons, const double* observations_colmajor, const double* dt,
               const int64_t* person_off, const int* person_id);

/* The integer replacement of the string-keyed parameterTable (setParameterValues_C / setParameterList_C,
 * inst/include/psydiffBasic.h:4-33, 63-121): theta_index[p*nfree + k] = position in theta (0..P-1) of the
 * label that person p's free slot k carries after grouping (makeGroups, R/prepareModel.R:704-734). */
int psy_set_slots(psy_model* mdl, int n_theta, const int* theta_index);

/* fitModel(psydiffModel, skipUpdate) (R/modelTemplate.R:165-420): per-person -2LL into m2ll[N]; optional
 * latentScores (rows x n) and predictedManifest (rows x m), COLUMN-major like the R matrices, may be NULL. */
int psy_fit(psy_model* mdl, const double* theta, int skip_update, double* m2ll, double* latent_scores,
            double* predicted_manifest);

/* fitModel for several parameter vectors in one launch (thetas: n_sets x P row-major): totals[k] = Σ_p -2LL over the persons
 * with a finite value, nonfinite[k] = how many persons failed.  What a caller that evaluates psydiffFitInternal
 * (R/optimWrapper.R:146-157) at several candidate points — GIST's inner line search (R/GIST.R:121-190), bestOfN
 * (R/bestOfN.R:23-38) — can use instead of n_sets sequential fitModel calls. */
int psy_fit_many(psy_model* mdl, const double* thetas, int n_sets, int skip_update, double* totals, double* nonfinite);

/* getGradients / getGradientsParallel (R/modelTemplate.R:852-930, R/prepareModel.R:608-616): the whole finite-
 * difference sweep as one batched launch per eps level.  grad[P] (NaN where both eps[] fallbacks fail,
 * R/modelTemplate.R:885-898); m2ll_total (may be NULL) receives Σ -2LL at theta. */
int psy_gradients(psy_model* mdl, const double* theta, const double* eps, int n_eps, int direction, double* grad,
                  double* m2ll_total);

/* getGradient(psydiffModel, currentParameterValues, par) (R/modelTemplate.R:932-1008): one parameter. */
int psy_gradient(psy_model* mdl, const double* theta, int par, const double* eps, int n_eps, int direction,
                 double* grad);

/* ---- multi-PROCESS building blocks (one rank per GPU or per node; one all-reduce per eps level) ---------------------
 * psy_grad_level runs this rank's instances at one eps value and leaves the partial-sum vector
 *   [ Σf0, #nonfinite f0, D+[P], D-[P], nf+[P], nf-[P], nbt[P] ]           (length psy_partial_len = 5P + 2)
 * in device memory — on the handle's first device, already summed over the handle's own devices — (D± = Σ_p (f±_p − f0_p), nf± = # non-finite f±, nbt = # touched persons with non-finite f0).
 * need_side (2P bytes, [par*2 + side], NULL = all) restricts the launch to the (parameter, side) pairs still
 * unresolved — the eps[] fallback of R/modelTemplate.R:885-898.  The caller all-reduces (sum) the vector across
 * ranks (NCCL on the device pointer) and feeds the HOST copy of the reduced vector to psy_grad_resolve. */
int64_t psy_partial_len(const psy_model* mdl);
int psy_grad_level(psy_model* mdl, const double* theta, double eps, int direction, const unsigned char* need_side,
                   void** partial_device);
/* Pure host bookkeeping, identical on every rank: consumes the reduced partial of eps level `level`, updates
 * state (which sides are resolved, the eps actually used, R/modelTemplate.R:893, 914), writes need_side for the
 * next level and returns in *done whether the sweep is finished (all resolved or eps[] exhausted). */
int psy_grad_resolve(psy_model* mdl, const double* partial_host, double eps, int level, int n_eps, int direction,
                     unsigned char* need_side, int* done);
int psy_grad_finish(psy_model* mdl, int direction, double* grad, double* m2ll_total);
/* The same bookkeeping without a device handle (psy_grad_resolve / psy_grad_finish delegate to it): usable on any
 * host, e.g. by a coordinator process or in CPU tests of the multi-rank protocol. */
typedef struct psy_sweep psy_sweep;
psy_sweep* psy_sweep_create(int n_theta);
void psy_sweep_destroy(psy_sweep* s);
int psy_sweep_resolve(psy_sweep* s, const double* partial_host, double eps, int level, int n_eps, int direction,
                      unsigned char* need_side, int* done);
int psy_sweep_finish(psy_sweep* s, int direction, double* grad, double* m2ll_total);

/* ---- instrumentation ------------------------------------------------------------------------------------ */
/* device time (ms, CUDA events on the launch stream) of the last filter-kernel launch, its instance count and
 * the number of kernels this library has launched since the last psy_counters_reset; with several devices the time is the
 * slowest device's (they run concurrently) and the instance count the sum */
double psy_last_kernel_ms(const psy_model* mdl);
int64_t psy_last_kernel_instances(const psy_model* mdl);
int64_t psy_launch_count(void);
void psy_counters_reset(void);
/* number of filter instances one full gradient sweep launches on this rank: Σ_p (2 P_p + 1) */
int64_t psy_n_instances(const psy_model* mdl);
/* dependent-free DFMA throughput of the current device in TFLOP/s (FMA = 2 flops); the FP64 roofline peak */
int psy_measure_fp64_peak(double* tflops);
/* kernel resource usage of the loaded module */
int psy_kernel_info(const psy_model* mdl, int* regs, int* local_bytes, int* max_threads);
/* Lanes that share one filter instance in the RK4 stage loop of the loaded kernel: 1 = one thread per instance, 2 = the lane-pair
   kernel (even n >= 8; every lane still owns one instance in the measurement update).  Diagnostics only; no reference counterpart. */
int psy_kernel_lanes(const psy_model* mdl);

/* ---- the filter primitives the reference exports to R one by one (src/exposeInlineFunctions.cpp:13-308,
 * R signatures in R/RcppExports.R:12-257).  Tiny host-side utilities on column-major buffers; they are NOT on
 * the fit/gradient path (that is the CUDA kernel) and exist so R code calling them keeps working. ----------- */
void psy_getMeanWeights(int n, double alpha, double beta, double kappa, double* mean_weights /*2n+1*/);
void psy_getCovWeights(int n, double alpha, double beta, double kappa, double* cov_weights /*2n+1*/);
void psy_getWMatrix(int s, const double* mean_weights, const double* cov_weights, double* W /*s x s*/);
int psy_getSigmaPoints(int n, const double* m, const double* A, int covariance_is_root, double c, double* X);
void psy_computeMeanFromSigmaPoints(int n, int s, const double* X, const double* mean_weights, double* mean);
void psy_getAFromSigmaPoints(int n, const double* X, const double* mean_weights, double c, double* A);
void psy_computePFromSigmaPoints(int n, const double* X, const double* mean_weights, double c, double* P);
void psy_getPhi(int n, const double* Mx, double* Phi);
void psy_predictMu(int m, int s, const double* Y, const double* mean_weights, double* mu);
void psy_predictS(int m, int s, const double* Y, const double* W, const double* R, double* S);
void psy_predictC(int n, int m, int s, const double* X, const double* Y, const double* W, double* C);
int psy_computeK(int n, int m, const double* C, const double* S, double* K);
void psy_updateM(int n, int m, const double* m_pred, const double* K, const double* residual, double* m_upd);
void psy_updateP(int n, int m, const double* P_pred, const double* K, const double* S, double* P_upd);
int psy_computeIndividualM2LL(int n_observed, const double* raw, const double* expected_means,
                              const double* expected_cov, double* m2ll);
int psy_computeIndividualM2LLChol(int n_observed, const double* raw, const double* expected_means,
                                  const double* expected_cov_chol, double* m2ll);
int psy_cholupdate(int n, double* L, const double* x, double v, const char* direction);
void psy_qr(int rows, int cols, const double* X, double* R /*cols x cols*/);
void psy_predictSquareRootOfS(int m, int s, const double* Y, const double* mu, const double* Rchol, double wc0,
                              double wc1, double* srS);
int psy_computeK_SR(int n, int m, const double* C, const double* srS, double* K);
void psy_logChol2Chol(int n, const double* log_chol, double* chol);

#ifdef __cplusplus
}
#endif
#endif /* PSYDIFF_B200_H */
