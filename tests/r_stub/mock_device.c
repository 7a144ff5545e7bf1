/* tests/r_stub/mock_device.c — TEST INFRASTRUCTURE: stand-in for the DEVICE entry points of libpsydiff_b200 (compile, create,
 * upload, slots, settings, fit, gradients) with trivial deterministic arithmetic, so that the .Call shim's whole compile /
 * sync / fit / gradient flow — argument plumbing, by-reference side effects, re-upload logic, naming, finalizer — can be
 * executed by the CPU test suite (tests/test_r_shim.py links a second copy of the shim against this file, listed before the
 * real library so these definitions win).  It computes nothing of the filter; device parity of the same flow is the gpu-marked
 * test.  Error texts are the real library's. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "psydiff_b200.h"

struct psy_model {
  int n, m, F, P, N;
  int64_t rows;
  int uploaded, slotted, settings, uploads, destroyed_marker;
  double checksum;  /* Σ of the finite observations, so that results depend on WHICH dataset is resident */
  double alpha;
  int stepper, sr;
  int n_dev;        /* devices in use (psy_model_set_devices) */
  int module_loads; /* psy_model_load_module calls */
};

static char g_err[512];
static int g_live = 0;
static int fail(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
const char* psy_last_error(void) { return g_err; }
int psy_mock_live_models(void) { return g_live; }
int psy_mock_uploads(const psy_model* M) { return M->uploads; }

int psy_model_compile(const char* src, const char* inc, const char* cache, const char* flags, char* out, size_t cap) {
  (void)inc; (void)cache; (void)flags;
  if (!src || !strstr(src, "PSY_INSTANTIATE")) return fail(PSY_ERR_COMPILE, "model does not compile for the device");
  snprintf(out, cap, "mock.cubin");
  return PSY_OK;
}
psy_model* psy_model_create(int n, int m, int nfree, const char* cubin) {
  if (!cubin || strcmp(cubin, "mock.cubin") != 0) { fail(PSY_ERR_ARG, "bad cubin"); return NULL; }
  psy_model* M = (psy_model*)calloc(1, sizeof *M);
  M->n = n; M->m = m; M->F = nfree; M->n_dev = 1;
  g_live++;
  return M;
}
void psy_model_destroy(psy_model* M) { if (M) { g_live--; free(M); } }
int psy_model_load_module(psy_model* M, const char* cubin) { (void)cubin; if (M) M->module_loads++; return PSY_OK; }
int psy_mock_module_loads(const psy_model* M) { return M->module_loads; }
/* a box with 4 devices */
int psy_device_count(void) { return 4; }
int psy_model_set_devices(psy_model* M, int n, const int* ids) {
  (void)ids;
  if (!M || n < 0) return fail(PSY_ERR_ARG, "psy_model_set_devices: bad arguments");
  if (n > 4) return fail(PSY_ERR_ARG, "psy_model_set_devices: more devices asked for than visible");
  M->n_dev = n == 0 ? 4 : n;
  return PSY_OK;
}
int psy_model_devices(const psy_model* M, int* ids, int cap) {
  for (int k = 0; k < M->n_dev && k < cap; k++) ids[k] = k;
  return M->n_dev;
}
int psy_model_settings(psy_model* M, double alpha, double beta, double kappa, int stepper, const double* ts, int nts, int sr) {
  (void)beta; (void)kappa;
  if (!M || !ts || nts < 1) return fail(PSY_ERR_ARG, "psy_model_settings: bad arguments");
  M->alpha = alpha; M->stepper = stepper; M->sr = sr; M->settings = 1;
  return PSY_OK;
}
int psy_upload(psy_model* M, int64_t rows, int N, const double* obs, const double* dt, const int64_t* off, const int* pid) {
  if (!M || rows < 1 || N < 1 || !obs || !dt || !off || !pid) return fail(PSY_ERR_ARG, "psy_upload: bad arguments");
  if (off[0] != 0 || off[N] != rows) return fail(PSY_ERR_ARG, "psy_upload: person_off must run from 0 to rows");
  M->rows = rows; M->N = N; M->uploaded = 1; M->slotted = 0; M->uploads++;
  M->checksum = 0.0;
  for (int64_t i = 0; i < rows * M->m; i++) if (obs[i] == obs[i]) M->checksum += obs[i];
  return PSY_OK;
}
int psy_set_slots(psy_model* M, int P, const int* tidx) {
  if (!M || !M->uploaded) return fail(PSY_ERR_STATE, "psy_set_slots: call psy_upload first");
  for (long i = 0; i < (long)M->N * M->F; i++)
    if (tidx[i] < 0 || tidx[i] >= P) return fail(PSY_ERR_ARG, "psy_set_slots: theta_index out of range");
  M->P = P; M->slotted = 1;
  return PSY_OK;
}
static double tsum(const psy_model* M, const double* th) { double s = 0; for (int i = 0; i < M->P; i++) s += th[i]; return s; }
int psy_fit(psy_model* M, const double* theta, int skip, double* m2ll, double* lat, double* pred) {
  if (!M || !m2ll) return fail(PSY_ERR_ARG, "psy_fit: bad arguments");
  if (!M->uploaded || !M->slotted) return fail(PSY_ERR_STATE, "psy_fit: call psy_upload and psy_set_slots first");
  if (!M->settings) return fail(PSY_ERR_STATE, "psy_model_settings has not been called");
  for (int p = 0; p < M->N; p++) m2ll[p] = skip ? 0.0 : M->checksum + p + tsum(M, theta);
  if (lat) for (int64_t i = 0; i < M->rows * M->n; i++) lat[i] = 1000.0 + (double)i;
  if (pred) for (int64_t i = 0; i < M->rows * M->m; i++) pred[i] = 2000.0 + (double)i;
  return PSY_OK;
}
int psy_gradients(psy_model* M, const double* theta, const double* eps, int n_eps, int direction, double* grad, double* total) {
  if (!M || !eps || n_eps < 1 || !grad) return fail(PSY_ERR_ARG, "psy_gradients: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  if (!M->uploaded || !M->slotted) return fail(PSY_ERR_STATE, "psy_grad_level: call psy_upload and psy_set_slots first");
  for (int t = 0; t < M->P; t++) grad[t] = 2.0 * theta[t] + M->N + 0.001 * direction + eps[0];
  if (total) *total = M->checksum;
  return PSY_OK;
}
int psy_gradient(psy_model* M, const double* theta, int par, const double* eps, int n_eps, int direction, double* grad) {
  if (!M || par < 0 || par >= M->P) return fail(PSY_ERR_ARG, "psy_gradient: parameter index out of range");
  double* g = (double*)malloc(sizeof(double) * (size_t)M->P);
  int rc = psy_gradients(M, theta, eps, n_eps, direction, g, NULL);
  if (!rc) *grad = g[par];
  free(g);
  return rc;
}
