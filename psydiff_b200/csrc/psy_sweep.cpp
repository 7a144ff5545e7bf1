// psy_sweep.cpp — host-only bookkeeping of one finite-difference sweep: which (parameter, side) pairs are resolved at which
// eps level, the eps actually used, the final gradient.  This is the eps[] fallback of getGradients / getGradient
// (R/modelTemplate.R:885-898, 905-919, 926) applied to the reduced partial-sum vector of a batched launch; it needs no device
// and is identical on every rank of a multi-process run.
#include <cmath>

#include "psy_internal.h"

#define fail psy_fail

extern "C" {

psy_sweep* psy_sweep_create(int n_theta) {
  if (n_theta < 0) { fail(PSY_ERR_ARG, "psy_sweep_create: bad arguments"); return nullptr; }
  psy_sweep* S = new psy_sweep;
  S->P = n_theta;
  return S;
}
void psy_sweep_destroy(psy_sweep* S) { delete S; }

int psy_sweep_resolve(psy_sweep* S, const double* part, double eps, int level, int n_eps, int direction, unsigned char* need, int* done) {
  if (!S || !part || !need || !done) return fail(PSY_ERR_ARG, "psy_sweep_resolve: bad arguments");
  if (direction < 0 || direction > 2) return fail(PSY_ERR_ARG, "Unknown direction argument. Possible are central, left and right");
  const int P = S->P;
  if (level == 0) {
    S->base_total = part[0];
    S->base_nonfinite = part[1];
    S->Dval.assign((size_t)P * 2, 0.0);
    S->eps_used.assign(P, 0.0);
    S->resolved.assign((size_t)P * 2, 0);
    S->wanted.assign((size_t)P * 2, 0);
    for (int t = 0; t < P; t++) {
      S->wanted[2 * t + 0] = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_LEFT);
      S->wanted[2 * t + 1] = (direction == PSY_DIR_CENTRAL || direction == PSY_DIR_RIGHT);
    }
  }
  if (S->wanted.size() != (size_t)P * 2) return fail(PSY_ERR_STATE, "psy_sweep_resolve: level 0 has not been resolved");
  const double* Dp = part + 2;
  const double* Dm = part + 2 + (size_t)P;
  const double* nfp = part + 2 + 2 * (size_t)P;
  const double* nfm = part + 2 + 3 * (size_t)P;
  const double* nbt = part + 2 + 4 * (size_t)P;
  int pending = 0;
  for (int t = 0; t < P; t++) {
    for (int side = 0; side < 2; side++) {
      const size_t k = 2 * (size_t)t + side;
      need[k] = 0;
      if (!S->wanted[k] || S->resolved[k]) continue;
      const double D = side ? Dp[t] : Dm[t];
      const double nf = side ? nfp[t] : nfm[t];
      // Σ -2LL over ALL persons is finite  <=>  untouched persons finite at base, touched persons finite here
      const bool ok = (S->base_nonfinite - nbt[t] == 0.0) && nf == 0.0 && std::isfinite(D);
      if (ok) {
        S->resolved[k] = 1;
        S->Dval[k] = D;
        S->eps_used[t] += eps;  // R/modelTemplate.R:893, 914
      } else if (level + 1 < n_eps) {
        need[k] = 1;
        pending++;
      }
    }
  }
  *done = (pending == 0);
  return PSY_OK;
}

int psy_sweep_finish(psy_sweep* S, int direction, double* grad, double* total) {
  if (!S || !grad) return fail(PSY_ERR_ARG, "psy_sweep_finish: bad arguments");
  if (S->wanted.size() != (size_t)S->P * 2) return fail(PSY_ERR_STATE, "psy_sweep_finish: nothing resolved");
  (void)direction;
  const double nan = std::nan("");
  for (int t = 0; t < S->P; t++) {
    const bool wl = S->wanted[2 * t], wr = S->wanted[2 * t + 1];
    if ((wl && !S->resolved[2 * t]) || (wr && !S->resolved[2 * t + 1]) || (!wl && !wr)) { grad[t] = nan; continue; }
    if (!(wl && wr) && S->base_nonfinite != 0.0) { grad[t] = nan; continue; }  // one-sided: the other side is Σ f0
    const double dplus = wr ? S->Dval[2 * t + 1] : 0.0, dminus = wl ? S->Dval[2 * t] : 0.0;
    grad[t] = (dplus - dminus) / S->eps_used[t];  // R/modelTemplate.R:926
  }
  if (total) *total = S->base_total;
  return PSY_OK;
}


}  // extern "C"
