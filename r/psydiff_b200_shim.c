/* psydiff_b200_shim.c — the .Call layer an R build of psydiff would link against libpsydiff_b200.so.
 *
 * Replaces the Rcpp-generated glue of the reference (src/RcppExports.cpp:15-321) for the SR-UKF likelihood / gradient
 * path and keeps the R-level names (R/RcppExports.R:12-257; fitModel / getGradients / getGradient of
 * R/modelTemplate.R:165-166, 853-854, 932-933).  R is not installed in the build image of this repository, so this file
 * is compiled only by a maintainer with R headers:
 *     R CMD SHLIB psydiff_b200_shim.c -I../include -L../psydiff_b200 -lpsydiff_b200
 * tests/test_host.py checks that every psy_* function used here is declared in include/psydiff_b200.h.
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include "psydiff_b200.h"

#define CHECK(rc) do { if ((rc) != PSY_OK) Rf_error("%s", psy_last_error()); } while (0)

static void model_finalizer(SEXP p) {
  psy_model* m = (psy_model*)R_ExternalPtrAddr(p);
  if (m) { psy_model_destroy(m); R_ClearExternalPtr(p); }
}
static psy_model* model_of(SEXP h) {
  psy_model* m = (psy_model*)R_ExternalPtrAddr(h);
  if (!m) Rf_error("model is not compiled: call compileModel(model) first");
  return m;
}

/* compileModel(psydiffModel) — R/prepareModel.R:553-562 */
SEXP R_psy_compile(SEXP cuda_src, SEXP incdir, SEXP cachedir, SEXP n, SEXP m, SEXP nfree, SEXP obs, SEXP dt,
                   SEXP person_off, SEXP person_id, SEXP n_theta, SEXP theta_index) {
  char cubin[4096];
  CHECK(psy_model_compile(CHAR(STRING_ELT(cuda_src, 0)), CHAR(STRING_ELT(incdir, 0)), CHAR(STRING_ELT(cachedir, 0)), "",
                          cubin, sizeof cubin));
  psy_model* mdl = psy_model_create(asInteger(n), asInteger(m), asInteger(nfree), cubin);
  if (!mdl) Rf_error("%s", psy_last_error());
  const int N = LENGTH(person_id);
  int64_t* off = (int64_t*)R_alloc((size_t)N + 1, sizeof(int64_t));
  for (int i = 0; i <= N; i++) off[i] = (int64_t)REAL(person_off)[i];
  int rc = psy_upload(mdl, (int64_t)LENGTH(dt), N, REAL(obs), REAL(dt), off, INTEGER(person_id));
  if (rc == PSY_OK) rc = psy_set_slots(mdl, asInteger(n_theta), INTEGER(theta_index));
  if (rc != PSY_OK) { psy_model_destroy(mdl); Rf_error("%s", psy_last_error()); }
  SEXP p = PROTECT(R_MakeExternalPtr(mdl, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, model_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}

/* a changed model$integrateFunction / srUpdate needs the matching kernel (R/modelTemplate.R:174) */
SEXP R_psy_load_module(SEXP h, SEXP cuda_src, SEXP incdir, SEXP cachedir) {
  char cubin[4096];
  CHECK(psy_model_compile(CHAR(STRING_ELT(cuda_src, 0)), CHAR(STRING_ELT(incdir, 0)), CHAR(STRING_ELT(cachedir, 0)), "",
                          cubin, sizeof cubin));
  CHECK(psy_model_load_module(model_of(h), cubin));
  return R_NilValue;
}

static void push_settings(psy_model* mdl, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr) {
  CHECK(psy_model_settings(mdl, REAL(settings)[0], REAL(settings)[1], REAL(settings)[2], asInteger(stepper),
                           REAL(timeStep), LENGTH(timeStep), asLogical(sr)));
}

/* fitModel(psydiffModel, skipUpdate) — R/modelTemplate.R:165-420; returns list(latentScores, predictedManifest, m2LL) */
SEXP R_psy_fit(SEXP h, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP skip, SEXP N, SEXP rows,
               SEXP n, SEXP m) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  SEXP m2 = PROTECT(allocVector(REALSXP, asInteger(N)));
  SEXP ls = PROTECT(allocMatrix(REALSXP, asInteger(rows), asInteger(n)));
  SEXP pm = PROTECT(allocMatrix(REALSXP, asInteger(rows), asInteger(m)));
  CHECK(psy_fit(mdl, REAL(theta), asLogical(skip), REAL(m2), REAL(ls), REAL(pm)));
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, ls);
  SET_VECTOR_ELT(out, 1, pm);
  SET_VECTOR_ELT(out, 2, m2);
  SEXP nm = PROTECT(allocVector(STRSXP, 3));
  SET_STRING_ELT(nm, 0, mkChar("latentScores"));
  SET_STRING_ELT(nm, 1, mkChar("predictedManifest"));
  SET_STRING_ELT(nm, 2, mkChar("m2LL"));
  setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* getGradients(psydiffModel) — R/modelTemplate.R:852-930; direction: 0 central, 1 left, 2 right */
SEXP R_psy_gradients(SEXP h, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP eps, SEXP direction) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  SEXP g = PROTECT(allocVector(REALSXP, LENGTH(theta)));
  double total;
  CHECK(psy_gradients(mdl, REAL(theta), REAL(eps), LENGTH(eps), asInteger(direction), REAL(g), &total));
  UNPROTECT(1);
  return g;
}

/* getGradient(psydiffModel, currentParameterValues, par) — R/modelTemplate.R:932-1008 (par is 0-based) */
SEXP R_psy_gradient(SEXP h, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP par, SEXP eps, SEXP direction) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  double g;
  CHECK(psy_gradient(mdl, REAL(theta), asInteger(par), REAL(eps), LENGTH(eps), asInteger(direction), &g));
  return ScalarReal(g);
}

/* ---- the exported primitives, same R signatures as R/RcppExports.R ------------------------------------------------ */
SEXP _psydiff_getMeanWeights(SEXP n, SEXP alpha, SEXP beta, SEXP kappa) {          /* R/RcppExports.R: getMeanWeights */
  SEXP w = PROTECT(allocVector(REALSXP, 2 * asInteger(n) + 1));
  psy_getMeanWeights(asInteger(n), asReal(alpha), asReal(beta), asReal(kappa), REAL(w));
  UNPROTECT(1);
  return w;
}
SEXP _psydiff_getCovWeights(SEXP n, SEXP alpha, SEXP beta, SEXP kappa) {
  SEXP w = PROTECT(allocVector(REALSXP, 2 * asInteger(n) + 1));
  psy_getCovWeights(asInteger(n), asReal(alpha), asReal(beta), asReal(kappa), REAL(w));
  UNPROTECT(1);
  return w;
}
SEXP _psydiff_getSigmaPoints(SEXP m, SEXP A, SEXP covarianceIsRoot, SEXP c) {      /* R/RcppExports.R:49 */
  const int n = LENGTH(m);
  SEXP X = PROTECT(allocMatrix(REALSXP, n, 2 * n + 1));
  if (psy_getSigmaPoints(n, REAL(m), REAL(A), asLogical(covarianceIsRoot), asReal(c), REAL(X)) != PSY_OK)
    Rf_error("chol(): decomposition failed");
  UNPROTECT(1);
  return X;
}
SEXP _psydiff_cholupdate(SEXP L, SEXP x, SEXP v, SEXP direction) {                 /* R/RcppExports.R:227 */
  SEXP out = PROTECT(duplicate(L));
  if (psy_cholupdate(nrows(L), REAL(out), REAL(x), asReal(v), CHAR(STRING_ELT(direction, 0))) != PSY_OK)
    Rf_error("Unknown direction in cholupdate. Possible are downdate and update.");   /* psydiffInternals.h:145 */
  UNPROTECT(1);
  return out;
}
SEXP _psydiff_qr_(SEXP X) {                                                        /* R/RcppExports.R:237 */
  SEXP R = PROTECT(allocMatrix(REALSXP, ncols(X), ncols(X)));
  psy_qr(nrows(X), ncols(X), REAL(X), REAL(R));
  UNPROTECT(1);
  return R;
}
SEXP _psydiff_computeIndividualM2LL(SEXP nObservedVariables, SEXP rawData, SEXP expectedMeans, SEXP expectedCovariance) {
  double out;                                                                      /* R/RcppExports.R:214 */
  CHECK(psy_computeIndividualM2LL(asInteger(nObservedVariables), REAL(rawData), REAL(expectedMeans), REAL(expectedCovariance), &out));
  return ScalarReal(out);
}
SEXP _psydiff_predictSquareRootOfS(SEXP Y_, SEXP mu, SEXP Rchol, SEXP covWeight_0, SEXP covWeight_1) {
  const int m = nrows(Y_), s = ncols(Y_);
  SEXP S = PROTECT(allocMatrix(REALSXP, m, m));
  psy_predictSquareRootOfS(m, s, REAL(Y_), REAL(mu), REAL(Rchol), asReal(covWeight_0), asReal(covWeight_1), REAL(S));
  UNPROTECT(1);
  return S;
}
/* getWMatrix, getAFromSigmaPoints, computePFromSigmaPoints, getPhi, predictMu, predictS, predictC, computeK, updateM,
 * updateP, computeIndividualM2LLChol, computeK_SR, logChol2Chol follow the same pattern over psy_getWMatrix,
 * psy_getAFromSigmaPoints, psy_computePFromSigmaPoints, psy_getPhi, psy_predictMu, psy_predictS, psy_predictC,
 * psy_computeK, psy_updateM, psy_updateP, psy_computeIndividualM2LLChol, psy_computeK_SR, psy_logChol2Chol. */
