/* psydiff_b200_shim.c — the .Call layer an R build of psydiff links against libpsydiff_b200.so.
 *
 * Replaces the Rcpp-generated glue of the reference (src/RcppExports.cpp:15-321): the same 22 registered symbols
 * `_psydiff_<name>` with the same arity (src/RcppExports.cpp:292-316), so R/RcppExports.R:12-257 works unchanged, plus the
 * entry points behind compileModel / fitModel / getGradients / getGradient (R/prepareModel.R:553-562,
 * R/modelTemplate.R:165-166, 853-854, 932-933) and newPsydiff's source assembly (R/prepareModel.R:479-512).
 *
 * R is not installed in the build image of this repository.  A maintainer builds it with
 *     R CMD SHLIB psydiff_b200_shim.c -I../include -L../psydiff_b200 -lpsydiff_b200
 * Here it is compiled against tests/r_stub (a small functional stand-in for the R C API) and EXECUTED by
 * tests/test_r_shim.py, which drives every entry point with stand-in SEXPs; only the public Rf_ API is used.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>
#include "psydiff_b200.h"

#define CHECK(rc) do { if ((rc) != PSY_OK) Rf_error("%s", psy_last_error()); } while (0)

/* ---- helpers -------------------------------------------------------------------------------------------------- */
static SEXP list_get(SEXP list, const char* name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (names != R_NilValue)
    for (int i = 0; i < LENGTH(names); i++)
      if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  Rf_error("element '%s' not found", name);
  return R_NilValue;
}
static double num_at(SEXP v, int i) { return TYPEOF(v) == INTSXP ? (double)INTEGER(v)[i] : REAL(v)[i]; }
static SEXP real_matrix(int r, int c) { return Rf_allocMatrix(REALSXP, r, c); }

static void model_finalizer(SEXP p) {
  psy_model* m = (psy_model*)R_ExternalPtrAddr(p);
  if (m) { psy_model_destroy(m); R_ClearExternalPtr(p); }
}
static psy_model* model_of(SEXP h) {
  psy_model* m = (psy_model*)R_ExternalPtrAddr(h);
  if (!m) Rf_error("model is not compiled: call compileModel(model) first");
  return m;
}

/* ---- newPsydiff's source assembly (R/prepareModel.R:479-512): equation strings + containers -> model$cppCode ------ */
SEXP R_psy_emit_source(SEXP n, SEXP m, SEXP n_slots, SEXP latent, SEXP manifest, SEXP names, SEXP is_scalar, SEXP values,
                       SEXP slots, SEXP sr, SEXP stepper) {
  const int nc = LENGTH(names);
  psy_container* c = (psy_container*)R_alloc((size_t)(nc > 0 ? nc : 1), sizeof(psy_container));
  for (int k = 0; k < nc; k++) {
    SEXP v = VECTOR_ELT(values, k);
    c[k].name = CHAR(STRING_ELT(names, k));
    c[k].is_scalar = LOGICAL(is_scalar)[k];
    c[k].rows = Rf_nrows(v);
    c[k].cols = Rf_ncols(v);
    c[k].values = REAL(v);
    c[k].slots = INTEGER(VECTOR_ELT(slots, k));
  }
  size_t need = 0;
  CHECK(psy_emit_model_source(Rf_asInteger(n), Rf_asInteger(m), Rf_asInteger(n_slots), CHAR(STRING_ELT(latent, 0)),
                              CHAR(STRING_ELT(manifest, 0)), c, nc, Rf_asLogical(sr), Rf_asInteger(stepper), NULL, 0, &need));
  char* buf = R_alloc(need, 1);
  CHECK(psy_emit_model_source(Rf_asInteger(n), Rf_asInteger(m), Rf_asInteger(n_slots), CHAR(STRING_ELT(latent, 0)),
                              CHAR(STRING_ELT(manifest, 0)), c, nc, Rf_asLogical(sr), Rf_asInteger(stepper), buf, need, NULL));
  return Rf_mkString(buf);
}

/* ---- compileModel(psydiffModel) — R/prepareModel.R:553-562 ----------------------------------------------------------
 * The reference's compileModel only loads compiled code; the functions it defines work on ANY model object of the same
 * structure, reading model$data on every call.  Here the dataset is resident in HBM, so the handle remembers WHICH R
 * objects it holds (observations, dt, person, parameterTable$label — kept alive in the external pointer's protected slot)
 * and R_psy_is_current lets the R wrappers re-upload when they are handed different ones.  R's copy-on-modify makes object
 * identity a sound test: a modified dataset is a new object. */
SEXP R_psy_compile(SEXP cuda_src, SEXP incdir, SEXP cachedir, SEXP n, SEXP m, SEXP nfree) {
  char cubin[4096];
  CHECK(psy_model_compile(CHAR(STRING_ELT(cuda_src, 0)), CHAR(STRING_ELT(incdir, 0)), CHAR(STRING_ELT(cachedir, 0)), "",
                          cubin, sizeof cubin));
  psy_model* mdl = psy_model_create(Rf_asInteger(n), Rf_asInteger(m), Rf_asInteger(nfree), cubin);
  if (!mdl) Rf_error("%s", psy_last_error());
  SEXP held = PROTECT(Rf_allocVector(VECSXP, 4));
  SEXP p = PROTECT(R_MakeExternalPtr(mdl, R_NilValue, held));
  R_RegisterCFinalizerEx(p, model_finalizer, TRUE);
  UNPROTECT(2);
  return p;
}

SEXP R_psy_is_current(SEXP h, SEXP obs, SEXP dt, SEXP person, SEXP labels) {
  model_of(h);
  SEXP held = R_ExternalPtrProtected(h);
  const int same = VECTOR_ELT(held, 0) == obs && VECTOR_ELT(held, 1) == dt && VECTOR_ELT(held, 2) == person &&
                   VECTOR_ELT(held, 3) == labels;
  SEXP out = PROTECT(Rf_allocVector(LGLSXP, 1));
  LOGICAL(out)[0] = same;
  UNPROTECT(1);
  return out;
}

/* model$data -> HBM and the slot map (theta_index = match(label, unique(label)) - 1, person-major) */
SEXP R_psy_upload(SEXP h, SEXP obs, SEXP dt, SEXP person, SEXP person_off, SEXP person_id, SEXP n_theta, SEXP theta_index,
                  SEXP labels) {
  psy_model* mdl = model_of(h);
  const int N = LENGTH(person_id);
  int64_t* off = (int64_t*)R_alloc((size_t)N + 1, sizeof(int64_t));
  for (int i = 0; i <= N; i++) off[i] = (int64_t)REAL(person_off)[i];
  CHECK(psy_upload(mdl, (int64_t)LENGTH(dt), N, REAL(obs), REAL(dt), off, INTEGER(person_id)));
  CHECK(psy_set_slots(mdl, Rf_asInteger(n_theta), INTEGER(theta_index)));
  SEXP held = R_ExternalPtrProtected(h);
  SET_VECTOR_ELT(held, 0, obs);
  SET_VECTOR_ELT(held, 1, dt);
  SET_VECTOR_ELT(held, 2, person);
  SET_VECTOR_ELT(held, 3, labels);
  return R_NilValue;
}

/* a changed model$integrateFunction / srUpdate needs the matching kernel (R/modelTemplate.R:174) */
SEXP R_psy_load_module(SEXP h, SEXP cuda_src, SEXP incdir, SEXP cachedir) {
  char cubin[4096];
  CHECK(psy_model_compile(CHAR(STRING_ELT(cuda_src, 0)), CHAR(STRING_ELT(incdir, 0)), CHAR(STRING_ELT(cachedir, 0)), "",
                          cubin, sizeof cubin));
  CHECK(psy_model_load_module(model_of(h), cubin));
  return R_NilValue;
}

/* compileParallelGradients(psydiffModel, nCores) (R/prepareModel.R:572-589): the "workers" are GPUs inside the library — the
 * handle's dataset is spread over min(nCores, visible devices) devices, one host thread (psy_model_set_devices).  Returns the
 * device ids in use. */
SEXP R_psy_set_devices(SEXP h, SEXP n_cores) {
  psy_model* mdl = model_of(h);
  int want = Rf_asInteger(n_cores);
  const int visible = psy_device_count();
  if (visible < 1) Rf_error("%s", psy_last_error());
  if (want < 1) want = 1;
  if (want > visible) want = visible;
  CHECK(psy_model_set_devices(mdl, want, NULL));
  int ids[64];
  const int n = psy_model_devices(mdl, ids, 64);
  SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
  for (int i = 0; i < n; i++) INTEGER(out)[i] = ids[i];
  UNPROTECT(1);
  return out;
}

static void push_settings(psy_model* mdl, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr) {
  CHECK(psy_model_settings(mdl, REAL(settings)[0], REAL(settings)[1], REAL(settings)[2], Rf_asInteger(stepper),
                           REAL(timeStep), LENGTH(timeStep), Rf_asLogical(sr)));
}

/* list[name] = value on the list object itself (what Rcpp's `psydiffModel["m2LL"] = m2LL` does); a missing element is an error */
static void list_set(SEXP list, const char* name, SEXP value) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (names != R_NilValue)
    for (int i = 0; i < LENGTH(names); i++)
      if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) { SET_VECTOR_ELT(list, i, value); return; }
  Rf_error("element '%s' not found", name);
}

/* fitModel(psydiffModel, skipUpdate) — R/modelTemplate.R:165-420.  Returns list(latentScores, predictedManifest, m2LL) AND, like
 * the reference (which works on the Rcpp::List by reference, R/modelTemplate.R:410-414; relied on at R/prepareModel.R:286),
 * overwrites model$m2LL / $latentScores / $predictedManifest in place and clears pars$parameterTable$changed. */
SEXP R_psy_fit(SEXP h, SEXP model, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP skip) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  SEXP obs = list_get(list_get(model, "data"), "observations");
  const int rows = Rf_nrows(obs), m = Rf_ncols(obs), n = Rf_asInteger(list_get(model, "nlatent"));
  const int N = LENGTH(list_get(model, "m2LL"));
  SEXP m2 = PROTECT(Rf_allocVector(REALSXP, N));
  SEXP ls = PROTECT(real_matrix(rows, n));
  SEXP pm = PROTECT(real_matrix(rows, m));
  CHECK(psy_fit(mdl, REAL(theta), Rf_asLogical(skip), REAL(m2), REAL(ls), REAL(pm)));
  /* by-reference side effects on the caller's model object: psydiffModel["m2LL"] = m2LL; ... (R/modelTemplate.R:412-414) */
  list_set(model, "m2LL", m2);
  list_set(model, "latentScores", ls);
  list_set(model, "predictedManifest", pm);
  SEXP changed = list_get(list_get(list_get(model, "pars"), "parameterTable"), "changed");
  for (int i = 0; i < LENGTH(changed); i++) LOGICAL(changed)[i] = 0;      /* :410 */
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, ls);
  SET_VECTOR_ELT(out, 1, pm);
  SET_VECTOR_ELT(out, 2, m2);
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, 3));
  SET_STRING_ELT(nm, 0, Rf_mkChar("latentScores"));
  SET_STRING_ELT(nm, 1, Rf_mkChar("predictedManifest"));
  SET_STRING_ELT(nm, 2, Rf_mkChar("m2LL"));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(5);
  return out;
}

/* getGradients(psydiffModel) — R/modelTemplate.R:852-930; direction: 0 central, 1 left, 2 right.  Named like
 * getParameterValues; the model object is not modified (the reference works on a clone, :855). */
SEXP R_psy_gradients(SEXP h, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP eps, SEXP direction) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  SEXP g = PROTECT(Rf_allocVector(REALSXP, LENGTH(theta)));
  double total;
  CHECK(psy_gradients(mdl, REAL(theta), REAL(eps), LENGTH(eps), Rf_asInteger(direction), REAL(g), &total));
  Rf_setAttrib(g, R_NamesSymbol, Rf_getAttrib(theta, R_NamesSymbol));
  UNPROTECT(1);
  return g;
}

/* getGradient(psydiffModel, currentParameterValues, par) — R/modelTemplate.R:932-1008 (par is 0-based) */
SEXP R_psy_gradient(SEXP h, SEXP settings, SEXP stepper, SEXP timeStep, SEXP sr, SEXP theta, SEXP par, SEXP eps, SEXP direction) {
  psy_model* mdl = model_of(h);
  push_settings(mdl, settings, stepper, timeStep, sr);
  double g;
  CHECK(psy_gradient(mdl, REAL(theta), Rf_asInteger(par), REAL(eps), LENGTH(eps), Rf_asInteger(direction), &g));
  SEXP out = PROTECT(Rf_ScalarReal(g));
  SEXP names = Rf_getAttrib(theta, R_NamesSymbol);
  if (names != R_NilValue && Rf_asInteger(par) >= 0 && Rf_asInteger(par) < LENGTH(names)) {
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 1));
    SET_STRING_ELT(nm, 0, STRING_ELT(names, Rf_asInteger(par)));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}

/* ---- parameter plumbing with the reference's by-reference semantics (inst/include/psydiffBasic.h) --------------------- */
/* setParameterValues(parameterTable, parameterValues, parameterLabels) — psydiffBasic.h:4-33: mutates the data.frame */
SEXP _psydiff_setParameterValues(SEXP parameterTable, SEXP parameterValues, SEXP parameterLabels) {
  if (LENGTH(parameterValues) != LENGTH(parameterLabels))
    Rf_error("Length of parameterValues and parameterLabels does not match");
  SEXP values = list_get(parameterTable, "value"), labels = list_get(parameterTable, "label"),
       changed = list_get(parameterTable, "changed");
  for (int i = 0; i < LENGTH(parameterLabels); i++) {
    const char* lab = CHAR(STRING_ELT(parameterLabels, i));
    const double v = REAL(parameterValues)[i];
    int found = 0;
    for (int j = 0; j < LENGTH(labels); j++)
      if (strcmp(CHAR(STRING_ELT(labels, j)), lab) == 0) {
        found = 1;
        if (REAL(values)[j] != v) {  /* only a changed value raises the flag */
          REAL(values)[j] = v;
          LOGICAL(changed)[j] = 1;
        }
      }
    if (!found) Rf_warning("The following parameter was not found in the parameterTable: %s", lab);
  }
  return R_NilValue;
}

/* getParameterValues(psydiffModel) — psydiffBasic.h:36-58: unique label -> first value, named (first-appearance order) */
SEXP _psydiff_getParameterValues(SEXP psydiffModel) {
  SEXP pt = list_get(list_get(psydiffModel, "pars"), "parameterTable");
  SEXP labels = list_get(pt, "label"), values = list_get(pt, "value");
  const int R = LENGTH(labels);
  int* first = (int*)R_alloc((size_t)(R > 0 ? R : 1), sizeof(int));
  int P = 0;
  for (int j = 0; j < R; j++) {
    int seen = 0;
    for (int k = 0; k < P && !seen; k++) seen = strcmp(CHAR(STRING_ELT(labels, first[k])), CHAR(STRING_ELT(labels, j))) == 0;
    if (!seen) first[P++] = j;
  }
  SEXP out = PROTECT(Rf_allocVector(REALSXP, P));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, P));
  for (int k = 0; k < P; k++) {
    REAL(out)[k] = REAL(values)[first[k]];
    SET_STRING_ELT(nm, k, STRING_ELT(labels, first[k]));
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

/* setParameterList(parameterTable, parameterList, person) — psydiffBasic.h:63-121: the person's CHANGED rows are written
 * into the shared parameter list at (target, row, col), by reference */
SEXP _psydiff_setParameterList(SEXP parameterTable, SEXP parameterList, SEXP person) {
  const int who = Rf_asInteger(person);
  SEXP persons = list_get(parameterTable, "person"), target = list_get(parameterTable, "target"),
       value = list_get(parameterTable, "value"), cls = list_get(parameterTable, "targetClass"),
       rw = list_get(parameterTable, "row"), cl = list_get(parameterTable, "col"), changed = list_get(parameterTable, "changed");
  SEXP names = Rf_getAttrib(parameterList, R_NamesSymbol);
  int any = 0;
  for (int i = 0; i < LENGTH(persons); i++) {
    if (num_at(persons, i) != (double)who) continue;
    any = 1;
    if (!LOGICAL(changed)[i]) continue;
    const char* tg = CHAR(STRING_ELT(target, i));
    int found = 0;
    for (int j = 0; j < LENGTH(names) && !found; j++) {
      if (strcmp(CHAR(STRING_ELT(names, j)), tg) != 0) continue;
      found = 1;
      const char* k = CHAR(STRING_ELT(cls, i));
      SEXP dst = VECTOR_ELT(parameterList, j);
      if (strcmp(k, "matrix") == 0) REAL(dst)[(int)num_at(rw, i) + (size_t)(int)num_at(cl, i) * Rf_nrows(dst)] = REAL(value)[i];
      else if (strcmp(k, "double") == 0) SET_VECTOR_ELT(parameterList, j, Rf_ScalarReal(REAL(value)[i]));
      else if (strcmp(k, "vec") == 0) REAL(dst)[(int)num_at(rw, i)] = REAL(value)[i];
      else Rf_error("Parameter class not found");
    }
    if (!found) Rf_warning("Target %s not found", tg);
  }
  if (!any) Rf_error("Person %d not found in data set.", who);
  return R_NilValue;
}

SEXP _psydiff_clonePsydiffModel(SEXP model) { return Rf_duplicate(model); }  /* src/exposeInlineFunctions.cpp:305-308 */

/* ---- the exported filter primitives, same R signatures as R/RcppExports.R:49-247 -------------------------------------- */
SEXP _psydiff_getSigmaPoints(SEXP m, SEXP A, SEXP covarianceIsRoot, SEXP c) {
  const int n = LENGTH(m);
  SEXP X = PROTECT(real_matrix(n, 2 * n + 1));
  if (psy_getSigmaPoints(n, REAL(m), REAL(A), Rf_asLogical(covarianceIsRoot), Rf_asReal(c), REAL(X)) != PSY_OK)
    Rf_error("chol(): decomposition failed");
  UNPROTECT(1);
  return X;
}
SEXP _psydiff_getMeanWeights(SEXP n, SEXP alpha, SEXP beta, SEXP kappa) {
  SEXP w = PROTECT(Rf_allocVector(REALSXP, 2 * Rf_asInteger(n) + 1));
  psy_getMeanWeights(Rf_asInteger(n), Rf_asReal(alpha), Rf_asReal(beta), Rf_asReal(kappa), REAL(w));
  UNPROTECT(1);
  return w;
}
SEXP _psydiff_getCovWeights(SEXP n, SEXP alpha, SEXP beta, SEXP kappa) {
  SEXP w = PROTECT(Rf_allocVector(REALSXP, 2 * Rf_asInteger(n) + 1));
  psy_getCovWeights(Rf_asInteger(n), Rf_asReal(alpha), Rf_asReal(beta), Rf_asReal(kappa), REAL(w));
  UNPROTECT(1);
  return w;
}
SEXP _psydiff_getWMatrix(SEXP meanWeights, SEXP covWeights) {
  const int s = LENGTH(meanWeights);
  SEXP W = PROTECT(real_matrix(s, s));
  psy_getWMatrix(s, REAL(meanWeights), REAL(covWeights), REAL(W));
  UNPROTECT(1);
  return W;
}
SEXP _psydiff_computeMeanFromSigmaPoints(SEXP sigmaPoints, SEXP meanWeights) {
  const int n = Rf_nrows(sigmaPoints), s = Rf_ncols(sigmaPoints);
  SEXP mu = PROTECT(real_matrix(n, 1));
  psy_computeMeanFromSigmaPoints(n, s, REAL(sigmaPoints), REAL(meanWeights), REAL(mu));
  UNPROTECT(1);
  return mu;
}
SEXP _psydiff_getAFromSigmaPoints(SEXP sigmaPoints, SEXP meanWeights, SEXP c) {
  const int n = Rf_nrows(sigmaPoints);
  SEXP A = PROTECT(real_matrix(n, n));
  psy_getAFromSigmaPoints(n, REAL(sigmaPoints), REAL(meanWeights), Rf_asReal(c), REAL(A));
  UNPROTECT(1);
  return A;
}
SEXP _psydiff_computePFromSigmaPoints(SEXP sigmaPoints, SEXP meanWeights, SEXP c) {
  const int n = Rf_nrows(sigmaPoints);
  SEXP P = PROTECT(real_matrix(n, n));
  psy_computePFromSigmaPoints(n, REAL(sigmaPoints), REAL(meanWeights), Rf_asReal(c), REAL(P));
  UNPROTECT(1);
  return P;
}
SEXP _psydiff_getPhi(SEXP squareMatrix) {
  const int n = Rf_nrows(squareMatrix);
  SEXP Phi = PROTECT(real_matrix(n, n));
  psy_getPhi(n, REAL(squareMatrix), REAL(Phi));
  UNPROTECT(1);
  return Phi;
}
SEXP _psydiff_predictMu(SEXP Y_, SEXP meanWeights) {
  const int m = Rf_nrows(Y_), s = Rf_ncols(Y_);
  SEXP mu = PROTECT(real_matrix(m, 1));
  psy_predictMu(m, s, REAL(Y_), REAL(meanWeights), REAL(mu));
  UNPROTECT(1);
  return mu;
}
SEXP _psydiff_predictS(SEXP Y_, SEXP W, SEXP R) {
  const int m = Rf_nrows(Y_), s = Rf_ncols(Y_);
  SEXP S = PROTECT(real_matrix(m, m));
  psy_predictS(m, s, REAL(Y_), REAL(W), REAL(R), REAL(S));
  UNPROTECT(1);
  return S;
}
SEXP _psydiff_predictC(SEXP sigmaPoints, SEXP Y_, SEXP W) {
  const int n = Rf_nrows(sigmaPoints), m = Rf_nrows(Y_), s = Rf_ncols(Y_);
  SEXP C = PROTECT(real_matrix(n, m));
  psy_predictC(n, m, s, REAL(sigmaPoints), REAL(Y_), REAL(W), REAL(C));
  UNPROTECT(1);
  return C;
}
SEXP _psydiff_computeK(SEXP C, SEXP S) {
  const int n = Rf_nrows(C), m = Rf_ncols(C);
  SEXP K = PROTECT(real_matrix(n, m));
  if (psy_computeK(n, m, REAL(C), REAL(S), REAL(K)) != PSY_OK) Rf_error("inv(): matrix is singular");
  UNPROTECT(1);
  return K;
}
SEXP _psydiff_updateM(SEXP m_, SEXP K, SEXP residual) {
  const int n = Rf_nrows(K), m = Rf_ncols(K);
  SEXP out = PROTECT(real_matrix(n, 1));
  psy_updateM(n, m, REAL(m_), REAL(K), REAL(residual), REAL(out));
  UNPROTECT(1);
  return out;
}
SEXP _psydiff_updateP(SEXP P_, SEXP K, SEXP S) {
  const int n = Rf_nrows(K), m = Rf_ncols(K);
  SEXP out = PROTECT(real_matrix(n, n));
  psy_updateP(n, m, REAL(P_), REAL(K), REAL(S), REAL(out));
  UNPROTECT(1);
  return out;
}
SEXP _psydiff_computeIndividualM2LL(SEXP nObservedVariables, SEXP rawData, SEXP expectedMeans, SEXP expectedCovariance) {
  double out;
  CHECK(psy_computeIndividualM2LL(Rf_asInteger(nObservedVariables), REAL(rawData), REAL(expectedMeans), REAL(expectedCovariance), &out));
  return Rf_ScalarReal(out);
}
SEXP _psydiff_cholupdate(SEXP L, SEXP x, SEXP v, SEXP direction) {
  SEXP out = PROTECT(Rf_duplicate(L));
  if (psy_cholupdate(Rf_nrows(L), REAL(out), REAL(x), Rf_asReal(v), CHAR(STRING_ELT(direction, 0))) != PSY_OK)
    Rf_error("Unknown direction in cholupdate. Possible are downdate and update.");   /* psydiffInternals.h:145 */
  UNPROTECT(1);
  return out;
}
SEXP _psydiff_qr_(SEXP X) {
  SEXP R = PROTECT(real_matrix(Rf_ncols(X), Rf_ncols(X)));
  psy_qr(Rf_nrows(X), Rf_ncols(X), REAL(X), REAL(R));
  UNPROTECT(1);
  return R;
}
SEXP _psydiff_logChol2Chol(SEXP logChol) {
  const int n = Rf_nrows(logChol);
  SEXP out = PROTECT(real_matrix(n, n));
  psy_logChol2Chol(n, REAL(logChol), REAL(out));
  UNPROTECT(1);
  return out;
}

/* ---- registration: the reference's table (src/RcppExports.cpp:292-316) plus the fit / gradient / compile entry points ---- */
static const R_CallMethodDef CallEntries[] = {
    {"_psydiff_setParameterValues", (DL_FUNC)&_psydiff_setParameterValues, 3},
    {"_psydiff_getParameterValues", (DL_FUNC)&_psydiff_getParameterValues, 1},
    {"_psydiff_setParameterList", (DL_FUNC)&_psydiff_setParameterList, 3},
    {"_psydiff_getSigmaPoints", (DL_FUNC)&_psydiff_getSigmaPoints, 4},
    {"_psydiff_getMeanWeights", (DL_FUNC)&_psydiff_getMeanWeights, 4},
    {"_psydiff_getCovWeights", (DL_FUNC)&_psydiff_getCovWeights, 4},
    {"_psydiff_getWMatrix", (DL_FUNC)&_psydiff_getWMatrix, 2},
    {"_psydiff_computeMeanFromSigmaPoints", (DL_FUNC)&_psydiff_computeMeanFromSigmaPoints, 2},
    {"_psydiff_getAFromSigmaPoints", (DL_FUNC)&_psydiff_getAFromSigmaPoints, 3},
    {"_psydiff_computePFromSigmaPoints", (DL_FUNC)&_psydiff_computePFromSigmaPoints, 3},
    {"_psydiff_getPhi", (DL_FUNC)&_psydiff_getPhi, 1},
    {"_psydiff_predictMu", (DL_FUNC)&_psydiff_predictMu, 2},
    {"_psydiff_predictS", (DL_FUNC)&_psydiff_predictS, 3},
    {"_psydiff_predictC", (DL_FUNC)&_psydiff_predictC, 3},
    {"_psydiff_computeK", (DL_FUNC)&_psydiff_computeK, 2},
    {"_psydiff_updateM", (DL_FUNC)&_psydiff_updateM, 3},
    {"_psydiff_updateP", (DL_FUNC)&_psydiff_updateP, 3},
    {"_psydiff_computeIndividualM2LL", (DL_FUNC)&_psydiff_computeIndividualM2LL, 4},
    {"_psydiff_cholupdate", (DL_FUNC)&_psydiff_cholupdate, 4},
    {"_psydiff_qr_", (DL_FUNC)&_psydiff_qr_, 1},
    {"_psydiff_logChol2Chol", (DL_FUNC)&_psydiff_logChol2Chol, 1},
    {"_psydiff_clonePsydiffModel", (DL_FUNC)&_psydiff_clonePsydiffModel, 1},
    {"R_psy_emit_source", (DL_FUNC)&R_psy_emit_source, 11},
    {"R_psy_compile", (DL_FUNC)&R_psy_compile, 6},
    {"R_psy_is_current", (DL_FUNC)&R_psy_is_current, 5},
    {"R_psy_upload", (DL_FUNC)&R_psy_upload, 9},
    {"R_psy_load_module", (DL_FUNC)&R_psy_load_module, 4},
    {"R_psy_set_devices", (DL_FUNC)&R_psy_set_devices, 2},
    {"R_psy_fit", (DL_FUNC)&R_psy_fit, 8},
    {"R_psy_gradients", (DL_FUNC)&R_psy_gradients, 8},
    {"R_psy_gradient", (DL_FUNC)&R_psy_gradient, 9},
    {NULL, NULL, 0}};

void R_init_psydiff(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
