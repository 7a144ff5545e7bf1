# psydiff_b200.R — R side of the drop-in: same function names and return values as the reference
# (compileModel R/prepareModel.R:553-562; fitModel / getGradients / getGradient R/modelTemplate.R:165-166, 853-854, 932-933;
# getGradientsParallel R/prepareModel.R:608-616).  newPsydiff is unchanged except that model$cppCode is produced by
# psydiffCudaSource() below — the library's own code generator (psy_emit_model_source, psy_codegen.cpp) — instead of
# pasting prepareEquations() into modelTemplate() (R/prepareModel.R:479-512):
#
#   modelCpp <- psydiffCudaSource(nlatent, nmanifest, latentEquations, manifestEquations, parameterList,
#                                 parameterTableInit, srUpdate, integrateFunction)

# parameterList: named list of start values (numeric scalars / matrices), parameterTableInit: one person's rows of the
# parameter table (label, value, target, targetClass, row, col; row/col 0-based as extractParameters writes them,
# R/prepareModel.R:646-690).  The k-th row of parameterTableInit is free slot k-1 of every filter instance.
psydiffCudaSource <- function(nlatent, nmanifest, latentEquations, manifestEquations, parameterList, parameterTableInit,
                              srUpdate = TRUE, integrateFunction = "default") {
  values <- lapply(parameterList, function(v) { v <- as.matrix(v); storage.mode(v) <- "double"; v })
  slots <- lapply(values, function(v) matrix(-1L, nrow(v), ncol(v)))
  for (k in seq_len(nrow(parameterTableInit))) {
    tg <- as.character(parameterTableInit$target[k])
    slots[[tg]][parameterTableInit$row[k] + 1L, parameterTableInit$col[k] + 1L] <- k - 1L
  }
  isScalar <- vapply(parameterList, function(v) is.null(dim(v)) && length(v) == 1L, logical(1))
  .Call("R_psy_emit_source", as.integer(nlatent), as.integer(nmanifest), nrow(parameterTableInit),
        paste(latentEquations, collapse = "\n"), paste(manifestEquations, collapse = "\n"),
        names(parameterList), isScalar, values, slots, isTRUE(srUpdate),
        if (identical(integrateFunction, "runge_kutta_dopri5")) 1L else 0L)
}

.psy_stepper <- function(f) switch(f, "rk4" = 0L, "runge_kutta_dopri5" = 1L, 2L)
.psy_direction <- function(d) {
  k <- match(d, c("central", "left", "right"))
  if (is.na(k)) stop("Unknown direction argument. Possible are central, left and right")
  k - 1L
}
# srUpdate is a compile-time property of the kernel.  The reference bakes it into model$cppCode and keeps no field for it
# (R/prepareModel.R:518-540); the patched newPsydiff therefore stores `srUpdate` and `cudaSpec` (the arguments of
# psydiffCudaSource) in the model list, so that the variant can be read back and the source regenerated.
.psy_sr <- function(m) if (is.null(m$srUpdate)) TRUE else isTRUE(m$srUpdate)
.psy_settings <- function(m) list(c(m$alpha, m$beta, m$kappa), .psy_stepper(m$integrateFunction), as.numeric(m$timeStep), .psy_sr(m))
.psy_variant <- function(m) c(if (identical(m$integrateFunction, "runge_kutta_dopri5")) 1L else 0L, as.integer(.psy_sr(m)))

# state of the compiled model: the device handle, what its module was generated for, where the kernel headers and cache are
.psy_env <- new.env(parent = emptyenv())

# The reference reads model$integrateFunction on every fitModel call (R/modelTemplate.R:174); here the stepper and the update
# variant are baked into the kernel, so a model whose (stepper, srUpdate) differ from the loaded module gets the matching
# module (regenerated from model$cudaSpec, compiled once, cached by content hash) before it runs.
.psy_variant_sync <- function(m) {
  v <- .psy_variant(m)
  if (!identical(v, .psy_env$variant)) {
    sp <- m$cudaSpec
    if (is.null(sp)) stop("model$cudaSpec is missing: build the model with the patched newPsydiff to change integrateFunction / srUpdate after compileModel")
    src <- psydiffCudaSource(m$nlatent, m$nmanifest, sp$latentEquations, sp$manifestEquations, sp$parameterList,
                             sp$parameterTableInit, .psy_sr(m), m$integrateFunction)
    .Call("R_psy_load_module", .psy_env$handle, src, .psy_env$include, .psy_env$cache)
    .psy_env$variant <- v
  }
  invisible(NULL)
}

# Make the handle hold THIS model's dataset and slot map (re-upload only when it is handed different R objects).
.psy_sync <- function(handle, m) {
  .psy_variant_sync(m)
  pt <- m$pars$parameterTable
  if (!.Call("R_psy_is_current", handle, m$data$observations, m$data$dt, m$data$person, pt$label)) {
    persons <- unique(m$data$person)
    labs <- unique(pt$label)                                   # theta, first-appearance order
    tidx <- as.integer(match(pt$label, labs) - 1L)             # person-major, slot-minor: exactly psy_set_slots' layout
    off <- c(0, cumsum(rle(as.vector(m$data$person))$lengths))
    .Call("R_psy_upload", handle, m$data$observations, m$data$dt, m$data$person, as.numeric(off), as.integer(persons),
          length(labs), tidx, pt$label)
  }
  invisible(NULL)
}

# compileModel(psydiffModel) is called for its side effect, like the reference's (R/prepareModel.R:553-562): it defines
# fitModel / getGradients / getGradient in the global environment.  The compiled kernel and the resident dataset live in
# `handle`, captured by those closures (the counterpart of the shared object Rcpp::sourceCpp loads); the functions accept
# any model object of the same structure, as the reference's do.
compileModel <- function(psydiffModel) {
  nfree <- nrow(psydiffModel$pars$parameterTable) / length(unique(psydiffModel$data$person))
  .psy_env$include <- system.file("include", package = "psydiff")
  .psy_env$cache <- tools::R_user_dir("psydiff", "cache")       # created recursively by psy_model_compile
  handle <- .Call("R_psy_compile", psydiffModel$cppCode, .psy_env$include, .psy_env$cache, psydiffModel$nlatent,
                  psydiffModel$nmanifest, nfree)
  .psy_env$handle <- handle
  .psy_env$variant <- .psy_variant(psydiffModel)                # model$cppCode was generated for these by newPsydiff
  assign("fitModel", function(psydiffModel, skipUpdate = FALSE) {
    .psy_sync(handle, psydiffModel)
    s <- .psy_settings(psydiffModel)
    # overwrites psydiffModel$m2LL / $latentScores / $predictedManifest and clears parameterTable$changed IN PLACE, as the
    # reference does through Rcpp (R/modelTemplate.R:410-414), and returns list(latentScores, predictedManifest, m2LL)
    .Call("R_psy_fit", handle, psydiffModel, s[[1]], s[[2]], s[[3]], s[[4]], getParameterValues(psydiffModel), skipUpdate)
  }, envir = globalenv())
  assign("getGradients", function(psydiffModel) {
    .psy_sync(handle, psydiffModel)
    s <- .psy_settings(psydiffModel)
    .Call("R_psy_gradients", handle, s[[1]], s[[2]], s[[3]], s[[4]], getParameterValues(psydiffModel),
          as.numeric(psydiffModel$eps), .psy_direction(psydiffModel$direction))
  }, envir = globalenv())
  assign("getGradient", function(psydiffModel, currentParameterValues, par) {
    .psy_sync(handle, psydiffModel)
    s <- .psy_settings(psydiffModel)
    .Call("R_psy_gradient", handle, s[[1]], s[[2]], s[[3]], s[[4]], currentParameterValues, as.integer(par),
          as.numeric(psydiffModel$eps), .psy_direction(psydiffModel$direction))
  }, envir = globalenv())
  invisible(NULL)
}

# compileParallelGradients (R/prepareModel.R:572-589): the reference starts nCores PSOCK workers, each recompiling the model,
# and ships them the whole model on every getGradientsParallel call.  Here the workers are the GPUs of the box, inside the
# library: the compiled model's persons are spread over min(nCores, visible devices) devices (one host thread, partial sums
# added over NVLink).  fitModel / getGradients / getGradient use them from then on; getGradientsParallel is getGradients.
compileParallelGradients <- function(psydiffModel, nCores) {
  if (is.null(.psy_env$handle)) compileModel(psydiffModel)
  .psy_sync(.psy_env$handle, psydiffModel)
  devices <- .Call("R_psy_set_devices", .psy_env$handle, as.integer(nCores))
  psydiffModel$cl <- list(nCores = nCores, devices = devices)
  psydiffModel
}
stopParallelGradients <- function(psydiffModel) {
  if (!is.null(.psy_env$handle)) .Call("R_psy_set_devices", .psy_env$handle, 1L)
  invisible(NULL)
}
getGradientsParallel <- function(psydiffModel) getGradients(psydiffModel)
