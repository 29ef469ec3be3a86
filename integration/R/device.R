# device.R - the R side of the drop-in (goes to R/device.R of the reference package fmcmc).
#
# What changes in the reference's R code: ONE branch in MCMC.default (R/mcmc.R:426) that hands device families to
# MCMC_device() below, three family constructors (the replacement of the user closure `fun`), and an attribute on the
# kernel constructors.  Everything else - S3 methods, append_chains, MCMC_OUTPUT, get_*(), messages, coda objects - stays
# as it is and keeps working on the buffers src/shim.c fills.
#
# R is not installed in the build image: this file is written against src/shim.c (which IS compiled and driven in the
# test-suite, integration/test/drive_shim.c building exactly the lists built here) and against the Python mirror
# fmcmc_b200/api.py, which follows it line by line and is what tests/ and bench.py run.

# ---- device families: the replacement of the user closure `fun` ------------------------------------------------------
ll_gaussian_lm <- function(X, y, intercept = TRUE, guard = TRUE)           # README.md:128-139 / 356-360
  structure(list(family = 1L, flags = intercept * 1L + guard * 2L, X = as.matrix(X) + 0, y = as.double(y),
                 group = NULL, n_groups = 0L, hyper = numeric(4),
                 k = ncol(as.matrix(X)) + intercept + 1L, env = new.env()), class = "fmcmc_device_family")
ll_logistic <- function(X, y, prior_sd = 2)                                 # vignettes/workflow-with-fmcmc.Rmd:35-41
  structure(list(family = 2L, flags = 0L, X = as.matrix(X) + 0, y = as.double(y), group = NULL, n_groups = 0L,
                 hyper = c(prior_sd, 0, 0, 0), k = ncol(as.matrix(X)), env = new.env()), class = "fmcmc_device_family")
ll_hier_normal <- function(y, group, gamma_bounds = c(-1, 1), estimate_scales = FALSE) {  # playground/hierarchical-bayes.Rmd:45-51
  g <- as.integer(factor(group)) - 1L
  structure(list(family = 3L, flags = estimate_scales * 4L, X = NULL, y = as.double(y), group = g,
                 n_groups = max(g) + 1L, hyper = c(gamma_bounds, 0, 0),
                 k = max(g) + 2L + 2L * estimate_scales, env = new.env()), class = "fmcmc_device_family")
}
# X / y are uploaded (and packed for the tensor-core path) once per family object, like the data a closure captures;
# the external pointer's finalizer frees the HBM copy when the family is collected
device_model <- function(fun, device = getOption("fmcmc.b200.device", 0L)) {
  key <- paste0("dev", device)
  if (is.null(fun$env[[key]]))
    fun$env[[key]] <- .Call(C_fmcmc_model_create, fun$family, fun$flags, fun$X, fun$y, fun$group, fun$n_groups,
                            fun$hyper, as.integer(device))
  fun$env[[key]]
}

# ---- MCMC.default (R/mcmc.R:426): one added branch, before MCMC_init --------------------------------------------------
#   if (inherits(fun, "fmcmc_device_family"))
#     return(MCMC_device(initial, fun, nsteps, seed, nchains, burnin, thin, kernel, conv_checker, ...))
#   else if (getOption("fmcmc.b200.only", FALSE))
#     stop("`fun` must be a device family (ll_gaussian_lm, ll_logistic, ll_hier_normal): a closure cannot run on the ",
#          "GPU and there is no CPU fallback.", call. = FALSE)

`%||%` <- function(a, b) if (is.null(a)) b else a

kernel_to_spec <- function(kernel, k) {             # reads the hyper-parameters the constructors stored in the env
  cls <- attr(kernel, "b200_type")                  # set by the (unchanged-signature) kernel_*() constructors, 1..8
  if (is.null(cls))
    stop("this kernel runs R closures (kernel_new) and cannot run on the device", call. = FALSE)
  rec <- function(x, d) if (is.null(x)) rep(d, k) else check_dimensions(x, k)            # R/kernel.R:1-15
  list(type = cls, k = k,
       scheme = if (is.numeric(kernel$scheme)) 3L else match(kernel$scheme %||% "joint", c("joint", "ordered", "random")) - 1L,
       order = if (is.numeric(kernel$scheme)) as.integer(kernel$scheme),
       mu = rec(kernel$mu, 0), scale = rec(kernel$scale, 1), min = rec(kernel$min., -1), max = rec(kernel$max., 1),
       lb = process_bounds(rec(kernel$lb, -.Machine$double.xmax), TRUE),                   # R/kernel.R:25-41
       ub = process_bounds(rec(kernel$ub,  .Machine$double.xmax), FALSE),
       fixed = as.raw(rec(kernel$fixed, FALSE)), warmup = kernel$warmup %||% 0, freq = kernel$freq %||% 1,
       bw = kernel$bw %||% 0, until = kernel$until %||% Inf, eps = kernel$eps %||% 1e-4, Sd = kernel$Sd %||% -1,
       arate = kernel$arate %||% .234, nadapt = as.double(kernel$nadapt), constr = kernel$constr,
       mvn_method = 0L)
}

# per-chain state <-> the flat arrays of the ABI (include/fmcmc_b200.h: fmcmc_kernel_state): what rep_kernel copies
# (R/kernel.R:348-377) and R/mcmc.R:629-631 writes back
kernel_state_get <- function(kernel, nchains, spec) {
  kf <- sum(spec$fixed == as.raw(0)); k <- spec$k
  dlen <- switch(as.character(spec$type), "5" = kf * kf + kf, "6" = kf * kf, "7" = 3 * k, "8" = 3 * k, 0)
  ist <- matrix(0, 4L, nchains); dst <- matrix(0, max(dlen, 1), nchains)
  envs <- if (inherits(kernel, "fmcmc_kernel_list")) kernel else rep(list(kernel), nchains)
  for (c in seq_len(nchains)) {
    e <- envs[[c]]
    ist[1, c] <- e$abs_iter %||% 0; ist[3, c] <- e$nerrors %||% 0
    if (spec$type %in% 5:6 && !is.null(e$Sigma)) {
      dst[seq_len(kf * kf), c] <- as.vector(e$Sigma); ist[2, c] <- 2                      # FMCMC_STATE_INIT
      if (spec$type == 5 && !is.null(e$Mean_t_prev)) {
        dst[kf * kf + seq_len(kf), c] <- e$Mean_t_prev; ist[2, c] <- 3                    # + FMCMC_STATE_HAS_MEAN
      }
    } else if (spec$type %in% 7:8 && (e$abs_iter %||% 0) > 0) {
      obs <- e$obs_arate
      dst[, c] <- c(rep_len(e$mu, k), rep_len(e$scale, k), if (is.null(obs)) rep(0, k) else rep_len(obs, k))
      ist[2, c] <- 2 + 4 * (if (is.null(obs)) 0 else if (length(obs) == 1) 1 else 2)      # obs_arate kind, quirk D6
    }
  }
  list(istate = ist, dstate = dst)
}
kernel_state_set <- function(kernel, state, nchains, spec) {
  kf <- sum(spec$fixed == as.raw(0)); k <- spec$k
  envs <- if (nchains > 1L) rep_kernel(kernel, nchains) else list(kernel)                  # R/mcmc.R:526-527
  for (c in seq_len(nchains)) {
    e <- envs[[c]]; d <- state$dstate[, c]
    e$abs_iter <- state$istate[1, c]; e$nerrors <- state$istate[3, c]
    if (spec$type %in% 5:6) e$Sigma <- matrix(d[seq_len(kf * kf)], kf, kf)
    if (spec$type == 5 && bitwAnd(state$istate[2, c], 1L)) e$Mean_t_prev <- d[kf * kf + seq_len(kf)]
    if (spec$type %in% 7:8) {
      e$mu <- d[seq_len(k)]; e$scale <- d[k + seq_len(k)]
      kind <- bitwAnd(state$istate[2, c] %/% 4, 3L)
      e$obs_arate <- if (kind == 0) NULL else if (kind == 1) d[2 * k + 1] else d[2 * k + seq_len(k)]
    }
  }
  if (nchains > 1L) update_kernel(kernel, envs)                                            # R/kernel.R:400-422
  invisible(kernel)
}

plan_bulks <- function(nsteps, burnin, conv_checker) {                                     # R/mcmc.R:862-889, same arithmetic
  if (is.null(conv_checker)) return(nsteps)
  freq <- attr(conv_checker, "freq")
  if (freq * 2 > nsteps) freq <- 0
  bulks <- if (freq > 0) {
    b <- rep(freq, (nsteps - burnin) %/% freq)
    if ((nsteps - burnin) %% freq) c(b, (nsteps - burnin) - sum(b)) else b
  } else nsteps
  bulks[1] <- bulks[1] + burnin
  bulks
}

MCMC_device <- function(initial, fun, nsteps, seed, nchains, burnin, thin, kernel, conv_checker, ...) {
  if (...length()) stop("device families carry their data; `...` must be empty", call. = FALSE)
  initial <- check_initial(initial, nchains)                                              # R/checks.R:22-58
  if (is.null(seed)) seed <- sample.int(.Machine$integer.max, 1L)
  model <- device_model(fun)
  spec  <- kernel_to_spec(kernel, ncol(initial))
  free  <- spec$fixed == as.raw(0)
  state <- kernel_state_get(kernel, nchains, spec)          # abs_iter, Sigma, Mean_t_prev, mu, scale, obs_arate, nerrors
  bulks <- plan_bulks(nsteps, burnin, conv_checker)
  device_gelman <- inherits(conv_checker, "fmcmc_gelman_device")
  if (device_gelman)
    .Call(C_fmcmc_store_reset, model, nchains, ncol(initial), sum((bulks - c(burnin, rep(0, length(bulks) - 1))) %/% thin))
  ans <- NULL; start1 <- NA
  for (b in seq_along(bulks)) {                             # R/mcmc.R:901
    # R/mcmc.R:909-911: restart from the last KEPT row; when thin divides the previous bulk that is the row still resident
    # on the device (initial = NULL), otherwise it is passed explicitly
    init_b <- if (b == 1) t(initial)
              else if ((bulks[b - 1] - (if (b == 2) burnin else 0)) %% thin == 0) NULL
              else t(do.call(rbind, lapply(ans, function(x) x[nrow(x), ])))
    out <- .Call(C_fmcmc_run, model,
                 list(nsteps = bulks[b], burnin = if (b == 1) burnin else 0, thin = thin, nchains = nchains,
                      flags = if (device_gelman) 4L else 0L, chain_offset = 0, nchains_total = nchains, initial = init_b),
                 spec, state, list(mode = 0L, seed = seed, run_index = b - 1, kdraw = 0L, logu = NULL, z = NULL))
    state <- out[c("istate", "dstate")]
    if (b == 1) start1 <- out$report[2]
    tmp <- coda::as.mcmc.list(lapply(seq_len(nchains), function(c)
      coda::mcmc(structure(matrix(out$ans[, , c], ncol = ncol(initial)), dimnames = list(NULL, colnames(initial))),
                 start = out$report[2], end = out$report[3], thin = thin)))              # R/mcmc.R:829-836
    ans <- if (is.null(ans)) tmp else append_chains(ans, tmp)                             # R/mcmc.R:947
    for (c in seq_len(nchains)) {                                                         # R/mcmc.R:979-987
      MCMC_OUTPUT$append_("logpost", out$logpost[, c], c)
      MCMC_OUTPUT$append_("draws", matrix(out$draws[, , c], ncol = ncol(initial)), c)
    }
    if (is.null(conv_checker)) break
    R_CheckUserInterrupt_()
    if (device_gelman) {                                                                  # R/convergence.R:198-234
      g <- .Call(C_fmcmc_gelman, model, as.raw(free), start1, thin)
      val <- if (length(g[[1]]) > 1) g[[2]] else g[[1]][1]
      if (is.na(val)) {                                                                   # chol(W) failed, R/convergence.R:207-217
        warning("At ", sum(bulks[1:b]), " the Gelman diagnostic failed; trying with the next bulk.", call. = FALSE)
        next
      }
      convergence_msg_set(sprintf("Gelman-Rubin's R: %.4f.", val))
      if (val < attr(conv_checker, "threshold")) break
    } else if (conv_checker(ans[, free, drop = FALSE])) break                             # any R checker still works
  }
  kernel_state_set(kernel, state, nchains, spec)            # write-back like R/mcmc.R:629-631 (fmcmc_kernel_list)
  if (nchains == 1L) ans[[1]] else ans
}

# exported helpers on the device (optional: the R versions keep working)
cov_recursive_device <- function(X_t, Cov_t, Mean_t_prev, t., eps = 0, Sd = 1, Ik = NULL)  # R/recursive.R:63-120
  .Call(C_fmcmc_cov_recursive, t(X_t), as.double(Mean_t_prev), Cov_t + 0, as.double(t.), eps, Sd, Ik,
        getOption("fmcmc.b200.device", 0L))
reflect_on_boundaries_device <- function(x, lb, ub, which)                                  # R/kernel.R:450-493
  .Call(C_fmcmc_reflect, as.double(x), as.double(lb), as.double(ub),
        as.raw(seq_along(x) %in% which), getOption("fmcmc.b200.device", 0L))
