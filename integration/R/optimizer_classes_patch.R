## Patch for R/optimizer_classes.R: one public field per optimiser class, read by Kernel$para_update to configure the
## device-resident handle (the classes are `class = FALSE`, so there is no S3 class attribute to dispatch on).
## Everything else in R/optimizer_classes.R stays as it is: `update()` keeps calling norm_clip_cpp and
## Nadam_cpp / Adam_cpp / Nesterov_cpp (now bodies in src/ace_b200_shim.cpp), which is what the unfused path
## (getinv_kernel + grad_*_cpp from R) uses.
##
##   optAdam      <- R6::R6Class("AdamOpt",     ..., public = list(code = 1L, m = NULL, v = NULL, ...
##   optNadam     <- R6::R6Class("NadamOpt",    ..., public = list(code = 0L, m = NULL, v = NULL, ...
##   optNesterov  <- R6::R6Class("NesterovOpt", ..., public = list(code = 2L, nu = NULL, beta1 = 0.9, beta2 = 0.999, ...
##
## optNesterov has no beta1 / beta2 and calls its rate `momentum`; the kernel classes pass defaults for whatever
## is missing (integration/R/kernel_SE_R6.R, make_fit).

## robust_treatment (R/robust_treatment.R:93-128) with ONE batched native call instead of n.steps + 1 predictions;
## `ace_fit_predict_marginal_batch_R` is the Rcpp face of ace_fit_predict_marginal_batch (same marshalling pattern as
## ace_fit_predict_marginal_R in src/ace_b200_shim.cpp: subsets as a raw n.pred x (n.steps + 1) matrix):
##
##   idx <- sapply(discard.steps, function(s) prediction$var <= s)
##   res <- ace_fit_predict_marginal_batch_R(fit, X.pred, B.pred, dB.pred, mean_y, std_y, std_Z, idx)
##   ate_results[, 1:4] <- res$ate ; att_results[, 1:4] <- res$att ; atu_results[, 1:4] <- res$atu
