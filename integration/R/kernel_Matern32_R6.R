## KernelClass_Matern32_R6 with the B200 path: same class name, fields and method signatures as R/kernel_Matern32_R6.R of the
## reference, so ace.train (R/main_ace.R:204-230) and predict.ace (R/predict.ace.R:78-96) are unchanged.
## para_update / get_train_stats / predict / predict_marginal run on a device-resident handle (external pointer):
## K, K^-1 and the optimiser moments never travel through R.  kernel_mat / kernel_mat_sym / getinv_kernel remain
## available through the per-function exports (same names as before, src/ace_b200_shim.cpp).
KernelClass_Matern32_R6 <- R6::R6Class("Matern32",
  cloneable = FALSE, class = FALSE, portable = FALSE,
  public = list(
    parameters = NULL,
    Kmat = NULL,
    Karray = NULL,
    B = NULL,
    p = NULL,
    stdy = 1,
    device = 0L,
    initialize = function(p_arg, B_arg, ext_init_parameters, std_y_arg = 1, verbose = FALSE) {
      if (verbose) cat("Using Matern 3/2 kernel\n")
      B <<- B_arg
      p <<- p_arg
      parameters <<- ext_init_parameters
      stdy <<- std_y_arg
    },
    kernel_mat = function(X1, X2, Z1, Z2) {
      Klist <- kernmat_Matern32_cpp(X1, X2, Z1, Z2, parameters)
    },
    kernel_mat_sym = function(X, Z) {
      Klist <- kernmat_Matern32_symmetric_cpp(X, Z, parameters)
      Kmat <<- Klist$full
      Karray <<- Klist$elements
      invisible(Klist)
    },
    getinv_kernel = function(X, Z) {
      kernel_mat_sym(X, Z)
      invKmatList <- invkernel_cpp(Kmat, c(parameters[1]))
      private$inv_host <- invKmatList$inv
      invisible(invKmatList)
    },
    para_update = function(iter, y, X, Z, Optim, printevery = 100, verbose = TRUE) {
      # one fused native call: kernel build, Cholesky + inverse, evidence + all gradients, clip, optimiser step and
      # mu refresh in the order of the reference's method body (R/kernel_Matern32_R6.R:39-60), quirks included
      if (is.null(private$fit)) private$fit <- private$make_fit(y, X, Z, Optim)
      out <- ace_fit_para_update_R(private$fit, iter)      # stop()s with the reference's message on non-finite gradients
      parameters <<- ace_fit_get_parameters_R(private$fit)
      stats <- out[1:2]
      if ((iter %% printevery == 0) && verbose) {
        cat(sprintf("%5d | log Evidence %9.4f | RMSE %9.4f | Norm. noise var: %3.4f | Gradient L2: %3.4f\n",
                    iter, stats[2], stats[1], exp(parameters[1]), out[3]))
      }
      invisible(stats)
    },
    get_train_stats = function(y, X, Z, invKmatList) {
      if (!is.null(private$fit) && missing(invKmatList)) {
        # local factorisation at the current parameters; the stored inverse is left untouched (R/kernel_Matern32_R6.R:61-72)
        return(ace_fit_get_train_stats_R(private$fit))
      }
      if (missing(invKmatList)) {
        Klist <- kernel_mat_sym(X, Z)
        invKmatList <- invkernel_cpp(Klist$full, c(parameters[1]))
      }
      stats <- stats_cpp(y, Kmat, invKmatList$inv, invKmatList$eigenval, c(parameters[2]), stdy)
    },
    predict = function(y, X, Z, X2, Z2, mean_y, std_y) {
      if (!is.null(private$fit)) return(ace_fit_predict_R(private$fit, X2, Z2, mean_y, std_y))
      K_xX <- kernmat_Matern32_cpp(X2, X, Z2, Z, parameters)$full
      K_xx <- kernmat_Matern32_symmetric_cpp(X2, Z2, parameters)$full
      outlist <- pred_cpp(y, parameters[1], parameters[2], invKmatn, K_xX, K_xx, mean_y, std_y)
    },
    predict_marginal = function(y, X, Z, X2, Z2, dZ2, mean_y, std_y, std_Z, calculate_ate) {
      if (!is.null(private$fit))
        return(ace_fit_predict_marginal_R(private$fit, X2, Z2, dZ2, mean_y, std_y, std_Z, calculate_ate))
      Kmarginal_xX <- kernmat_Matern32_cpp(X2, X, dZ2, Z, parameters)$elements
      Kmarginal_xx <- kernmat_Matern32_symmetric_cpp(X2, dZ2, parameters)$elements
      outlist <- pred_marginal_cpp(y, Z2, parameters[1], parameters[2], invKmatn, Kmarginal_xX, Kmarginal_xx,
                                   mean_y, std_y, std_Z, calculate_ate)
    }),
  active = list(
    # the reference's public field: downloaded from the device when somebody looks at it
    invKmatn = function(value) {
      if (!missing(value)) { private$inv_host <- value; return(invisible(value)) }
      if (!is.null(private$fit)) ace_fit_get_invKmatn_R(private$fit) else private$inv_host
    }),
  private = list(
    fit = NULL,
    inv_host = NULL,
    kernel_code = 1L,     # ACE_KERNEL_MATERN32
    make_fit = function(y, X, Z, Optim) {
      # the optimiser classes are class = FALSE environments: they tell which update they run through the `code`
      # field added in integration/R/optimizer_classes_patch.R (0 Nadam, 1 Adam, 2 GD / NAG)
      code <- if (is.null(Optim$code)) 0L else Optim$code
      ace_fit_create_R(y, X, Z, parameters, private$kernel_code, code, Optim$lr,
                       if (is.null(Optim$beta1)) 0.9 else Optim$beta1,
                       if (is.null(Optim$beta2)) 0.999 else Optim$beta2,
                       if (is.null(Optim$momentum)) 0 else Optim$momentum,
                       Optim$norm.clip, Optim$clip.at, stdy, device)
    },
    mean_solution = function(y) {
      parameters[2] <<- mu_solution_cpp(y, invKmatn)
    })
)
