# model_WRMF_b200.R -- the only R-side change needed in rsparse to run WRMF on libb200als.so.
#
# (1) Nothing: with src/b200als_shim.c linked instead of src/wrmf_implicit.cpp / src/wrmf_explicit.cpp,
#     `als_implicit_float()` etc. in R/RcppExports.R:88-102 resolve to the shim's
#     `_rsparse_als_*` routines, so WRMF$fit_transform() / $transform() run unchanged -- one
#     host<->device round trip per half-iteration (the stateless C-ABI calls).
#
# (2) Optional fast path below: keep both factor matrices in HBM for the whole fit
#     (R/model_WRMF.R:318-338 moves into b200als_fit).  Drop-in replacement for the loop body of
#     WRMF$fit_transform when precision == "float"; bias terms (with_user_item_bias / with_global_bias) stay on the
#     device too: private$rank already counts the two bias rows (R/model_WRMF.R:162-166) and b200als_R_set_bias passes the
#     flags after the reference's own initialisation code (R/model_WRMF.R:260-297) has filled the bias rows.
#
# Not executed in the authoring image (no R there); see INTEGRATION.md.

b200als_fit_transform = function(self, private, c_ui, c_iu, n_iter, convergence_tol) {
  solver_code = private$solver_code              # 0 cholesky, 1 conjugate_gradient (R/model_WRMF.R:99-100)
  feedback_code = if (private$feedback == "implicit") 0L else 1L
  session = .Call("b200als_R_create", c_ui, c_iu, private$rank, feedback_code, solver_code,
                  private$cg_steps, private$dynamic_lambda, private$lambda)
  if (private$with_user_item_bias || self$global_bias != 0)
    .Call("b200als_R_set_bias", session, private$with_user_item_bias,
          if (private$feedback == "implicit") self$global_bias else 0)
  .Call("b200als_R_set_factors", session, 1L, private$U)          # users  (rank x n_user float32)
  .Call("b200als_R_set_factors", session, 0L, self$components)    # items  (rank x n_item float32)
  trace = .Call("b200als_R_fit", session, as.integer(n_iter), as.numeric(convergence_tol))
  for (i in seq_len(length(trace) / 2)) {
    logger$info("iter %d (items) loss = %.4f", i, trace[2 * i - 1])
    logger$info("iter %d (users) loss = %.4f", i, trace[2 * i])
  }
  .Call("b200als_R_get_factors", session, 0L, self$components)    # in place, like the reference's solver
  .Call("b200als_R_get_factors", session, 1L, private$U)
  if (private$feedback == "implicit" && !private$with_user_item_bias && self$global_bias != 0)   # wrmf_implicit.hpp:111-112
    self$global_bias_base = float::fl(-self$global_bias * rowSums(float::dbl(self$components)))
  res = float::float(0, nrow = private$rank, ncol = ncol(c_iu))
  .Call("b200als_R_transform", session, res)                      # transform_ with avoid_cg (R/model_WRMF.R:412-452)
  t(res)
}
