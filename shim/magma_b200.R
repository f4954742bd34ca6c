## magma_b200.R -- R side of the drop-in: the accelerated bodies of MagmaClustR's internal functions.
##
## Each function keeps the reference's signature and return value (file:line of the body it replaces is cited) and
## delegates the arithmetic to ONE `.Call` routine of shim/magma_b200_shim.c.  Reference keys, their order and the index
## mapping are computed HERE with the package's own code (signif / tidyr::unite / dplyr::arrange / match), so the union-grid
## mapping is the reference's by construction; only numbers and 0-based integers cross the boundary.
## Custom kernel FUNCTIONS (is.function(kern)) fall back to the original R bodies, kept as `.<name>_R`, except
## `convolution_kernel`, which maps to the library's "CONV<O>" kernel type.
## Not runnable in the build image (no R): the C half is compiled and exercised against a stand-in runtime
## (tests/test_shim.py); this half is the transcription a maintainer drops into R/.

.mb_kern <- function(kern, data = NULL) {
  if (is.character(kern)) return(kern)
  if (identical(kern, convolution_kernel)) return(paste0("CONV", dplyr::n_distinct(data$Output_ID)))
  NULL                                                   # any other function: R path
}
.mb_hp_matrix <- function(hp, ids = NULL) {              # tibble -> P x M (column-major == M x P row-major)
  if (!is.null(ids)) hp <- hp[match(ids, hp$ID), ]
  t(as.matrix(hp[, setdiff(names(hp), c("ID", "prop_mixture")), drop = FALSE]))
}
.mb_input_cols <- function(db) setdiff(names(db), c("ID", "Output", "Reference"))
## library switches (include/magma_b200.h: mb_set_option): magma_b200_option("logdet_chol", 0) evaluates log-determinants like
## R/likelihoods.R:37-41 (LU of the explicit inverse); magma_b200_option("pair_tables", 0) evaluates the SE kernel element by element
magma_b200_option <- function(name, value) invisible(.Call("mbR_set_option", as.character(name), as.double(value)))

## what e_step / m_step / ve_step / vm_step need from `db`: uploaded once per (db, grid)
.mb_upload <- function(db, grid_inputs = NULL) {
  cols <- .mb_input_cols(db)
  all_inputs <- db %>% dplyr::select(-"ID", -"Output") %>% unique()
  if (!is.null(grid_inputs)) all_inputs <- dplyr::union(all_inputs, grid_inputs)
  all_inputs <- all_inputs %>% dplyr::arrange(.data$Reference)          # em-magma.R:27-30
  ids  <- unique(db$ID)                                                  # group order of training.R:699-701
  rows <- split(seq_len(nrow(db)), factor(db$ID, levels = ids))
  ord  <- unlist(rows, use.names = FALSE)
  offs <- c(0, cumsum(lengths(rows)))
  .Call("mbR_upload_dataset", as.double(offs), t(as.matrix(db[ord, cols])), as.double(db$Output[ord]),
        match(db$Reference[ord], all_inputs$Reference) - 1L, t(as.matrix(all_inputs[, cols])), length(cols))
  list(all_inputs = all_inputs, ids = ids, U = nrow(all_inputs), ord = ord, cols = cols)
}

## ---- R/kernel-to-matrices.R ------------------------------------------------------------------------------------
kern_to_cov <- function(input, kern = "SE", hp, deriv = NULL, input_2 = NULL) {            # :71-503
  k <- .mb_kern(kern, input)
  if (is.null(k)) return(.kern_to_cov_R(input, kern, hp, deriv, input_2))
  x1 <- as.matrix(if (is.data.frame(input)) input[, setdiff(names(input), "Reference")] else input)
  x2 <- if (is.null(input_2)) NULL else
    as.matrix(if (is.data.frame(input_2)) input_2[, setdiff(names(input_2), "Reference")] else input_2)
  hpv <- unlist(hp[setdiff(names(hp), "ID")])
  d <- if (is.null(deriv)) -1L else match(deriv, names(hpv)) - 1L
  mat <- .Call("mbR_kern_to_cov", k, "noise" %in% names(hpv), as.double(hpv), x1, x2, d)
  ref <- function(m, df) if (is.data.frame(df) && "Reference" %in% names(df)) as.character(df$Reference) else
    apply(m, 1, paste, collapse = ":")
  dimnames(mat) <- list(ref(x1, input), if (is.null(x2)) ref(x1, input) else ref(x2, input_2))
  mat
}
chol_inv_jitter <- function(mat, pen_diag) {                                               # :670-680
  inv <- .Call("mbR_chol_inv_jitter", mat, pen_diag)
  dimnames(inv) <- dimnames(mat)
  inv
}
.list_kern_to_mat <- function(db, kern, hp, deriv, inverse, pen_diag) {                    # :586-652
  k <- .mb_kern(kern, db)
  ids <- unique(db$ID)
  rows <- split(seq_len(nrow(db)), factor(db$ID, levels = ids))
  ord <- unlist(rows, use.names = FALSE)
  cols <- .mb_input_cols(db)
  hpm <- .mb_hp_matrix(hp, ids)
  d <- if (is.null(deriv)) -1L else match(deriv, rownames(hpm)) - 1L
  mats <- .Call("mbR_list_kern_to_mat", k, "noise" %in% rownames(hpm), as.double(c(0, cumsum(lengths(rows)))),
                t(as.matrix(db[ord, cols])), length(cols), hpm, FALSE, d, inverse, pen_diag)
  for (t in seq_along(ids)) dimnames(mats[[t]]) <- rep(list(as.character(db$Reference[rows[[t]]])), 2)
  stats::setNames(mats, ids)
}
list_kern_to_cov <- function(data, kern, hp, deriv = NULL) .list_kern_to_mat(data, kern, hp, deriv, FALSE, 0)
list_kern_to_inv <- function(db, kern, hp, pen_diag, deriv = NULL) .list_kern_to_mat(db, kern, hp, deriv, TRUE, pen_diag)

## ---- R/em-magma.R ------------------------------------------------------------------------------------------------
e_step <- function(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag) {                        # :25-82
  k0 <- .mb_kern(kern_0, db); ki <- .mb_kern(kern_i, db)
  if (is.null(k0) || is.null(ki)) return(.e_step_R(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag))
  up  <- .mb_upload(db)
  res <- .Call("mbR_e_step", k0, as.double(.mb_hp_matrix(hp_0)), ki, .mb_hp_matrix(hp_i, up$ids), FALSE,
               as.double(m_0), pen_diag)
  cov <- res$cov
  dimnames(cov) <- list(up$all_inputs$Reference, up$all_inputs$Reference)
  list("mean" = tibble::tibble(up$all_inputs, "Output" = res$mean), "cov" = cov)
}
m_step <- function(db, m_0, kern_0, kern_i, old_hp_0, old_hp_i, post_mean, post_cov, shared_hp, pen_diag) {   # :119-235
  k0 <- .mb_kern(kern_0, db); ki <- .mb_kern(kern_i, db)
  if (is.null(k0) || is.null(ki))
    return(.m_step_R(db, m_0, kern_0, kern_i, old_hp_0, old_hp_i, post_mean, post_cov, shared_hp, pen_diag))
  up  <- .mb_upload(db)
  res <- .Call("mbR_m_step", k0, as.double(.mb_hp_matrix(old_hp_0)), ki, .mb_hp_matrix(old_hp_i, up$ids), shared_hp,
               as.double(m_0), as.double(post_mean$Output), post_cov, pen_diag)
  n0 <- setdiff(names(old_hp_0), "ID"); ni <- setdiff(names(old_hp_i), "ID")
  hp_0 <- tibble::as_tibble_row(stats::setNames(res$hp_0, n0))
  hp_i <- tibble::tibble(ID = up$ids, tibble::as_tibble(stats::setNames(as.data.frame(t(matrix(res$hp_i, length(ni)))), ni)))
  list("hp_0" = hp_0, "hp_i" = hp_i)
}
logL_monitoring <- function(hp_0, hp_i, db, m_0, kern_0, kern_i, post_mean, post_cov, pen_diag) {   # likelihoods.R:253-312
  k0 <- .mb_kern(kern_0, db); ki <- .mb_kern(kern_i, db)
  if (is.null(k0) || is.null(ki))
    return(.logL_monitoring_R(hp_0, hp_i, db, m_0, kern_0, kern_i, post_mean, post_cov, pen_diag))
  up <- .mb_upload(db)
  .Call("mbR_logL_monitoring", k0, as.double(.mb_hp_matrix(hp_0)), ki, .mb_hp_matrix(hp_i, up$ids), FALSE,
        as.double(m_0), as.double(post_mean$Output), post_cov, pen_diag)
}

## ---- R/vem-magmaclust.R, R/elbos.R ---------------------------------------------------------------------------------
.mb_clusters <- function(lst, U) vapply(lst, function(m) as.double(if (is.data.frame(m)) m$Output else m), numeric(U))
ve_step <- function(db, m_k, kern_k, kern_i, hp_k, hp_i, old_mixture, iter, pen_diag) {    # vem-magmaclust.R:32-150
  kk <- .mb_kern(kern_k, db); ki <- .mb_kern(kern_i, db)
  if (is.null(kk) || is.null(ki)) return(.ve_step_R(db, m_k, kern_k, kern_i, hp_k, hp_i, old_mixture, iter, pen_diag))
  up  <- .mb_upload(db)
  IDk <- names(m_k)
  tau <- t(as.matrix(old_mixture[match(up$ids, old_mixture$ID), IDk]))
  res <- .Call("mbR_ve_step", kk, .mb_hp_matrix(hp_k), ki, .mb_hp_matrix(hp_i, up$ids), FALSE, .mb_clusters(m_k, up$U),
               tau, as.double(hp_k$prop_mixture), iter > 1, pen_diag)
  cov <- array(res$cov, c(up$U, up$U, length(IDk)))
  list(
    "mean" = stats::setNames(lapply(seq_along(IDk), function(k) tibble::tibble(up$all_inputs, "Output" = res$mean[, k])), IDk),
    "cov" = stats::setNames(lapply(seq_along(IDk), function(k)
      `dimnames<-`(cov[, , k], rep(list(up$all_inputs$Reference), 2))), IDk),
    "mixture" = tibble::tibble(ID = up$ids, tibble::as_tibble(stats::setNames(as.data.frame(t(res$mixture)), IDk))))
}
vm_step <- function(db, old_hp_k, old_hp_i, list_mu_param, kern_k, kern_i, m_k, shared_hp_k, shared_hp_i, pen_diag) {  # :191-303
  kk <- .mb_kern(kern_k, db); ki <- .mb_kern(kern_i, db)
  if (is.null(kk) || is.null(ki))
    return(.vm_step_R(db, old_hp_k, old_hp_i, list_mu_param, kern_k, kern_i, m_k, shared_hp_k, shared_hp_i, pen_diag))
  up  <- .mb_upload(db)
  IDk <- names(m_k)
  mix <- list_mu_param$mixture
  res <- .Call("mbR_vm_step", kk, .mb_hp_matrix(old_hp_k), ki, .mb_hp_matrix(old_hp_i, up$ids), shared_hp_k, shared_hp_i,
               .mb_clusters(list_mu_param$mean, up$U), simplify2array(list_mu_param$cov),
               t(as.matrix(mix[match(up$ids, mix$ID), IDk])), pen_diag)
  nk <- setdiff(names(old_hp_k), c("ID", "prop_mixture")); ni <- setdiff(names(old_hp_i), "ID")
  hp_k <- tibble::tibble(ID = IDk, tibble::as_tibble(stats::setNames(as.data.frame(t(matrix(res$hp_k, length(nk)))), nk)),
                         prop_mixture = res$prop_mixture)
  hp_i <- tibble::tibble(ID = up$ids, tibble::as_tibble(stats::setNames(as.data.frame(t(matrix(res$hp_i, length(ni)))), ni)))
  list("m_k" = stats::setNames(lapply(seq_along(IDk), function(k) res$m_k[, k]), IDk), "hp_k" = hp_k, "hp_i" = hp_i)
}
update_mixture <- function(db, mean_k, cov_k, hp, kern, prop_mixture, pen_diag) {          # :330-396
  ki <- .mb_kern(kern, db)
  if (is.null(ki)) return(.update_mixture_R(db, mean_k, cov_k, hp, kern, prop_mixture, pen_diag))
  up  <- .mb_upload(db)
  IDk <- names(mean_k)
  tau <- .Call("mbR_update_mixture", ki, .mb_hp_matrix(hp, up$ids), FALSE, .mb_clusters(mean_k, up$U),
               simplify2array(cov_k), as.double(unlist(prop_mixture[IDk])), pen_diag, length(up$ids))
  tibble::tibble(ID = up$ids, tibble::as_tibble(stats::setNames(as.data.frame(t(tau)), IDk)))
}
elbo_monitoring_VEM <- function(hp_k, hp_i, db, kern_i, kern_k, hyperpost, m_k, pen_diag) {   # elbos.R:200-272
  kk <- .mb_kern(kern_k, db); ki <- .mb_kern(kern_i, db)
  if (is.null(kk) || is.null(ki)) return(.elbo_monitoring_VEM_R(hp_k, hp_i, db, kern_i, kern_k, hyperpost, m_k, pen_diag))
  up  <- .mb_upload(db)
  IDk <- names(m_k)
  mix <- hyperpost$mixture
  .Call("mbR_elbo_monitoring", kk, .mb_hp_matrix(hp_k), ki, .mb_hp_matrix(hp_i, up$ids), FALSE, .mb_clusters(m_k, up$U),
        .mb_clusters(hyperpost$mean, up$U), simplify2array(hyperpost$cov), t(as.matrix(mix[match(up$ids, mix$ID), IDk])),
        as.double(hp_k$prop_mixture), pen_diag)
}

## ---- the whole loops on device-resident state: the bodies of the `for (i in 1:n_iter_max)` loops of
## train_magma (R/training.R:896-1065) and train_magmaclust (R/training.R:1586-1758) become one call each
.train_magma_loop <- function(db, m_0, kern_0, kern_i, hp_0, hp_i, shared_hp, pen_diag, n_iter_max, cv_threshold) {
  up  <- .mb_upload(db)
  .Call("mbR_train_magma", .mb_kern(kern_0, db), as.double(.mb_hp_matrix(hp_0)), .mb_kern(kern_i, db),
        .mb_hp_matrix(hp_i, up$ids), shared_hp, as.double(m_0), pen_diag, as.integer(n_iter_max), cv_threshold)
}
.train_magmaclust_loop <- function(db, m_k, kern_k, kern_i, hp_k, hp_i, mixture, shared_hp_k, shared_hp_i, pen_diag,
                                   n_iter_max, cv_threshold, fast_approx) {
  up  <- .mb_upload(db)
  IDk <- names(m_k)
  .Call("mbR_train_magmaclust", .mb_kern(kern_k, db), .mb_hp_matrix(hp_k), .mb_kern(kern_i, db), .mb_hp_matrix(hp_i, up$ids),
        shared_hp_k, shared_hp_i, .mb_clusters(m_k, up$U), t(as.matrix(mixture[match(up$ids, mixture$ID), IDk])), pen_diag,
        as.integer(n_iter_max), cv_threshold, fast_approx)
}

## ---- R/prediction.R: the algebra blocks ------------------------------------------------------------------------------
## hyperposterior (:670-707): the E step re-evaluated on data U grid_inputs
.hyperposterior_math <- function(db, grid_inputs, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag) {
  up  <- .mb_upload(db, grid_inputs)
  res <- .Call("mbR_hyperposterior", .mb_kern(kern_0, db), as.double(.mb_hp_matrix(hp_0)), .mb_kern(kern_i, db),
               .mb_hp_matrix(hp_i, up$ids), FALSE, t(as.matrix(up$all_inputs[, up$cols])),
               match(db$Reference[up$ord], up$all_inputs$Reference) - 1L, as.double(m_0), pen_diag)
  cov <- res$cov
  dimnames(cov) <- rep(list(up$all_inputs$Reference), 2)
  list("mean" = tibble::tibble(up$all_inputs, "Output" = res$mean), "cov" = cov)
}
## pred_magma (:1146-1197): conditional of the new task given its observations
.pred_magma_math <- function(kern, hp, input_obs, y_obs, input_pred, mean_obs, mean_pred, cov_obs, cov_pred, cov_crossed,
                             pen_diag, get_full_cov) {
  hpv <- unlist(hp[setdiff(names(hp), "ID")])
  .Call("mbR_pred_magma", .mb_kern(kern), as.double(hpv), t(as.matrix(input_obs)), as.double(y_obs), t(as.matrix(input_pred)),
        ncol(as.matrix(input_obs)), as.double(mean_obs), as.double(mean_pred), cov_obs, cov_pred, cov_crossed, pen_diag,
        get_full_cov)
}
