# R side of the drop-in (RUN-UNTESTED here: no R in the build container; the C side it calls is exercised under a mock of the
# R API, tests/test_shim.py).
# ruvIII(), fastruvIII() and ruvIIIMultipleK() keep their formals and return shapes (ruvIII.R:38-56,
# fastruvIII.R:40-59, ruvIIIMultipleK.R:34-52); only the arithmetic between `Y <- RUV1(...)` and
# `newY <- Y - W %*% alpha` (ruvIII.R:116-146) is replaced by this call.

ruvIIICore <- function(assay, replicate.data, row.names.Y, ctl, k, mode = 0L, apply.log = FALSE,
                       pseudo.count = 1, return.info = FALSE) {
    # tological() of ruvIII.R:80-84: ctl may be a logical vector or a vector of indices
    ctl.logical <- rep(FALSE, nrow(assay))
    ctl.logical[ctl] <- TRUE
    res <- .Call('_RUVIIIPRPS_ruvIIICore', PACKAGE = 'RUVIIIPRPS',
                 assay,                                   # n x m1, untouched (no t(), no log2 on the R side)
                 t(replicate.data),                       # n x n_ps
                 as.integer(factor(row.names.Y)),         # replicate.matrix() is never materialised
                 ctl.logical, as.integer(k), as.integer(mode), as.logical(apply.log),
                 as.numeric(pseudo.count), as.logical(return.info))
    names(res) <- c('newY', 'fullalpha', 'W', 'r')
    if (return.info && length(res$fullalpha) > 0) {
        res$fullalpha <- t(res$fullalpha)                 # r x n as in ruvIII.R:140
        m.w <- if (mode == 1L) ncol(assay) else ncol(assay) + nrow(replicate.data)
        k.eff <- pmin(as.integer(k), res$r)
        ends <- cumsum(m.w * k.eff)
        res$W <- lapply(seq_along(k), function(i)         # each block is t(W): k x m_w column-major
            t(matrix(res$W[(ends[i] - m.w * k.eff[i] + 1):ends[i]], nrow = k.eff[i], ncol = m.w)))
    }
    res
}

# Inside ruvIII() the replaced block becomes (save.se.obj branch, ruvIII.R:152-158):
#   core <- ruvIIICore(assay(se.obj, assay.name), replicate.data,
#                      c(colnames(se.obj), row.names(replicate.data)), ctl, k,
#                      mode = 0L, apply.log = apply.log, pseudo.count = pseudo.count, return.info = return.info)
#   se.obj@assays@data[[paste0('RUV_K', k, '_on_', assay.name)]] <- core$newY[[1]]
# and ruvIIIMultipleK()'s loop (ruvIIIMultipleK.R:75-94) becomes one call with the whole k vector.
