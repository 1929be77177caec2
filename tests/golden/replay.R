#!/usr/bin/env Rscript
# replay.R -- pins the golden fixtures of this directory against the REFERENCE ITSELF (RMolania/RUVIIIPRPS in R).
#
#   Rscript tests/golden/replay.R [golden.dir] [reference.R.dir]
#
# The build container of this repo has no R, so every parity claim of the CUDA engine is "matches the numpy restatement
# (oracle/ruviii_oracle.py)" and the fixtures here are outputs of that restatement, stored exactly as R holds a matrix
# (raw little-endian float64, column-major).  This script closes the loop in ONE command on any machine with R,
# SummarizedExperiment, ruv and BiocSingular: it reads the same bytes, runs the reference's own
#     ruvIII()           R/ruvIII.R:38-203
#     fastruvIII()       R/fastruvIII.R:40-234
#     ruvIIIMultipleK()  R/ruvIIIMultipleK.R:34-187   (save.se.obj = TRUE branch, :75-94 -- the one normalise() uses)
# and compares with the stored outputs: relative Frobenius error on newY (tolerance 1e-9, BASELINE.json north_star),
# fullalpha rows and W columns after sign alignment (singular vectors are sign-free).  Exit status 1 when any comparison
# is above tolerance, 0 otherwise.  The reference functions are taken from the installed package RUVIIIPRPS when there is
# one, otherwise source()d from `reference.R.dir` (default: $RUVIIIPRPS_SRC or /root/reference/R).
#
# A first part needs BASE R ONLY and pins the statistics either side of the path (lm residuals of regress.out.*, the one-way
# ANOVA F, Pearson / Spearman correlations, the full svd of computePCA) against stats::lm / oneway.test / cor / svd; it runs
# even when the Bioconductor packages are missing (the RUV-III part is then skipped, and the script says so).
#
# COMPILE-/RUN-UNTESTED in the build container (no R there); base R only apart from the reference's own imports.

args <- commandArgs(trailingOnly = TRUE)
this.file <- sub("^--file=", "", grep("^--file=", commandArgs(FALSE), value = TRUE))
golden.dir <- if (length(args) >= 1) args[1] else if (length(this.file)) dirname(normalizePath(this.file)) else "."
ref.dir <- if (length(args) >= 2) args[2] else Sys.getenv("RUVIIIPRPS_SRC", "/root/reference/R")
tol.newY <- 1e-9
tol.signed <- 1e-7

have.pkgs <- all(vapply(c("SummarizedExperiment", "ruv", "BiocSingular"), requireNamespace, logical(1), quietly = TRUE))
if (!have.pkgs) {
    message("SummarizedExperiment / ruv / BiocSingular not installed: only the BASE-R part (lm residuals, ANOVA F, correlations, svd) runs")
} else {
suppressPackageStartupMessages({
    library(SummarizedExperiment)
    library(ruv)
    library(BiocSingular)
})
}
if (!have.pkgs) {
    ruvIII <- fastruvIII <- ruvIIIMultipleK <- NULL
} else if (requireNamespace("RUVIIIPRPS", quietly = TRUE)) {
    ruvIII <- RUVIIIPRPS::ruvIII
    fastruvIII <- RUVIIIPRPS::fastruvIII
    ruvIIIMultipleK <- RUVIIIPRPS::ruvIIIMultipleK
    message("reference: installed package RUVIIIPRPS ", as.character(utils::packageVersion("RUVIIIPRPS")))
} else {
    suppressPackageStartupMessages(library(DelayedArray))
    for (f in c("printColoredMessage.R", "CheckSeObj.R", "ruvIII.R", "fastruvIII.R", "ruvIIIMultipleK.R"))
        source(file.path(ref.dir, f))
    message("reference: sourced from ", ref.dir)
}
if (have.pkgs) message("ruv ", as.character(utils::packageVersion("ruv")), ", BiocSingular ", as.character(utils::packageVersion("BiocSingular")))
message(R.version.string)

read.mat <- function(case, name, dims) {
    d <- dims[dims$name == name, ]
    con <- file(file.path(golden.dir, paste0(case, ".", name, ".bin")), "rb")
    on.exit(close(con))
    matrix(readBin(con, "double", n = d$nrow * d$ncol, endian = "little"), d$nrow, d$ncol)
}
rel.frob <- function(a, b) sqrt(sum((a - b)^2)) / sqrt(sum(b^2))
align.rows <- function(a, ref) {                      # flip each row of a so that <a_i, ref_i> >= 0
    s <- sign(rowSums(a * ref)); s[s == 0] <- 1
    a * s
}
failures <- 0
check <- function(what, err, tol) {
    ok <- is.finite(err) && err <= tol
    cat(sprintf("%-58s %.3e  %s\n", what, err, if (ok) "ok" else paste0("ABOVE ", tol)))
    if (!ok) failures <<- failures + 1
}

# ---- BASE R ONLY: the statistics either side of the path (SURVEY.md 8f) -------------------------------------------------
# supervisedFindNCG.R:111-121 lm(t(expr) ~ ...)$residuals; :179-191 row_oneway_equalvar(x, g)$statistic (the classical
# equal-variance one-way F: stats::oneway.test(var.equal = TRUE)); :214-229 correls(y, x, type)[, "correlation"] (stats::cor);
# computePCA.R:150-190 svd(scale(t(x), center, scale)).  matrixTests / Rfast are compared too when they happen to be installed.
for (case in c("mid")) {
    dims <- read.delim(file.path(golden.dir, paste0(case, ".manifest.tsv")), stringsAsFactors = FALSE)
    if (!("side.covariate" %in% dims$name)) next
    assay.mat <- read.mat(case, "assay", dims)                          # genes x samples
    batch <- factor(readLines(file.path(golden.dir, paste0(case, ".side_batch.txt"))))
    cov <- as.numeric(read.mat(case, "side.covariate", dims))
    y <- lm(t(assay.mat) ~ batch + cov)
    check(paste(case, "lm(t(expr) ~ batch + cov)$residuals"), rel.frob(t(unname(y$residuals)), read.mat(case, "side.lm_residuals", dims)), tol.newY)
    Fstat <- apply(assay.mat, 1, function(v) unname(oneway.test(v ~ batch, var.equal = TRUE)$statistic))
    F.ref <- as.numeric(read.mat(case, "side.anovaF", dims))
    check(paste(case, "one-way ANOVA F (oneway.test, var.equal)"), rel.frob(Fstat, F.ref), tol.newY)
    if (requireNamespace("matrixTests", quietly = TRUE))
        check(paste(case, "matrixTests::row_oneway_equalvar statistic"),
              rel.frob(matrixTests::row_oneway_equalvar(assay.mat, batch)$statistic, F.ref), tol.newY)
    r.ref <- as.numeric(read.mat(case, "side.pearson", dims))
    rho.ref <- as.numeric(read.mat(case, "side.spearman", dims))
    check(paste(case, "Pearson r (stats::cor)"), max(abs(as.numeric(cor(t(assay.mat), cov)) - r.ref)), 1e-12)
    check(paste(case, "Spearman rho (stats::cor, ties averaged)"), max(abs(as.numeric(cor(t(assay.mat), cov, method = "spearman")) - rho.ref)), 1e-12)
    if (requireNamespace("Rfast", quietly = TRUE)) {
        check(paste(case, "Rfast::correls pearson"), max(abs(Rfast::correls(cov, t(assay.mat), type = "pearson")[, "correlation"] - r.ref)), 1e-10)
        check(paste(case, "Rfast::correls spearman"), max(abs(Rfast::correls(cov, t(assay.mat), type = "spearman")[, "correlation"] - rho.ref)), 1e-10)
    }
    sv <- svd(scale(t(assay.mat), center = TRUE, scale = FALSE))
    d.ref <- as.numeric(read.mat(case, "side.svd_d", dims))
    nz <- seq_len(length(d.ref) - 1)                                    # the centred matrix has rank m - 1: the last value is noise
    check(paste(case, "svd(scale(t(x)))$d"), max(abs(sv$d[nz] - d.ref[nz]) / d.ref[1]), 1e-12)
    check(paste(case, "svd rank-10 reconstruction"),
          rel.frob(sv$u[, 1:10] %*% (sv$d[1:10] * t(sv$v[, 1:10])), read.mat(case, "side.svd_rank10", dims)), tol.newY)
}

for (case in if (have.pkgs) c("tiny", "mid") else character(0)) {
    dims <- read.delim(file.path(golden.dir, paste0(case, ".manifest.tsv")), stringsAsFactors = FALSE)
    k <- dims$nrow[dims$name == "k"]
    ks.multi <- dims$nrow[dims$name == "k_multi"]
    sample.names <- readLines(file.path(golden.dir, paste0(case, ".sample_names.txt")))
    replicate.names <- readLines(file.path(golden.dir, paste0(case, ".replicate_names.txt")))
    ctl <- scan(file.path(golden.dir, paste0(case, ".ctl_index_1based.txt")), what = integer(), quiet = TRUE)
    assay.mat <- read.mat(case, "assay", dims)                          # genes x samples
    gene.names <- paste0("g", seq_len(nrow(assay.mat)))
    dimnames(assay.mat) <- list(gene.names, sample.names)
    replicate.data <- read.mat(case, "replicate_data", dims)            # pseudo-samples x genes (normalise.R:143)
    dimnames(replicate.data) <- list(replicate.names, gene.names)
    se <- SummarizedExperiment(assays = list(HTseq = assay.mat))

    # ---- ruvIII, list return (ruvIII.R:193-197) ----
    out <- ruvIII(se.obj = se, assay.name = "HTseq", apply.log = FALSE, replicate.data = replicate.data, ctl = ctl, k = k,
                  return.info = TRUE, save.se.obj = FALSE, assess.se.obj = FALSE, verbose = FALSE)
    check(paste(case, "ruvIII newY"), rel.frob(unname(out$newY), read.mat(case, "ruvIII.newY", dims)), tol.newY)
    fa.ref <- read.mat(case, "ruvIII.fullalpha", dims)
    check(paste(case, "ruvIII nrow(fullalpha)"), abs(nrow(out$fullalpha) - nrow(fa.ref)), 0)
    kk <- min(k, nrow(fa.ref))
    check(paste(case, "ruvIII fullalpha[1:k,] (sign aligned)"),
          rel.frob(align.rows(unname(out$fullalpha[1:kk, , drop = FALSE]), fa.ref[1:kk, , drop = FALSE]), fa.ref[1:kk, , drop = FALSE]), tol.signed)
    W.ref <- read.mat(case, "ruvIII.W", dims)
    check(paste(case, "ruvIII W (sign aligned)"), rel.frob(t(align.rows(t(unname(out$W)), t(W.ref))), W.ref), tol.signed)
    check(paste(case, "ruvIII W %*% alpha (sign free)"),
          rel.frob(unname(out$W %*% out$fullalpha[1:kk, , drop = FALSE]), W.ref %*% fa.ref[1:kk, , drop = FALSE]), 10 * tol.newY)

    # ---- fastruvIII ----
    f <- fastruvIII(se.obj = se, assay.name = "HTseq", apply.log = FALSE, replicate.data = replicate.data, ctl = ctl, k = k,
                    return.info = TRUE, save.se.obj = FALSE, assess.se.obj = FALSE, verbose = FALSE)
    check(paste(case, "fastruvIII newY"), rel.frob(unname(f$newY), read.mat(case, "fastruvIII.newY", dims)), tol.newY)
    ffa.ref <- read.mat(case, "fastruvIII.fullalpha", dims)
    check(paste(case, "fastruvIII fullalpha (sign aligned)"), rel.frob(align.rows(unname(f$fullalpha), ffa.ref), ffa.ref), tol.signed)
    fW.ref <- read.mat(case, "fastruvIII.W", dims)
    check(paste(case, "fastruvIII W (sign aligned)"), rel.frob(t(align.rows(t(unname(f$W)), t(fW.ref))), fW.ref), tol.signed)

    # ---- ruvIIIMultipleK, SE branch: assays RUV_K<k>_on_HTseq = t(newY[1:m1, ]) (ruvIII.R:154-158) ----
    se.multi <- ruvIIIMultipleK(se.obj = se, assay.name = "HTseq", apply.log = FALSE, replicate.data = replicate.data, ctl = ctl,
                                k = ks.multi, assess.se.obj = FALSE, save.se.obj = TRUE, verbose = FALSE)
    m1 <- ncol(assay.mat)
    for (kv in ks.multi) {
        ref <- read.mat(case, paste0("ruvIIIMultipleK.k", kv, ".newY"), dims)
        got <- assay(se.multi, paste0("RUV_K", kv, "_on_HTseq"))
        check(paste0(case, " ruvIIIMultipleK k=", kv, " assay"), rel.frob(unname(as.matrix(got)), t(ref[1:m1, , drop = FALSE])), tol.newY)
    }

    if (case == "mid") {
        # ---- apply.log = TRUE on counts (ruvIII.R:88-89) ----
        counts <- read.mat(case, "counts", dims)
        dimnames(counts) <- dimnames(assay.mat)
        se.c <- SummarizedExperiment(assays = list(HTseq = counts))
        o2 <- ruvIII(se.obj = se.c, assay.name = "HTseq", apply.log = TRUE, pseudo.count = 1, replicate.data = replicate.data,
                     ctl = ctl, k = k, return.info = TRUE, save.se.obj = FALSE, assess.se.obj = FALSE, verbose = FALSE)
        check(paste(case, "ruvIII apply.log newY"), rel.frob(unname(o2$newY), read.mat(case, "ruvIII.log.newY", dims)), tol.newY)
        # ---- eta != NULL: ruv::RUV1 pre-step (ruvIII.R:116); pins the oracle's recollection of ruv::RUV1 / design.matrix ----
        eta <- read.mat(case, "eta", dims)                                   # n x 1 gene-wise covariate
        o3 <- ruvIII(se.obj = se, assay.name = "HTseq", apply.log = FALSE, replicate.data = replicate.data, ctl = ctl, k = k,
                     eta = eta, include.intercept = TRUE, return.info = TRUE, save.se.obj = FALSE, assess.se.obj = FALSE,
                     verbose = FALSE)
        check(paste(case, "ruvIII eta (RUV1) newY"), rel.frob(unname(o3$newY), read.mat(case, "ruvIII.eta.newY", dims)), tol.newY)
    }
}
if (!have.pkgs) cat("NOTE: the RUV-III fixtures were NOT replayed (Bioconductor packages missing); only the base-R statistics were checked\n")
cat(if (failures == 0) "PARITY PINNED: every fixture that was replayed reproduces under R\n"
    else sprintf("%d comparison(s) above tolerance -- the numpy restatement and the reference disagree\n", failures))
quit(status = if (failures == 0) 0 else 1)
