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
# COMPILE-/RUN-UNTESTED in the build container (no R there); base R only apart from the reference's own imports.

args <- commandArgs(trailingOnly = TRUE)
this.file <- sub("^--file=", "", grep("^--file=", commandArgs(FALSE), value = TRUE))
golden.dir <- if (length(args) >= 1) args[1] else if (length(this.file)) dirname(normalizePath(this.file)) else "."
ref.dir <- if (length(args) >= 2) args[2] else Sys.getenv("RUVIIIPRPS_SRC", "/root/reference/R")
tol.newY <- 1e-9
tol.signed <- 1e-7

suppressPackageStartupMessages({
    library(SummarizedExperiment)
    library(ruv)
    library(BiocSingular)
})
if (requireNamespace("RUVIIIPRPS", quietly = TRUE)) {
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
message("ruv ", as.character(utils::packageVersion("ruv")), ", BiocSingular ", as.character(utils::packageVersion("BiocSingular")),
        ", ", R.version.string)

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

for (case in c("tiny", "mid")) {
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
cat(if (failures == 0) "PARITY PINNED: every fixture reproduces under the reference\n"
    else sprintf("%d comparison(s) above tolerance -- the numpy restatement and the reference disagree\n", failures))
quit(status = if (failures == 0) 0 else 1)
