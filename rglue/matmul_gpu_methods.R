### rglue/matmul_gpu_methods.R — the R side of the GPU drop-in.  Add this file to MatrixExtra's R/ directory
### (collated after R/matmul.R) when the package is built with -DMATRIXEXTRA_USE_MXGPU (rglue/mxgpu.patch).
###
### Nothing in R/matmul.R changes: its S4 methods keep calling the ten Rcpp exports, which
### rglue/matmul_gpu_glue.cpp re-implements on libmxgpu.so with identical signatures.  This file ADDS
###   1. the three signatures MatrixExtra leaves to the `Matrix` package (no setMethod for them in
###      R/matmul.R:200-767): crossprod(RsparseMatrix, matrix), CsparseMatrix %*% matrix, matrix %*% RsparseMatrix,
###      on the device CSR->CSC transpose (crossprod_csr_dense_{numeric,float32}, rglue/matmul_gpu_glue.cpp);
###   2. a device-resident sparse matrix class, `gpuRsparse`, for code that multiplies ONE matrix many times
###      (vignettes/Introducing_MatrixExtra.Rmd:454-476: `X %*% coefs` inside optim) — the CSR crosses PCIe once;
###   3. session settings: several GPUs for one product, and the operand cache of the unchanged exports.
### Same idioms as the methods they sit next to (R/matmul.R:436-469): dimension check, `mode<-"double"`,
### `as.csr.matrix`, `check_valid_matrix`, `set_dimnames`, MatrixExtra.nthreads.

mxgpu_nthreads <- function() {
    nthreads <- getOption("MatrixExtra.nthreads", default=parallel::detectCores())
    max(as.integer(nthreads), 1L)
}

#### 1. products the reference leaves to the Matrix package ----

crossprod_csr_dense <- function(x, y) {
    check_dimensions_match(x, y, crossprod=TRUE)
    if (typeof(y) != "double") mode(y) <- "double"
    x <- as.csr.matrix(x)
    check_valid_matrix(x)
    res <- crossprod_csr_dense_numeric(x@p, x@j, x@x, ncol(x), y, mxgpu_nthreads())
    set_dimnames(res, x, y, crossprod=TRUE)
}

crossprod_csr_f32 <- function(x, y) {
    check_dimensions_match(x, y, crossprod=TRUE)
    x <- as.csr.matrix(x)
    check_valid_matrix(x)
    res <- new("float32", Data=crossprod_csr_dense_float32(x@p, x@j, x@x, ncol(x), y@Data, mxgpu_nthreads()))
    set_dimnames(res, x, y, crossprod=TRUE)
}

#' @rdname matmult
#' @export
setMethod("crossprod", signature(x="RsparseMatrix", y="matrix"), crossprod_csr_dense)
#' @rdname matmult
#' @export
setMethod("crossprod", signature(x="RsparseMatrix", y="float32"), crossprod_csr_f32)

## CSC(x) is CSR(t(x)): x %*% y == crossprod(t_shallow(x), y)
#' @rdname matmult
#' @export
setMethod("%*%", signature(x="CsparseMatrix", y="matrix"), function(x, y) {
    check_dimensions_match(x, y, matmult=TRUE)
    res <- crossprod_csr_dense(t_shallow(x), y)
    set_dimnames(res, x, y, matmult=TRUE)
})

## x %*% y == t(crossprod(y, t(x)))
#' @rdname matmult
#' @export
setMethod("%*%", signature(x="matrix", y="RsparseMatrix"), function(x, y) {
    check_dimensions_match(x, y, matmult=TRUE)
    res <- t(crossprod_csr_dense(y, t(x)))
    set_dimnames(res, x, y, matmult=TRUE)
})

#### 2. device-resident matrices ----

#' @title Sparse matrix kept in GPU memory
#' @description A `dgRMatrix` uploaded once with `as.gpu.csr()`; `%*%`, `crossprod` and `tcrossprod` with dense
#' operands then move only the dense operand to the device and the result back.  Released by `gpu.free()` or by
#' the garbage collector (the external pointer carries a finalizer).
#' @export
setClass("gpuRsparse", representation(ptr="externalptr", Dim="integer", Dimnames="list"))

#' @export
as.gpu.csr <- function(x, float64=TRUE, float32=FALSE) {
    x <- as.csr.matrix(x)
    check_valid_matrix(x)
    ptr <- as_gpu_csr(x@p, x@j, x@x, ncol(x), as.logical(float64), as.logical(float32))
    new("gpuRsparse", ptr=ptr, Dim=x@Dim, Dimnames=x@Dimnames)
}

#' @export
gpu.free <- function(x) invisible(gpu_csr_free(x@ptr))

setMethod("dim", "gpuRsparse", function(x) x@Dim)
setMethod("dimnames", "gpuRsparse", function(x) x@Dimnames)

## x %*% t(y)   (the handle form of tcrossprod_csr_dense, R/matmul.R:436-457)
tcrossprod_gpucsr_dense <- function(x, y) {
    check_dimensions_match(x, y, tcrossprod=TRUE)
    if (inherits(y, "float32")) {
        res <- new("float32", Data=gpu_csr_tcrossprod_dense_float32(x@ptr, y@Data, mxgpu_nthreads()))
    } else {
        if (typeof(y) != "double") mode(y) <- "double"
        res <- gpu_csr_tcrossprod_dense_numeric(x@ptr, y, mxgpu_nthreads())
    }
    set_dimnames(res, x, y, tcrossprod=TRUE)
}
## x %*% t(y) with the sparse matrix on the right (tcrossprod_dense_csr, R/matmul.R:283-303)
tcrossprod_dense_gpucsr <- function(x, y) {
    check_dimensions_match(x, y, tcrossprod=TRUE)
    if (inherits(x, "float32")) {
        res <- new("float32", Data=gpu_csr_dense_tcrossprod_float32(x@Data, y@ptr, mxgpu_nthreads()))
    } else {
        if (typeof(x) != "double") mode(x) <- "double"
        res <- gpu_csr_dense_tcrossprod_numeric(x, y@ptr, mxgpu_nthreads())
    }
    set_dimnames(res, x, y, tcrossprod=TRUE)
}
crossprod_gpucsr_dense <- function(x, y) {
    check_dimensions_match(x, y, crossprod=TRUE)
    if (inherits(y, "float32")) {
        res <- new("float32", Data=gpu_csr_crossprod_dense_float32(x@ptr, y@Data, mxgpu_nthreads()))
    } else {
        if (typeof(y) != "double") mode(y) <- "double"
        res <- gpu_csr_crossprod_dense_numeric(x@ptr, y, mxgpu_nthreads())
    }
    set_dimnames(res, x, y, crossprod=TRUE)
}

for (cls in c("matrix", "float32")) {
    setMethod("tcrossprod", signature(x="gpuRsparse", y=cls), tcrossprod_gpucsr_dense)
    setMethod("tcrossprod", signature(x=cls, y="gpuRsparse"), tcrossprod_dense_gpucsr)
    setMethod("crossprod", signature(x="gpuRsparse", y=cls), crossprod_gpucsr_dense)
    setMethod("%*%", signature(x="gpuRsparse", y=cls), function(x, y) tcrossprod_gpucsr_dense(x, t(y)))
    setMethod("%*%", signature(x=cls, y="gpuRsparse"), function(x, y) t(crossprod_gpucsr_dense(y, t(x))))
}
setMethod("%*%", signature(x="gpuRsparse", y="numeric"), function(x, y) {
    if (ncol(x) != length(y)) stop("Matrix-vector dimensions do not match.")
    res <- gpu_csr_dvec_numeric(x@ptr, as.numeric(y), mxgpu_nthreads())
    if (!is.null(rownames(x))) names(res) <- rownames(x)
    matrix(res, ncol=1)
})

#### 3. session settings ----

#' @title GPU settings of the multiplication path
#' @param gpus number of GPUs of the box ONE product is spread over (row blocks, one host thread per device)
#' @param cache_mb device memory (MiB) the unchanged `%*%` / `crossprod` / `tcrossprod` methods may use to keep the
#' sparse matrices (and dense operands) they were last called with; 0 turns it off.  Off by default because R
#' objects can be modified in place (`options(MatrixExtra.inplace_sort=TRUE)`): call `mxgpu.settings(cache_mb=0)`
#' or use `as.gpu.csr()` when that is a concern.
#' @export
mxgpu.settings <- function(gpus=0L, cache_mb=-1L) mxgpu_configure(as.integer(gpus), as.integer(cache_mb))

.onLoad_mxgpu <- function(libname, pkgname) {
    ## also read by the glue itself on its first call: MATRIXEXTRA_GPUS, MATRIXEXTRA_GPU_CACHE_MB
    gpus <- suppressWarnings(as.integer(Sys.getenv("MATRIXEXTRA_GPUS", "0")))
    cache <- suppressWarnings(as.integer(Sys.getenv("MATRIXEXTRA_GPU_CACHE_MB", "-1")))
    if (!is.na(gpus) && gpus > 1L || !is.na(cache) && cache >= 0L)
        try(mxgpu_configure(if (is.na(gpus)) 0L else gpus, if (is.na(cache)) -1L else cache), silent=TRUE)
}
