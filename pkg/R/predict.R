## predict.NNGP(): the arguments of the reference's R/generics.R:532 for fits of type "response" (posterior predictive
## draws through .Call("rNNGPPredict", ...), 17 arguments, R/generics.R:674-675) and "conjugate".

sub.sample.indx <- function(n.samples, sub.sample){
    s <- if(missing(sub.sample) || is.null(sub.sample)) list() else sub.sample
    start <- if("start" %in% names(s)) s$start else floor(0.5*n.samples)
    end <- if("end" %in% names(s)) s$end else n.samples
    thin <- if("thin" %in% names(s)) s$thin else 1
    seq(as.integer(start), as.integer(end), by = as.integer(thin))
}

predict.NNGP <- function(object, X.0, coords.0, sub.sample, n.omp.threads = 1, verbose = TRUE, n.report = 100, ...){
    if(missing(X.0) || missing(coords.0)){stop("error: X.0 and coords.0 must be specified")}
    X.0 <- as.matrix(X.0); coords.0 <- as.matrix(coords.0)
    if(ncol(coords.0) != 2){stop("error: coords.0 must have two columns")}
    if(ncol(X.0) != ncol(object$X)){stop(paste0("error: X.0 must have ", ncol(object$X), " columns"))}
    if(nrow(X.0) != nrow(coords.0)){stop("error: X.0 and coords.0 must have the same number of rows")}
    X <- object$X; y <- object$y; coords <- object$coords
    n <- nrow(X); p <- ncol(X); q <- nrow(X.0); m <- object$n.neighbors
    ptm <- proc.time()
    out <- list()

    if(object$type[2] == "conjugate"){
        nn <- list(nnIndx = object$neighbor.info$n.indx, nnIndxLU = object$neighbor.info$nn.indx.lu)
        if(is.null(nn$nnIndx)){stop("error: the fit was made with return.neighbor.info = FALSE")}
        nn.indx.0 <- nn2(coords, coords.0, k = m)$nn.idx - 1
        theta.alpha <- t(object$theta.alpha)
        res <- call.cNNGP(y, X, coords, theta.alpha, nn, X.0, coords.0, nn.indx.0, m, object$sigma.sq.IG,
                          object$cov.model.indx, n.omp.threads, verbose)
        out$y.0.hat <- res$y.0.hat; out$y.0.hat.var <- res$y.0.hat.var
    }else if(object$type[2] == "response"){
        s.indx <- sub.sample.indx(nrow(object$p.beta.samples), if(missing(sub.sample)) NULL else sub.sample)
        beta <- t(as.matrix(object$p.beta.samples)[s.indx, , drop = FALSE])     ## p x S
        theta <- t(as.matrix(object$p.theta.samples)[s.indx, , drop = FALSE])   ## nTheta x S
        nn.indx.0 <- nn2(coords, coords.0, k = m)$nn.idx - 1
        storage.mode(X) <- "double"; storage.mode(coords) <- "double"; storage.mode(X.0) <- "double"
        storage.mode(coords.0) <- "double"; storage.mode(beta) <- "double"; storage.mode(theta) <- "double"
        res <- .Call("rNNGPPredict", X, as.double(y), coords, as.integer(n), as.integer(p), as.integer(m), X.0, coords.0,
                     as.integer(q), as.integer(nn.indx.0), beta, theta, as.integer(length(s.indx)),
                     as.integer(object$cov.model.indx), as.integer(n.omp.threads), as.integer(verbose), as.integer(n.report))
        out$p.y.0 <- matrix(res$p.y.0, nrow = q)
        out$sub.sample <- s.indx
    }else{
        stop("error: this build predicts from 'response' and 'conjugate' fits")
    }
    out$run.time <- proc.time() - ptm
    class(out) <- "predict.NNGP"
    out
}

print.predict.NNGP <- function(x, ...){
    cat("predict.NNGP object with elements:", paste(names(x), collapse = ", "), "\n")
    invisible(x)
}
