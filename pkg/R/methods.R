## The small S3 surface of an NNGP fit (print / summary / fitted / residuals), written for the two fit types this
## build produces.

print.NNGP <- function(x, ...){
    cat("\nCall:", deparse(x$call), "\n")
    cat("Model class is", paste(x$type, collapse = " "), "\n")
    cat("Fit on", nrow(x$X), "observations with", x$n.neighbors, "neighbors;", x$cov.model, "covariance.\n")
    invisible(x)
}

summary.NNGP <- function(object, sub.sample, quantiles = c(0.025, 0.25, 0.5, 0.75, 0.975), digits = max(3L, getOption("digits") - 3L), ...){
    print(object)
    if(object$type[2] == "response"){
        s.indx <- sub.sample.indx(nrow(object$p.beta.samples), if(missing(sub.sample)) NULL else sub.sample)
        smp <- cbind(as.matrix(object$p.beta.samples), as.matrix(object$p.theta.samples))[s.indx, , drop = FALSE]
        print(noquote(apply(t(apply(smp, 2, quantile, probs = quantiles)), 2, function(v) formatC(v, format = "f", digits = digits))))
    }else{
        est <- cbind(Estimate = object$beta.hat[, 1], Variance = diag(object$beta.var))
        est <- rbind(est, sigma.sq = c(object$sigma.sq.hat, object$sigma.sq.var))
        print(noquote(apply(est, 2, function(v) formatC(v, format = "f", digits = digits))))
    }
    invisible(object)
}

fitted.NNGP <- function(object, sub.sample, ...){
    if(object$type[2] == "conjugate"){return(as.vector(object$X %*% object$beta.hat[, 1]))}
    if(is.null(object$y.rep.samples)){stop("error: fit with fit.rep = TRUE to obtain replicated data")}
    list(y.rep.quants = t(apply(object$y.rep.samples, 1, quantile, probs = c(0.5, 0.025, 0.975))),
         y.rep.samples = object$y.rep.samples, sub.sample = object$s.indx)
}

residuals.NNGP <- function(object, sub.sample, ...){
    f <- fitted(object)
    if(is.list(f)){object$y - f$y.rep.samples}else{object$y - f}
}
