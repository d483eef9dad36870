## spConjNNGP(): same arguments and return value as the reference's R/spConjNNGP.R:1-6 for the NNGP (no knots) model.
## The k-fold loop of R/spConjNNGP.R:286-349 -- split, ordered search on the retained rows, RANN query for the held-out
## rows, one cNNGP per fold, crps / rmspe per parameter set -- is written here with the same .Call("cNNGP", ...) contract
## (18 arguments, R/spConjNNGP.R:336-338, 378-380); every call lands on the GPU.

crps <- function(y, y.hat, y.var){
    sd <- sqrt(y.var); z <- (y - y.hat)/sd
    mean(sd*(1/sqrt(pi) - 2*dnorm(z) - z*(2*pnorm(z) - 1)))
}
rmspe <- function(y, y.hat){sqrt(mean((y - y.hat)^2))}

call.cNNGP <- function(y, X, coords, theta.alpha, nn, X.0, coords.0, nn.indx.0, n.neighbors, sigma.sq.IG, cov.model.indx,
                       n.omp.threads, verbose){
    n <- nrow(X); p <- ncol(X); g <- ncol(theta.alpha)
    n.0 <- if(is.null(X.0)) 0L else nrow(X.0)
    if(n.0 == 0){X.0 <- matrix(0, 1, p); coords.0 <- matrix(0, 1, 2); nn.indx.0 <- 0L}
    storage.mode(X) <- "double"; storage.mode(coords) <- "double"; storage.mode(theta.alpha) <- "double"
    storage.mode(X.0) <- "double"; storage.mode(coords.0) <- "double"
    .Call("cNNGP", as.double(y), X, as.integer(p), as.integer(n), coords, theta.alpha, as.integer(nn$nnIndx), as.integer(nn$nnIndxLU),
          X.0, coords.0, as.integer(n.0), as.integer(nn.indx.0), as.integer(g),
          as.integer(n.neighbors), as.double(sigma.sq.IG), as.integer(cov.model.indx), as.integer(n.omp.threads), as.integer(verbose))
}

spConjNNGP <- function(formula, data = parent.frame(), coords, knots, n.neighbors = 15,
                       theta.alpha, sigma.sq.IG, cov.model = "exponential",
                       k.fold = 5, score.rule = "crps",
                       X.0, coords.0,
                       n.omp.threads = 1, search.type = "cb", ord, return.neighbor.info = TRUE,
                       neighbor.info, fit.rep = FALSE, n.samples, verbose = TRUE, ...){

    if(!missing(knots)){stop("error: the SLGP (knots) model is not part of this build; use the CPU package")}
    ptm <- proc.time()
    mm <- parse.model(formula, data)
    y <- mm$y; X <- mm$X; p <- ncol(X); n <- nrow(X)
    if(missing(coords)){stop("error: coords must be specified")}
    if(is.character(coords) && is.data.frame(data)){coords <- as.matrix(data[, coords])}
    coords <- as.matrix(coords)
    if(ncol(coords) != 2 || nrow(coords) != n){stop("error: coords must have two columns and one row per observation")}
    if(missing(ord)){ord <- order(coords[,1])}
    coords <- coords[ord, , drop = FALSE]; X <- X[ord, , drop = FALSE]; y <- y[ord]

    cov.model.indx <- match(cov.model, cov.model.names) - 1
    if(is.na(cov.model.indx)){stop("error: specified cov.model '", cov.model, "' is not a valid option")}
    if(missing(sigma.sq.IG) || length(sigma.sq.IG) != 2){stop("error: sigma.sq.IG must be a vector of length two")}
    if(!score.rule %in% c("rmspe", "crps")){stop("error: score.rule must be 'rmspe' or 'crps'")}

    ## theta.alpha: one column per parameter set, rows phi, alpha[, nu] (R/spConjNNGP.R:203-211)
    if(missing(theta.alpha)){stop("error: theta.alpha must be specified")}
    if(is.vector(theta.alpha)){theta.alpha <- t(as.matrix(theta.alpha))}
    want <- c("phi", "alpha", if(cov.model == "matern") "nu")
    if(!all(want %in% colnames(theta.alpha))){stop("error: column names of theta.alpha must include ", paste(want, collapse = ", "))}
    grid <- theta.alpha
    theta.alpha <- t(theta.alpha[, want, drop = FALSE])
    g <- ncol(theta.alpha)

    ## k-fold scores of every parameter set
    k.fold.scores <- NULL
    if(g > 1 && !is.na(k.fold) && k.fold > 1){
        scores <- matrix(0, g, 2, dimnames = list(NULL, c("rmspe", "crps")))
        for(f in seq_len(k.fold)){
            ho <- sort(sample(seq_len(n), floor(n/k.fold)))
            nn.mod <- nn.search(coords[-ho, , drop = FALSE], n.neighbors, search.type, n.omp.threads, FALSE)
            nn.indx.0 <- nn2(coords[-ho, , drop = FALSE], coords[ho, , drop = FALSE], k = n.neighbors)$nn.idx - 1
            out <- call.cNNGP(y[-ho], X[-ho, , drop = FALSE], coords[-ho, , drop = FALSE], theta.alpha, nn.mod,
                              X[ho, , drop = FALSE], coords[ho, , drop = FALSE], nn.indx.0, n.neighbors, sigma.sq.IG,
                              cov.model.indx, n.omp.threads, FALSE)
            y.hat <- matrix(out$y.0.hat, ncol = g); y.var <- matrix(out$y.0.hat.var, ncol = g)
            for(j in seq_len(g)){
                scores[j, "crps"] <- scores[j, "crps"] + crps(y[ho], y.hat[, j], y.var[, j])
                scores[j, "rmspe"] <- scores[j, "rmspe"] + rmspe(y[ho], y.hat[, j])
            }
        }
        k.fold.scores <- cbind(grid, scores/k.fold)
        best <- which.min(k.fold.scores[, score.rule])
        theta.alpha <- theta.alpha[, best, drop = FALSE]; g <- 1
    }

    ## final fit (and prediction) with the chosen set
    nn <- if(!missing(neighbor.info)) list(nnIndx = neighbor.info$n.indx, nnIndxLU = neighbor.info$nn.indx.lu)
          else nn.search(coords, n.neighbors, search.type, n.omp.threads, verbose)
    has.0 <- !missing(X.0) && !missing(coords.0)
    nn.indx.0 <- if(has.0) nn2(coords, as.matrix(coords.0), k = n.neighbors)$nn.idx - 1 else NULL
    out <- call.cNNGP(y, X, coords, theta.alpha, nn, if(has.0) as.matrix(X.0) else NULL, if(has.0) as.matrix(coords.0) else NULL,
                      nn.indx.0, n.neighbors, sigma.sq.IG, cov.model.indx, n.omp.threads, verbose)

    out$beta.hat <- matrix(out$beta.hat, nrow = p, dimnames = list(mm$x.names, NULL))
    out$beta.var <- out$ab[2]*matrix(out$bB.inv, p, p)/(out$ab[1] - 1)
    out$sigma.sq.hat <- out$ab[2]/(out$ab[1] - 1)
    out$sigma.sq.var <- out$ab[2]^2/((out$ab[1] - 1)^2*(out$ab[1] - 2))
    out$theta.alpha <- t(theta.alpha)
    if(!is.null(k.fold.scores)){out$k.fold.scores <- k.fold.scores}
    out$call <- match.call(); out$type <- c("NNGP", "conjugate", "gaussian")
    out$y <- y; out$X <- X; out$coords <- coords; out$ord <- ord
    out$n.neighbors <- n.neighbors; out$cov.model <- cov.model; out$cov.model.indx <- cov.model.indx
    out$sigma.sq.IG <- sigma.sq.IG
    if(return.neighbor.info){
        out$neighbor.info <- list(n.neighbors = n.neighbors, n.indx = nn$nnIndx, nn.indx.lu = nn$nnIndxLU, ord = ord)
    }
    out$run.time <- proc.time() - ptm
    class(out) <- "NNGP"
    out
}
