## spNNGP(): same arguments and return value as the reference's R/spNNGP.R:1-5 for family = "gaussian",
## method = "response"; the sampler behind .Call("rNNGP", ...) (28 arguments, R/spNNGP.R:401-405) runs on the GPU.

cov.model.names <- c("exponential", "spherical", "matern", "gaussian")  ## index order fixed by the C side (0..3)

parse.model <- function(formula, data){
    if(!inherits(formula, "formula")){stop("error: formula is misspecified")}
    mf <- model.frame(formula, data)
    X <- model.matrix(terms(mf), mf)
    list(y = as.double(model.response(mf)), X = X, x.names = colnames(X))
}

need <- function(lst, name, what){
    if(!name %in% names(lst)){stop(paste0("error: ", name, " must be specified in ", what))}
    lst[[name]]
}

spNNGP <- function(formula, data = parent.frame(), coords, method = "response", family = "gaussian", weights, n.neighbors = 15,
                   starting, tuning, priors, cov.model = "exponential",
                   n.samples, n.omp.threads = 1, search.type = "cb", ord,
                   return.neighbor.info = FALSE, neighbor.info,
                   fit.rep = FALSE, sub.sample, verbose = TRUE, n.report = 100, ...){

    extra <- names(list(...))
    if(length(extra)){for(a in extra) warning(paste0("'", a, "' is not an argument"))}
    if(family != "gaussian" || method != "response"){
        stop("error: this build provides family = 'gaussian' with method = 'response' (the B200 hot path); use the CPU package for the latent and binomial models")
    }
    ptm <- proc.time()

    mm <- parse.model(formula, data)
    y <- mm$y; X <- mm$X; p <- ncol(X); n <- nrow(X)

    ## coordinates: a matrix, or the names of two columns of data
    if(missing(coords)){stop("error: coords must be specified")}
    if(is.character(coords) && is.data.frame(data)){coords <- as.matrix(data[, coords])}
    coords <- as.matrix(coords)
    if(ncol(coords) != 2 || nrow(coords) != n){stop("error: coords must have two columns and one row per observation")}

    ## NNGP order (default: by the first coordinate) and neighbours
    if(!missing(neighbor.info)){
        ord <- neighbor.info$ord
        nn.indx <- neighbor.info$n.indx; nn.indx.lu <- neighbor.info$nn.indx.lu
        n.neighbors <- neighbor.info$n.neighbors
    }else{
        if(missing(ord)){ord <- order(coords[,1])}
        if(length(ord) != n){stop("error: supplied order vector ord must be of length n")}
    }
    coords <- coords[ord, , drop = FALSE]; X <- X[ord, , drop = FALSE]; y <- y[ord]
    if(n.neighbors >= n){stop("error: n.neighbors must be less than the number of observations")}

    cov.model.indx <- match(cov.model, cov.model.names) - 1
    if(is.na(cov.model.indx)){stop("error: specified cov.model '", cov.model, "' is not a valid option; choose from ", paste(cov.model.names, collapse = ", "), ".")}
    matern <- cov.model == "matern"

    ## priors, starting values, Metropolis tuning (names as in the reference)
    if(missing(priors)){stop("error: prior list for the parameters must be specified")}
    names(priors) <- tolower(names(priors))
    sigma.sq.IG <- as.double(need(priors, "sigma.sq.ig", "priors"))
    tau.sq.IG <- as.double(need(priors, "tau.sq.ig", "priors"))
    phi.Unif <- as.double(need(priors, "phi.unif", "priors"))
    nu.Unif <- if(matern) as.double(need(priors, "nu.unif", "priors")) else as.double(c(0, 0))
    if(missing(starting)){stop("error: starting value list for the parameters must be specified")}
    names(starting) <- tolower(names(starting))
    beta.starting <- if("beta" %in% names(starting)) as.double(starting[["beta"]]) else as.double(rep(0, p))
    if(length(beta.starting) != p){stop(paste0("error: starting values for beta must be of length p=", p))}
    sigma.sq.starting <- as.double(need(starting, "sigma.sq", "starting")[1])
    tau.sq.starting <- as.double(need(starting, "tau.sq", "starting")[1])
    phi.starting <- as.double(need(starting, "phi", "starting")[1])
    nu.starting <- if(matern) as.double(need(starting, "nu", "starting")[1]) else 0.0
    if(missing(tuning)){stop("error: tuning value vector for the spatial parameters must be specified")}
    names(tuning) <- tolower(names(tuning))
    sigma.sq.tuning <- as.double(need(tuning, "sigma.sq", "tuning")[1])
    tau.sq.tuning <- as.double(need(tuning, "tau.sq", "tuning")[1])
    phi.tuning <- as.double(need(tuning, "phi", "tuning")[1])
    nu.tuning <- if(matern) as.double(need(tuning, "nu", "tuning")[1]) else 0.0
    if(missing(n.samples)){stop("error: n.samples needs to be specified")}

    ## replicates: which iterations draw y.rep (start / end / thin of sub.sample)
    rep.indx <- integer(n.samples); n.rep <- 0L
    if(fit.rep){
        s <- if(missing(sub.sample)) list() else sub.sample
        start <- if("start" %in% names(s)) s$start else floor(0.5*n.samples)
        end <- if("end" %in% names(s)) s$end else n.samples
        thin <- if("thin" %in% names(s)) s$thin else 1
        s.indx <- seq(as.integer(start), as.integer(end), by = as.integer(thin))
        rep.indx[s.indx] <- 1L; n.rep <- length(s.indx)
    }

    if(missing(neighbor.info)){
        nn <- nn.search(coords, n.neighbors, search.type, n.omp.threads, verbose)
        nn.indx <- nn$nnIndx; nn.indx.lu <- as.integer(nn$nnIndxLU)
    }

    storage.mode(X) <- "double"; storage.mode(coords) <- "double"
    out <- .Call("rNNGP", y, X, as.integer(p), as.integer(n), as.integer(n.neighbors), coords, as.integer(cov.model.indx),
                 as.integer(nn.indx), as.integer(nn.indx.lu),
                 sigma.sq.IG, tau.sq.IG, phi.Unif, nu.Unif,
                 beta.starting, sigma.sq.starting, tau.sq.starting, phi.starting, nu.starting,
                 sigma.sq.tuning, tau.sq.tuning, phi.tuning, nu.tuning,
                 as.integer(n.samples), as.integer(n.omp.threads), as.integer(verbose), as.integer(n.report), as.integer(n.rep), rep.indx)

    out$p.beta.samples <- mcmc(matrix(out$p.beta.samples, ncol = p, byrow = TRUE, dimnames = list(NULL, mm$x.names)))
    theta.names <- c("sigma.sq", "tau.sq", "phi", if(matern) "nu")
    out$p.theta.samples <- mcmc(matrix(out$p.theta.samples, ncol = length(theta.names), byrow = TRUE, dimnames = list(NULL, theta.names)))
    if(fit.rep){out$y.rep.samples <- matrix(out$y.rep.samples, nrow = n); out$s.indx <- s.indx}

    out$call <- match.call(); out$type <- c("NNGP", "response", "gaussian")
    out$y <- y; out$X <- X; out$coords <- coords; out$ord <- ord
    out$n.neighbors <- n.neighbors; out$cov.model <- cov.model; out$cov.model.indx <- cov.model.indx
    out$starting <- starting; out$priors <- priors; out$tuning <- tuning
    if(return.neighbor.info){
        out$neighbor.info <- list(n.neighbors = n.neighbors, n.indx = nn.indx, nn.indx.lu = nn.indx.lu, ord = ord)
    }
    out$run.time <- proc.time() - ptm
    class(out) <- "NNGP"
    out
}
