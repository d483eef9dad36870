## Ordered neighbour search (contract of the reference's R/nn.R:27-76): both functions fill nnIndx / nnDist / nnIndxLU
## IN PLACE through .Call and return them with the run time.  coords must already be in NNGP order.

nn.storage <- function(n, m){
    n.indx <- (1+m)/2*m + (n-m-1)*m
    list(nnIndx = integer(n.indx), nnDist = double(n.indx), nnIndxLU = integer(2*n))
}

mkNNIndx <- function(coords, m, n.omp.threads = 1){
    n <- nrow(coords)
    s <- nn.storage(n, m)
    ptm <- proc.time()
    .Call("mkNNIndx", as.integer(n), as.integer(m), as.double(coords), s$nnIndx, s$nnDist, s$nnIndxLU,
          as.integer(n.omp.threads))
    list(run.time = proc.time() - ptm, nnIndx = s$nnIndx, nnDist = s$nnDist, nnIndxLU = matrix(s$nnIndxLU, n, 2))
}

mkNNIndxCB <- function(coords, m, n.omp.threads = 1){
    n <- nrow(coords)
    s <- nn.storage(n, m)
    ptm <- proc.time()
    .Call("mkNNIndxCB", as.integer(n), as.integer(m), as.double(coords), s$nnIndx, s$nnDist, s$nnIndxLU,
          as.integer(n.omp.threads))
    list(run.time = proc.time() - ptm, nnIndx = s$nnIndx, nnDist = s$nnDist, nnIndxLU = matrix(s$nnIndxLU, n, 2))
}

## search.type "brute" or "cb" -> neighbour structure of the ordered coordinates
nn.search <- function(coords, m, search.type, n.omp.threads, verbose){
    if(!search.type %in% c("brute", "cb")){stop("error: search.type must be either 'brute' or 'cb'")}
    if(verbose){
        cat("----------------------------------------\n\tBuilding the neighbor list\n----------------------------------------\n")
    }
    if(search.type == "brute") mkNNIndx(coords, m, n.omp.threads) else mkNNIndxCB(coords, m, n.omp.threads)
}
