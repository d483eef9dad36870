/* rshim: miniature stand-in for <R.h>, for building spNNGP-style .Call code on hosts without R.
 * Not R. Only the subset that the response/conjugate NNGP hot path touches.  When a real R is
 * available, build against R's own headers instead. */
#ifndef RSHIM_R_H
#define RSHIM_R_H
#include <stddef.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <limits.h>
#include "R_ext/Print.h"
#include "R_ext/Error.h"
#include "R_ext/Memory.h"
#include "R_ext/Utils.h"
#include "R_ext/Random.h"
#include "R_ext/RS.h"
#endif
