// rshim: the few Rmath entry points the NNGP hot path uses.  OUR restatements:
//  * qnorm: Wichura (1988) algorithm AS241 (PPND16), the algorithm R's qnorm5 documents.
//  * bessel_k / bessel_k_ex: modified Bessel function of the second kind via libstdc++'s
//    std::cyl_bessel_k (within 1.2e-14 relative of mpmath over the tested range);
//    R's own nmath/bessel_k.c (Campbell's RKBESL) is NOT reproduced here.  `bk` is R's caller-
//    provided work array; unused by this implementation.
//  * gammafn / lgammafn: std::tgamma / std::lgamma.
#include <cmath>
#include <limits>
#include "include/Rmath.h"

extern "C" {

double qnorm(double p, double mu, double sigma, int lower_tail, int log_p) {
  if (std::isnan(p) || std::isnan(mu) || std::isnan(sigma)) return p + mu + sigma;
  if (log_p) p = exp(p);
  if (!lower_tail) p = 1.0 - p;
  if (p <= 0) return -std::numeric_limits<double>::infinity();
  if (p >= 1) return std::numeric_limits<double>::infinity();
  if (sigma < 0) return std::numeric_limits<double>::quiet_NaN();
  if (sigma == 0) return mu;

  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                   + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                + 133.14166789178437745) * r + 3.387132872796366608)
          / (((((((r * 5226.495278852854561 + 28729.085735721942674) * r + 39307.89580009271061) * r
                 + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
              + 42.313330701600911252) * r + 1.0);
  } else {
    r = (q < 0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
      r += -1.6;
      val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r
                + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
             + 4.6303378461565452959) * r + 1.42343711074968357734)
            / (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r
                   + 0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r
                + 2.05319162663775882187) * r + 1.0);
    } else {
      r += -5.0;
      val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r
                + 0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r
             + 5.4637849111641143699) * r + 6.6579046435011037772)
            / (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                   + 7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r
                + 0.59983220655588793769) * r + 1.0);
    }
    if (q < 0.0) val = -val;
  }
  return mu + sigma * val;
}

double pnorm(double x, double mu, double sigma, int lower_tail, int log_p) {
  double z = (x - mu) / sigma;
  double p = lower_tail ? 0.5 * erfc(-z / M_SQRT2) : 0.5 * erfc(z / M_SQRT2);
  return log_p ? log(p) : p;
}

double dnorm(double x, double mu, double sigma, int give_log) {
  double z = (x - mu) / sigma;
  double l = -0.5 * z * z - log(sigma) - 0.918938533204672741780329736406;
  return give_log ? l : exp(l);
}

double gammafn(double x) { return std::tgamma(x); }
double lgammafn(double x) { return std::lgamma(x); }

double bessel_k(double x, double alpha, double expo) {
  if (std::isnan(x) || std::isnan(alpha)) return x + alpha;
  if (x < 0) return std::numeric_limits<double>::quiet_NaN();
  if (alpha < 0) alpha = -alpha;  // K_{-nu} = K_nu
  if (x == 0) return std::numeric_limits<double>::infinity();
  double k = std::cyl_bessel_k(alpha, x);
  if (expo == 2.0) k *= exp(x);
  return k;
}

double bessel_k_ex(double x, double alpha, double expo, double *bk) {
  (void)bk;
  return bessel_k(x, alpha, expo);
}

}  // extern "C"
