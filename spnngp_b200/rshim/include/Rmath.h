/* rshim stand-in for <Rmath.h> (subset). */
#ifndef RSHIM_RMATH_H
#define RSHIM_RMATH_H
#include <math.h>
#ifndef M_PI
#define M_PI 3.141592653589793238462643383280
#endif
#ifndef M_LN2
#define M_LN2 0.693147180559945309417232121458
#endif
#ifndef M_SQRT2
#define M_SQRT2 1.414213562373095048801688724210
#endif
#ifdef __cplusplus
extern "C" {
#endif
double rnorm(double mu, double sigma);
double runif(double a, double b);
double rgamma(double shape, double scale);
double rexp(double scale);
double pnorm(double x, double mu, double sigma, int lower_tail, int log_p);
double qnorm(double p, double mu, double sigma, int lower_tail, int log_p);
double dnorm(double x, double mu, double sigma, int give_log);
double gammafn(double x);
double lgammafn(double x);
double bessel_k(double x, double alpha, double expo);
double bessel_k_ex(double x, double alpha, double expo, double *bk);
#ifdef __cplusplus
}
#endif
#endif
