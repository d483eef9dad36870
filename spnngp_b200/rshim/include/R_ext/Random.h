#ifndef RSHIM_RANDOM_H
#define RSHIM_RANDOM_H
#ifdef __cplusplus
extern "C" {
#endif
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
double norm_rand(void);
double exp_rand(void);
#ifdef __cplusplus
}
#endif
#endif
