#ifndef RSHIM_UTILS_H
#define RSHIM_UTILS_H
#ifdef __cplusplus
extern "C" {
#endif
void R_CheckUserInterrupt(void);
/* restatement of R's src/main/sort.c rsort_with_index (Shell sort, NaN last) */
void rsort_with_index(double *x, int *indx, int n);
#ifdef __cplusplus
}
#endif
#endif
