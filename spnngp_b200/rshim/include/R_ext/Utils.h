#ifndef RSHIM_UTILS_H
#define RSHIM_UTILS_H
#ifdef __cplusplus
extern "C" {
#endif
void R_CheckUserInterrupt(void);
typedef int Rboolean;  /* R_ext/Boolean.h */
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
Rboolean R_ToplevelExec(void (*fun)(void *), void *data);
void rshim_raise_interrupt(void);  /* stand-in only: make the next R_ToplevelExec report an interrupt */
/* restatement of R's src/main/sort.c rsort_with_index (Shell sort, NaN last) */
void rsort_with_index(double *x, int *indx, int n);
#ifdef __cplusplus
}
#endif
#endif
