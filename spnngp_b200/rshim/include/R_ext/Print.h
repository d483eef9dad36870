#ifndef RSHIM_PRINT_H
#define RSHIM_PRINT_H
#ifdef __cplusplus
extern "C" {
#endif
void Rprintf(const char *fmt, ...);
void REprintf(const char *fmt, ...);
void R_FlushConsole(void);
#ifdef __cplusplus
}
#endif
#endif
