#ifndef RSHIM_RDYNLOAD_H
#define RSHIM_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef struct _DllInfo DllInfo;
typedef int Rboolean;
int R_registerRoutines(DllInfo *info, const R_CMethodDef *c, const R_CallMethodDef *call, const void *f, const void *e);
int R_useDynamicSymbols(DllInfo *info, int value);
int R_forceSymbols(DllInfo *info, int value);
#ifdef __cplusplus
}
#endif
#endif
