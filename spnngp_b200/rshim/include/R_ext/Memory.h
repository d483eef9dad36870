#ifndef RSHIM_MEMORY_H
#define RSHIM_MEMORY_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* transient storage released by rshim_reset() (R frees it when .Call returns) */
char *R_alloc(size_t n, int size);
#ifdef __cplusplus
}
#endif
#endif
