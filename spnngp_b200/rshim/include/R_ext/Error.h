#ifndef RSHIM_ERROR_H
#define RSHIM_ERROR_H
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
void Rf_error(const char *fmt, ...) __attribute__((noreturn));
#else
void Rf_error(const char *fmt, ...);
#endif
void Rf_warning(const char *fmt, ...);
#ifdef __cplusplus
}
#endif
#endif
