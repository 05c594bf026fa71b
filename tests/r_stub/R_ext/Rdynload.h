/* stand-in for <R_ext/Rdynload.h> (see ../Rinternals.h) */
#ifndef ZZ_R_STUB_RDYNLOAD_H
#define ZZ_R_STUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo *info, const void *cRoutines, const R_CallMethodDef *callRoutines, const void *fortranRoutines, const void *externalRoutines);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
#endif
