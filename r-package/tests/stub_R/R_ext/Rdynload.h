/* stand-in for R_ext/Rdynload.h (see ../Rinternals.h) */
#ifndef STUB_RDYNLOAD_H
#define STUB_RDYNLOAD_H
#include "../Rinternals.h"
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
typedef struct { const char* name; DL_FUNC fun; int numArgs; void* types; } R_CMethodDef;
int R_registerRoutines(DllInfo* info, const R_CMethodDef* const croutines, const R_CallMethodDef* const callRoutines,
                       const void* const fortranRoutines, const void* const externalRoutines);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#endif
