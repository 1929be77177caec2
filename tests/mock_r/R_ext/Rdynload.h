/* see R.h in this directory */
#ifndef MOCK_RDYNLOAD_H
#define MOCK_RDYNLOAD_H
#include "../Rinternals.h"
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct mock_dllinfo { const R_CallMethodDef* call_methods; int dynamic_symbols; } DllInfo;
int R_registerRoutines(DllInfo* info, const void* c_methods, const R_CallMethodDef* call_methods, const void* f_methods,
                       const void* ext_methods);
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value);
#endif
