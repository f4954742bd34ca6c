/* R_ext/Rdynload.h stand-in (see ../Rinternals.h) */
#ifndef MOCK_RDYNLOAD_H
#define MOCK_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct DllInfo DllInfo;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e);
int R_useDynamicSymbols(DllInfo* info, int value);
#ifdef __cplusplus
}
#endif
#endif
