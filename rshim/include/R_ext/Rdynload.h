/* rshim: subset of <R_ext/Rdynload.h> -- just enough for R_init_<pkg>(). */
#ifndef RSHIM_RDYNLOAD_H
#define RSHIM_RDYNLOAD_H

#include "../Rinternals.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *(*DL_FUNC)();

typedef struct {
	const char *name;
	DL_FUNC fun;
	int numArgs;
} R_CallMethodDef;
typedef R_CallMethodDef R_CMethodDef;
typedef R_CallMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;

typedef struct _DllInfo {
	const char *name;
	const R_CallMethodDef *call_methods;
	int n_call_methods;
	int use_dynamic_symbols;
} DllInfo;

int R_registerRoutines(DllInfo *info, const R_CMethodDef *const croutines,
		       const R_CallMethodDef *const call_routines,
		       const R_FortranMethodDef *const fortran_routines,
		       const R_ExternalMethodDef *const external_routines);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);

/* shim-only: look up a registered .Call routine by name (NULL if absent);
   '*nargs' receives its registered arity. */
DL_FUNC rshim_lookup_call_routine(const DllInfo *info, const char *name,
				  int *nargs);

#ifdef __cplusplus
}
#endif

#endif  /* RSHIM_RDYNLOAD_H */
