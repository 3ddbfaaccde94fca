/* rshim runtime -- see include/Rinternals.h for what this is and is not. */
#include "include/Rinternals.h"
#include "include/R_ext/Rdynload.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* constants                                                           */

static SEXPREC nil_rec = { NILSXP, 0, 0, NULL, &nil_rec, &nil_rec, &nil_rec };
SEXP R_NilValue = &nil_rec;

static char na_string_payload[] = "NA";
static SEXPREC na_string_rec =
	{ CHARSXP, 0, 2, na_string_payload, &nil_rec, &nil_rec, &nil_rec };
SEXP R_NaString = &na_string_rec;

static char blank_string_payload[] = "";
static SEXPREC blank_string_rec =
	{ CHARSXP, 0, 0, blank_string_payload, &nil_rec, &nil_rec, &nil_rec };
SEXP R_BlankString = &blank_string_rec;

int R_NaInt = INT_MIN;
double R_NaReal, R_NaN, R_PosInf, R_NegInf;

/* R's NA_real_ is the NaN whose low 32 bits are 1954 (arithmetic.c). */
static double make_na_real(void)
{
	union { double d; uint64_t u; } x;
	x.u = ((uint64_t) 0x7FF00000u << 32) | 1954u;
	return x.d;
}

__attribute__((constructor)) static void rshim_init_constants(void)
{
	R_NaReal = make_na_real();
	R_NaN = NAN;
	R_PosInf = INFINITY;
	R_NegInf = -INFINITY;
}

int R_IsNA(double x)
{
	if (!isnan(x))
		return 0;
	union { double d; uint64_t u; } y;
	y.d = x;
	return (uint32_t) (y.u & 0xFFFFFFFFu) == 1954u;
}

int R_IsNaN(double x)
{
	if (!isnan(x))
		return 0;
	return !R_IsNA(x);
}

int R_finite(double x)
{
	return isfinite(x) != 0;
}

/* ------------------------------------------------------------------ */
/* conditions                                                          */

static jmp_buf error_jmpbuf;
static int error_jmpbuf_armed = 0;
static char last_error[2048];

#define MAX_WARNINGS 64
static char *warnings[MAX_WARNINGS];
static int nwarnings = 0;

void Rf_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(last_error, sizeof(last_error), fmt, ap);
	va_end(ap);
	if (error_jmpbuf_armed)
		longjmp(error_jmpbuf, 1);
	fprintf(stderr, "rshim: uncaught R error: %s\n", last_error);
	abort();
}

void Rf_warning(const char *fmt, ...)
{
	char buf[2048];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (nwarnings < MAX_WARNINGS)
		warnings[nwarnings++] = strdup(buf);
}

const char *rshim_last_error(void) { return last_error; }
int rshim_warning_count(void) { return nwarnings; }
const char *rshim_warning_message(int i)
{
	return (i >= 0 && i < nwarnings) ? warnings[i] : NULL;
}
void rshim_clear_warnings(void)
{
	for (int i = 0; i < nwarnings; i++)
		free(warnings[i]);
	nwarnings = 0;
}

/* ------------------------------------------------------------------ */
/* allocation                                                          */

static long live_objects = 0;
long rshim_live_objects(void) { return live_objects; }

static size_t elt_size(SEXPTYPE type)
{
	switch (type) {
	    case LGLSXP: case INTSXP: return sizeof(int);
	    case REALSXP:             return sizeof(double);
	    case CPLXSXP:             return sizeof(Rcomplex);
	    case RAWSXP:              return sizeof(Rbyte);
	    case STRSXP: case VECSXP: return sizeof(SEXP);
	    case CHARSXP:             return 1;
	}
	Rf_error("rshim: allocVector(): unsupported type %u", type);
}

static SEXP new_record(SEXPTYPE type, R_xlen_t n, void *data, int owns)
{
	SEXP x = (SEXP) malloc(sizeof(SEXPREC));
	if (x == NULL)
		Rf_error("rshim: out of memory");
	x->type = type;
	x->owns_data = owns;
	x->length = n;
	x->data = data;
	x->dim = x->names = x->dimnames = R_NilValue;
	x->finalizer = NULL;
	__atomic_add_fetch(&live_objects, 1, __ATOMIC_RELAXED);
	return x;
}

/* ---- external pointers ---- */
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
	(void) tag; (void) prot;
	return new_record(EXTPTRSXP, 0, p, 0);
}

void *R_ExternalPtrAddr(SEXP s)
{
	if (s == NULL || TYPEOF(s) != EXTPTRSXP)
		Rf_error("rshim: R_ExternalPtrAddr(): not an external pointer");
	return s->data;
}

void R_ClearExternalPtr(SEXP s)
{
	if (s != NULL && TYPEOF(s) == EXTPTRSXP)
		s->data = NULL;
}

void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit)
{
	(void) onexit;
	if (s == NULL || TYPEOF(s) != EXTPTRSXP)
		Rf_error("rshim: R_RegisterCFinalizerEx(): not an external "
			 "pointer");
	s->finalizer = fun;
}

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n)
{
	if (n < 0)
		Rf_error("rshim: allocVector(): negative length");
	size_t sz = elt_size(type);
	size_t nbytes = sz * (size_t) n + (type == CHARSXP ? 1 : 0);
	void *data = calloc(nbytes > 0 ? nbytes : 1, 1);
	if (data == NULL)
		Rf_error("rshim: cannot allocate vector of %zu bytes", nbytes);
	SEXP x = new_record(type, n, data, 1);
	if (type == VECSXP) {
		for (R_xlen_t i = 0; i < n; i++)
			((SEXP *) data)[i] = R_NilValue;
	} else if (type == STRSXP) {
		for (R_xlen_t i = 0; i < n; i++)
			((SEXP *) data)[i] = R_BlankString;
	}
	return x;
}

SEXP rshim_wrap_vector(SEXPTYPE type, R_xlen_t n, void *data)
{
	return new_record(type, n, data, 0);
}

SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol)
{
	SEXP x = Rf_allocVector(type, (R_xlen_t) nrow * ncol);
	SEXP d = Rf_allocVector(INTSXP, 2);
	INTEGER(d)[0] = nrow;
	INTEGER(d)[1] = ncol;
	x->dim = d;
	return x;
}

SEXP Rf_allocArray(SEXPTYPE type, SEXP dims)
{
	R_xlen_t n = 1;
	for (int i = 0; i < LENGTH(dims); i++)
		n *= INTEGER(dims)[i];
	SEXP x = Rf_allocVector(type, n);
	x->dim = Rf_duplicate(dims);
	return x;
}

SEXP Rf_duplicate(SEXP x)
{
	if (x == R_NilValue || x == R_NaString || x == R_BlankString)
		return x;
	SEXP y = Rf_allocVector(TYPEOF(x), XLENGTH(x));
	if (TYPEOF(x) == VECSXP || TYPEOF(x) == STRSXP) {
		for (R_xlen_t i = 0; i < XLENGTH(x); i++)
			((SEXP *) y->data)[i] =
				Rf_duplicate(((SEXP *) x->data)[i]);
	} else {
		memcpy(y->data, x->data, elt_size(TYPEOF(x)) * XLENGTH(x));
	}
	y->dim = Rf_duplicate(x->dim);
	y->names = Rf_duplicate(x->names);
	y->dimnames = Rf_duplicate(x->dimnames);
	return y;
}

SEXP Rf_mkChar(const char *s)
{
	size_t n = strlen(s);
	SEXP x = Rf_allocVector(CHARSXP, (R_xlen_t) n);
	memcpy(x->data, s, n + 1);
	return x;
}

SEXP Rf_mkString(const char *s)
{
	SEXP x = Rf_allocVector(STRSXP, 1);
	((SEXP *) x->data)[0] = Rf_mkChar(s);
	return x;
}

SEXP Rf_ScalarInteger(int v)
{
	SEXP x = Rf_allocVector(INTSXP, 1);
	INTEGER(x)[0] = v;
	return x;
}

SEXP Rf_ScalarLogical(int v)
{
	SEXP x = Rf_allocVector(LGLSXP, 1);
	LOGICAL(x)[0] = v == NA_LOGICAL ? NA_LOGICAL : (v != 0);
	return x;
}

SEXP Rf_ScalarReal(double v)
{
	SEXP x = Rf_allocVector(REALSXP, 1);
	REAL(x)[0] = v;
	return x;
}

SEXP Rf_ScalarString(SEXP v)
{
	SEXP x = Rf_allocVector(STRSXP, 1);
	((SEXP *) x->data)[0] = v;
	return x;
}

SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v)
{
	if (TYPEOF(x) != VECSXP || i < 0 || i >= XLENGTH(x))
		Rf_error("rshim: SET_VECTOR_ELT(): bad list or index");
	((SEXP *) x->data)[i] = v;
	return v;
}

void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v)
{
	if (TYPEOF(x) != STRSXP || i < 0 || i >= XLENGTH(x))
		Rf_error("rshim: SET_STRING_ELT(): bad vector or index");
	((SEXP *) x->data)[i] = v;
}

static int is_static_record(SEXP x)
{
	return x == NULL || x == R_NilValue ||
	       x == R_NaString || x == R_BlankString;
}

void rshim_release(SEXP x)
{
	if (is_static_record(x))
		return;
	if (TYPEOF(x) == EXTPTRSXP && x->finalizer != NULL)
		x->finalizer(x);
	if (x->owns_data)
		free(x->data);
	free(x);
	__atomic_sub_fetch(&live_objects, 1, __ATOMIC_RELAXED);
}

void rshim_release_tree(SEXP x)
{
	if (is_static_record(x))
		return;
	if (TYPEOF(x) == VECSXP || TYPEOF(x) == STRSXP) {
		for (R_xlen_t i = 0; i < XLENGTH(x); i++)
			rshim_release_tree(((SEXP *) x->data)[i]);
	}
	rshim_release_tree(x->dim);
	rshim_release_tree(x->names);
	rshim_release_tree(x->dimnames);
	rshim_release(x);
}

/* R_alloc() arena: everything is released when rshim_try_call() returns. */
typedef struct ralloc_block { struct ralloc_block *next; } ralloc_block;
static ralloc_block *ralloc_head = NULL;

char *R_alloc(size_t n, int size)
{
	size_t nbytes = n * (size_t) size;
	ralloc_block *b = (ralloc_block *) malloc(sizeof(ralloc_block) + 16 +
						 (nbytes ? nbytes : 1));
	if (b == NULL)
		Rf_error("rshim: R_alloc(): cannot allocate %zu bytes", nbytes);
	b->next = ralloc_head;
	ralloc_head = b;
	return (char *) b + 16;  /* sizeof(ralloc_block) <= 16: keep alignment */
}

static void ralloc_reset(void)
{
	while (ralloc_head != NULL) {
		ralloc_block *next = ralloc_head->next;
		free(ralloc_head);
		ralloc_head = next;
	}
}

/* ------------------------------------------------------------------ */
/* predicates, type names, attributes                                  */

int Rf_isVectorList(SEXP x) { return TYPEOF(x) == VECSXP; }

int Rf_isBlankString(const char *s)
{
	for (; *s; s++)
		if (*s != ' ' && *s != '\t' && *s != '\n' && *s != '\r')
			return 0;
	return 1;
}

static const struct { const char *name; SEXPTYPE type; } type_table[] = {
	{ "NULL", NILSXP }, { "logical", LGLSXP }, { "integer", INTSXP },
	{ "double", REALSXP }, { "complex", CPLXSXP },
	{ "character", STRSXP }, { "list", VECSXP }, { "raw", RAWSXP },
	{ "char", CHARSXP }, { "symbol", SYMSXP }, { "pairlist", LISTSXP },
	{ "numeric", REALSXP },
	{ NULL, 0 }
};

const char *Rf_type2char(SEXPTYPE t)
{
	for (int i = 0; type_table[i].name != NULL; i++)
		if (type_table[i].type == t)
			return type_table[i].name;
	return "unknown";
}

SEXPTYPE Rf_str2type(const char *s)
{
	for (int i = 0; type_table[i].name != NULL; i++)
		if (strcmp(type_table[i].name, s) == 0)
			return type_table[i].type;
	return (SEXPTYPE) -1;
}

SEXP Rf_getDim(SEXP x) { return x->dim; }
SEXP Rf_setDim(SEXP x, SEXP v) { x->dim = v; return x; }
SEXP Rf_getNames(SEXP x) { return x->names; }
SEXP Rf_setNames(SEXP x, SEXP v) { x->names = v; return x; }
SEXP Rf_getDimnames(SEXP x) { return x->dimnames; }
SEXP Rf_setDimnames(SEXP x, SEXP v) { x->dimnames = v; return x; }

/* ------------------------------------------------------------------ */
/* harness helpers                                                     */

SEXP rshim_svt_from_csc(int ncol, const int64_t *ptr, int *offs, void *vals,
			SEXPTYPE vals_type, const unsigned char *lacunar)
{
	if (ncol == 0 || ptr[ncol] == ptr[0])
		return R_NilValue;
	size_t vsz = vals != NULL ? elt_size(vals_type) : 0;
	SEXP svt = Rf_allocVector(VECSXP, ncol);
	for (int j = 0; j < ncol; j++) {
		int64_t start = ptr[j], nz = ptr[j + 1] - start;
		if (nz == 0)
			continue;  /* stays R_NilValue */
		SEXP leaf = Rf_allocVector(VECSXP, 2);
		int is_lacunar = vals == NULL ||
				 (lacunar != NULL && lacunar[j]);
		if (!is_lacunar)
			SET_VECTOR_ELT(leaf, 0, rshim_wrap_vector(vals_type, nz,
					(char *) vals + vsz * (size_t) start));
		SET_VECTOR_ELT(leaf, 1, rshim_wrap_vector(INTSXP, nz,
					offs + start));
		SET_VECTOR_ELT(svt, j, leaf);
	}
	return svt;
}

typedef SEXP (*fn0)(void);
typedef SEXP (*fn1)(SEXP);
typedef SEXP (*fn2)(SEXP, SEXP);
typedef SEXP (*fn3)(SEXP, SEXP, SEXP);
typedef SEXP (*fn4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*fn9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

SEXP rshim_try_call(void *fn, int nargs, SEXP *a, int *status)
{
	volatile SEXP ans = NULL;
	*status = 0;
	last_error[0] = '\0';
	if (setjmp(error_jmpbuf)) {
		error_jmpbuf_armed = 0;
		ralloc_reset();
		*status = 1;
		return NULL;
	}
	error_jmpbuf_armed = 1;
	switch (nargs) {
	    case 0: ans = ((fn0) fn)(); break;
	    case 1: ans = ((fn1) fn)(a[0]); break;
	    case 2: ans = ((fn2) fn)(a[0], a[1]); break;
	    case 3: ans = ((fn3) fn)(a[0], a[1], a[2]); break;
	    case 4: ans = ((fn4) fn)(a[0], a[1], a[2], a[3]); break;
	    case 5: ans = ((fn5) fn)(a[0], a[1], a[2], a[3], a[4]); break;
	    case 6: ans = ((fn6) fn)(a[0], a[1], a[2], a[3], a[4], a[5]);
		    break;
	    case 7: ans = ((fn7) fn)(a[0], a[1], a[2], a[3], a[4], a[5],
				     a[6]); break;
	    case 8: ans = ((fn8) fn)(a[0], a[1], a[2], a[3], a[4], a[5],
				     a[6], a[7]); break;
	    case 9: ans = ((fn9) fn)(a[0], a[1], a[2], a[3], a[4], a[5],
				     a[6], a[7], a[8]); break;
	    default:
		Rf_error("rshim_try_call(): unsupported arity %d", nargs);
	}
	error_jmpbuf_armed = 0;
	ralloc_reset();
	return ans;
}

/* ------------------------------------------------------------------ */
/* Rdynload                                                            */

int R_registerRoutines(DllInfo *info, const R_CMethodDef *const croutines,
		       const R_CallMethodDef *const call_routines,
		       const R_FortranMethodDef *const fortran_routines,
		       const R_ExternalMethodDef *const external_routines)
{
	(void) croutines; (void) fortran_routines; (void) external_routines;
	info->call_methods = call_routines;
	info->n_call_methods = 0;
	if (call_routines != NULL)
		while (call_routines[info->n_call_methods].name != NULL)
			info->n_call_methods++;
	return 1;
}

Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value)
{
	Rboolean old = info->use_dynamic_symbols ? TRUE : FALSE;
	info->use_dynamic_symbols = value;
	return old;
}

DL_FUNC rshim_lookup_call_routine(const DllInfo *info, const char *name,
				  int *nargs)
{
	for (int i = 0; i < info->n_call_methods; i++) {
		if (strcmp(info->call_methods[i].name, name) == 0) {
			if (nargs != NULL)
				*nargs = info->call_methods[i].numArgs;
			return info->call_methods[i].fun;
		}
	}
	return NULL;
}
