#include "rglue_common.h"

#include <limits.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#undef match
#include <omp.h>
#endif

SEXPTYPE rglue_get_and_check_Rtype(SEXP type, const char *fun,
				   const char *argname)
{
	SEXPTYPE Rtype = 0;
	if (IS_CHARACTER(type) && LENGTH(type) == 1 &&
	    STRING_ELT(type, 0) != NA_STRING)
	{
		const char *s = CHAR(STRING_ELT(type, 0));
		SEXPTYPE t = str2type(s);
		/* the vector types an SVT can hold
		   (_get_Rtype_from_Rstring(), src/Rvector_utils.c:27-48) */
		if (t == LGLSXP || t == INTSXP || t == REALSXP ||
		    t == CPLXSXP || t == STRSXP || t == VECSXP || t == RAWSXP)
			Rtype = t;
	}
	if (Rtype == 0)
		error("SparseArray internal error in %s():\n"
		      "    invalid '%s' value", fun, argname);
	return Rtype;
}

int rglue_get_and_check_na_background(SEXP na_background, const char *fun,
				      const char *argname)
{
	if (!(IS_LOGICAL(na_background) && LENGTH(na_background) == 1))
		error("SparseArray internal error in %s():\n"
		      "    '%s' must be TRUE or FALSE", fun, argname);
	return LOGICAL(na_background)[0] != 0;
}

int rglue_get_summarize_opcode(SEXP op, SEXPTYPE Rtype)
{
	static const struct { const char *name; int opcode; int numeric_only; }
	ops[] = {
		{ "anyNA", SVTGPU_OP_ANYNA, 0 },
		{ "countNAs", SVTGPU_OP_COUNTNAS, 0 },
		{ "min", SVTGPU_OP_MIN, 1 }, { "max", SVTGPU_OP_MAX, 1 },
		{ "range", SVTGPU_OP_RANGE, 1 }, { "sum", SVTGPU_OP_SUM, 1 },
		{ "prod", SVTGPU_OP_PROD, 1 }, { "mean", SVTGPU_OP_MEAN, 1 },
		{ "centered_X2_sum", SVTGPU_OP_CENTERED_X2_SUM, 1 },
		{ "sum_X_X2", SVTGPU_OP_SUM_X_X2, 1 },
		{ "var1", SVTGPU_OP_VAR1, 1 }, { "var2", SVTGPU_OP_VAR2, 1 },
		{ "sd1", SVTGPU_OP_SD1, 1 }, { "sd2", SVTGPU_OP_SD2, 1 },
		{ "any", SVTGPU_OP_ANY, 2 }, { "all", SVTGPU_OP_ALL, 2 },
	};
	if (!(IS_CHARACTER(op) && LENGTH(op) == 1))
		error("'op' must be a single string");
	op = STRING_ELT(op, 0);
	if (op == NA_STRING)
		error("'op' cannot be NA");
	const char *s = CHAR(op);
	if (Rtype != LGLSXP && Rtype != INTSXP && Rtype != REALSXP &&
	    Rtype != CPLXSXP && Rtype != STRSXP)
		error("%s() does not support SparseArray objects "
		      "of type() \"%s\"", s, type2char(Rtype));
	for (size_t i = 0; i < sizeof(ops) / sizeof(ops[0]); i++) {
		if (strcmp(s, ops[i].name) != 0)
			continue;
		int ok = 1;
		if (ops[i].numeric_only >= 1)
			ok = Rtype == LGLSXP || Rtype == INTSXP ||
			     Rtype == REALSXP;
		if (ops[i].numeric_only == 2)
			ok = Rtype == LGLSXP || Rtype == INTSXP;
		if (!ok)
			error("%s() does not support SparseArray objects "
			      "of type() \"%s\"", s, type2char(Rtype));
		return ops[i].opcode;
	}
	error("'op' must be one of: "
	      "\"anyNA\", \"countNAs\", \"any\", \"all\",\n"
	      "                       \"min\", \"max\", "
	      "\"range\", \"sum\", \"prod\", \"mean\",\n"
	      "                       \"centered_X2_sum\", \"sum_X_X2\",\n"
	      "                       \"var1\", \"var2\", \"sd1\", \"sd2\"");
	return 0;
}

void rglue_fail(int rc, const char *fun)
{
	if (rc == SVT_FLATTEN_BAD_OFFSETS)
		error("SparseArray internal error in svt_index_leaves():\n"
		      "    invalid SVT leaf ('nzoffs' must be strictly ascending "
		      "and inside the first dimension)");
	if (rc == SVTGPU_ERR_ARG || rc == SVTGPU_ERR_UNSUPPORTED)
		error("%s", svtgpu_last_error());
	error("SparseArray GPU path: %s() failed (status %d):\n    %s",
	      fun, rc, svtgpu_last_error());
}

double rglue_now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void rglue_trace(const char *fun, double t_index, double t_upload,
		 double t_op, double t_free)
{
	const char *v = getenv("SVTGPU_TRACE");
	if (v == NULL || v[0] == '\0' || v[0] == '0')
		return;
	fprintf(stderr, "[svtgpu] %s: index %.1f ms, flatten+upload %.1f ms, "
		"op %.1f ms, free %.1f ms\n", fun, t_index, t_upload, t_op,
		t_free);
}

static double last_timings[7];

void rglue_record_timings(const svtgpu_matrix *m, double flatten_ms)
{
	svtgpu_timings t;
	memset(&t, 0, sizeof(t));
	if (m != NULL)
		svtgpu_matrix_timings(m, &t);
	last_timings[0] = flatten_ms;
	last_timings[1] = t.h2d_ms;
	last_timings[2] = t.kernel_ms;
	last_timings[3] = t.d2h_ms;
	last_timings[4] = t.h2d_bytes;
	last_timings[5] = t.d2h_bytes;
	last_timings[6] = (double) t.launches;
}

/* c(flatten_ms, h2d_ms, kernel_ms, d2h_ms, h2d_bytes, d2h_bytes, launches)
   of the most recent .Call that ran on the GPU */
SEXP C_svtgpu_last_timings(void)
{
	SEXP ans = PROTECT(NEW_NUMERIC(7));
	memcpy(REAL(ans), last_timings, sizeof(last_timings));
	UNPROTECT(1);
	return ans;
}

/* ---- resident handles ---- */

static void resident_finalizer(SEXP handle)
{
	svtgpu_matrix *m = (svtgpu_matrix *) R_ExternalPtrAddr(handle);
	if (m != NULL) {
		svtgpu_matrix_free(m);
		R_ClearExternalPtr(handle);
	}
}

/* ---- device cache of the most recent SVT (opt-in) ----
 *
 * The R methods call the same SVT several times in a row: rowVars() is
 * C_rowStats_SVT x 3 (countNAs, sum, centered_X2_sum), colMeans() after
 * colSums() ...  -- and every stateless call flattens and uploads the same
 * leaves again (the host reads them at memory speed: ~150 ms per 2.3e9
 * nonzeros, whatever the GPU does).  With the cache on (options(
 * SparseArray.gpu.cache = TRUE) -> C_svtgpu_set_cache, or SVTGPU_CACHE=1) the
 * device CSC of the last SVT stays in HBM and is reused when the next call
 * presents the same object.  "The same" is decided on every call by a
 * fingerprint over ALL leaves -- address of the SVT list, dim, type, and per
 * leaf the addresses and length of nzoffs / nzvals plus their first and last
 * elements -- O(number of leaves), no payload pass.
 *
 * Caveat (why it is opt-in): R objects are immutable from R code, so a leaf
 * whose address, length and end elements are unchanged is unchanged, but C
 * code that modifies a leaf in place, or a new SVT that the allocator places
 * at exactly the addresses of a garbage-collected one with the same lengths
 * and end elements, would be served stale data.  C_svtgpu_set_cache(FALSE)
 * (or a call on a different SVT) releases the cached matrix. */
static struct {
	int enabled;            /* -1: not decided yet (environment) */
	svtgpu_matrix *m;
	uint64_t fp;
	int64_t nrow, nleaf, nnz;
	int Rtype;
	int64_t hits, misses;
} g_cache = { -1, NULL, 0, 0, 0, 0, 0, 0, 0 };

static int cache_enabled(void)
{
	if (g_cache.enabled < 0) {
		const char *v = getenv("SVTGPU_CACHE");
		g_cache.enabled = v != NULL && v[0] != '\0' && v[0] != '0';
	}
	return g_cache.enabled;
}

static void cache_drop(void)
{
	if (g_cache.m != NULL)
		svtgpu_matrix_free(g_cache.m);
	g_cache.m = NULL;
	g_cache.fp = 0;
}

static inline uint64_t fp_mix(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/* one leaf's share of the fingerprint: its place, the addresses and length
 * of nzoffs / nzvals, their first and last elements */
static inline uint64_t fp_leaf(int64_t l, int64_t n, const int *o,
			       const char *v, size_t vsz)
{
	uint64_t z = (uint64_t) (uintptr_t) o * 0x9E3779B97F4A7C15ull +
		     (uint64_t) (uintptr_t) v + ((uint64_t) n << 40) +
		     (uint64_t) l;
	z = fp_mix(z) ^ (((uint64_t) (uint32_t) o[0] << 32) |
			 (uint32_t) o[n - 1]);
	if (v != NULL) {
		uint64_t a = 0, b = 0;
		memcpy(&a, v, vsz);
		memcpy(&b, v + vsz * (size_t) (n - 1), vsz);
		z = fp_mix(z + a) ^ b;
	}
	return fp_mix(z);
}

static uint64_t fp_head(SEXP x_SVT, const int *dim, int ndim, SEXPTYPE Rtype)
{
	uint64_t h = fp_mix((uint64_t) (uintptr_t) x_SVT) ^
		     fp_mix(0x5BD1E995u + (uint64_t) Rtype);
	for (int along = 0; along < ndim; along++)
		h = fp_mix(h + (uint64_t) dim[along] * 0x9E3779B97F4A7C15ull);
	return h;
}

static uint64_t svt_fingerprint(SEXP x_SVT, const int *dim, int ndim,
				SEXPTYPE Rtype, const svt_leaf_index *ix)
{
	const uint64_t h = fp_head(x_SVT, dim, ndim, Rtype);
	const size_t vsz = Rtype == REALSXP ? 8 : 4;
	uint64_t acc = 0;
	#pragma omp parallel for schedule(static) reduction(+:acc)
	for (int64_t l = 0; l < ix->nleaf; l++) {
		const int64_t n = ix->leaf_ptr[l + 1] - ix->leaf_ptr[l];
		if (n == 0)
			continue;
		acc += fp_leaf(l, n, ix->offs[l], (const char *) ix->vals[l],
			       vsz);
	}
	return fp_mix(h ^ acc) | 1u;     /* never 0 */
}

/* The same fingerprint straight from the leaf list of a MATRIX, without
 * building the leaf index (three 8-byte arrays per leaf + a prefix sum): what
 * a call that is going to hit the cache needs.  A leaf costs a chain of three
 * dependent cache misses (leaf -> its two vectors -> their last elements), so
 * each level is prefetched SVTGPU_FP_AHEAD leaves ahead (default 4; measured
 * on the 16-core GPU hosts at 1e6 leaves: 12.8 ms without, 7.1 ms at 4, 7.6 at
 * 8, 9.5 from 16 on; index + fingerprint took 13-16 ms).  Returns 0 when a leaf
 * does not look like one (the caller then takes the full path, which raises
 * the proper error); *nnz receives the number of nonzeros. */
static int fp_ahead(void)
{
	static int v = -1;
	if (v < 0) {
		const char *e = getenv("SVTGPU_FP_AHEAD");
		v = e != NULL ? atoi(e) : 4;
		if (v < 0) v = 0;
	}
	return v;
}
static uint64_t svt_fingerprint_matrix(SEXP x_SVT, const int *dim,
				       SEXPTYPE Rtype, int64_t *nnz)
{
	*nnz = 0;
	if (!isVectorList(x_SVT) || LENGTH(x_SVT) != dim[1])
		return 0;
	const uint64_t h = fp_head(x_SVT, dim, 2, Rtype);
	const size_t vsz = Rtype == REALSXP ? 8 : 4;
	const int64_t nleaf = dim[1];
	uint64_t acc = 0;
	int64_t total = 0;
	int bad = 0;
	#pragma omp parallel reduction(+:acc, total) reduction(max:bad)
	{
		int64_t lo = 0, hi = nleaf;
#ifdef _OPENMP
		const int nt = omp_get_num_threads(), me = omp_get_thread_num();
		lo = nleaf * me / nt;
		hi = nleaf * (me + 1) / nt;
#endif
		const int FP_AHEAD = fp_ahead();
		for (int64_t l = lo; l < hi; l++) {
			if (FP_AHEAD > 0 && l + 3 * FP_AHEAD < hi)
				__builtin_prefetch(VECTOR_ELT(x_SVT,
							      l + 3 * FP_AHEAD));
			if (FP_AHEAD > 0 && l + 2 * FP_AHEAD < hi) {
				SEXP lf = VECTOR_ELT(x_SVT, l + 2 * FP_AHEAD);
				if (lf != R_NilValue && isVectorList(lf) &&
				    LENGTH(lf) >= 2) {
					__builtin_prefetch(VECTOR_ELT(lf, 0));
					__builtin_prefetch(VECTOR_ELT(lf, 1));
				}
			}
			if (FP_AHEAD > 0 && l + FP_AHEAD < hi) {
				SEXP lf = VECTOR_ELT(x_SVT, l + FP_AHEAD);
				if (lf != R_NilValue && isVectorList(lf) &&
				    LENGTH(lf) >= 2) {
					SEXP pv = VECTOR_ELT(lf, 0);
					SEXP po = VECTOR_ELT(lf, 1);
					if (IS_INTEGER(po) && XLENGTH(po) > 0) {
						const int *o = INTEGER(po);
						__builtin_prefetch(o);
						__builtin_prefetch(o + XLENGTH(po) - 1);
					}
					if (pv != R_NilValue &&
					    TYPEOF(pv) == Rtype && XLENGTH(pv) > 0) {
						const char *v = (const char *)
							DATAPTR(pv);
						__builtin_prefetch(v);
						__builtin_prefetch(v + vsz *
							(size_t) (XLENGTH(pv) - 1));
					}
				}
			}
			SEXP leaf = VECTOR_ELT(x_SVT, l);
			if (leaf == R_NilValue)
				continue;
			if (!isVectorList(leaf) || LENGTH(leaf) < 2) {
				bad = 1;
				continue;
			}
			SEXP nzvals = VECTOR_ELT(leaf, 0);
			SEXP nzoffs = VECTOR_ELT(leaf, 1);
			if (!IS_INTEGER(nzoffs)) {
				bad = 1;
				continue;
			}
			const R_xlen_t n = XLENGTH(nzoffs);
			if (n == 0 || n > INT_MAX) {
				bad = 1;
				continue;
			}
			const char *v = NULL;
			if (nzvals != R_NilValue) {
				if (TYPEOF(nzvals) != Rtype || XLENGTH(nzvals) != n) {
					bad = 1;
					continue;
				}
				v = (const char *) DATAPTR(nzvals);
			}
			acc += fp_leaf(l, (int64_t) n, INTEGER(nzoffs), v, vsz);
			total += n;
		}
	}
	if (bad)
		return 0;
	*nnz = total;
	return fp_mix(h ^ acc) | 1u;
}

/* --- .Call ENTRY POINT (extension) --- switch the device cache; returns the
 * previous setting.  Turning it off (or passing NA to just clear) releases
 * the cached matrix. */
SEXP C_svtgpu_set_cache(SEXP on)
{
	if (!(IS_LOGICAL(on) && LENGTH(on) == 1))
		error("'on' must be TRUE, FALSE or NA");
	int prev = cache_enabled();
	int v = LOGICAL(on)[0];
	if (v == NA_LOGICAL) {
		cache_drop();
	} else {
		g_cache.enabled = v != 0;
		if (!g_cache.enabled)
			cache_drop();
	}
	return ScalarLogical(prev);
}

/* c(hits, misses) since the library was loaded */
SEXP C_svtgpu_cache_stats(void)
{
	SEXP ans = PROTECT(NEW_NUMERIC(2));
	REAL(ans)[0] = (double) g_cache.hits;
	REAL(ans)[1] = (double) g_cache.misses;
	UNPROTECT(1);
	return ans;
}

void rglue_acquire2(SEXP x_SVT, const int *dim, int ndim, SEXPTYPE Rtype,
		    int want_offs, int want_vals, int may_share,
		    rglue_input *in)
{
	memset(in, 0, sizeof(*in));
	double t0 = rglue_now_ms();
	if (TYPEOF(x_SVT) == EXTPTRSXP) {
		svtgpu_matrix *m = (svtgpu_matrix *) R_ExternalPtrAddr(x_SVT);
		if (m == NULL)
			error("the device-resident SVT handle has been "
			      "released");
		int64_t nrow = 0, nleaf = 0, nnz = 0, want_nleaf = 1;
		int val_type = 0, flags = 0;
		if (svtgpu_matrix_info(m, &nrow, &nleaf, &nnz, &val_type,
				       &flags) != SVTGPU_OK)
			rglue_fail(SVTGPU_ERR_ARG, "svtgpu_matrix_info");
		for (int along = 1; along < ndim; along++)
			want_nleaf *= dim[along];
		if (ndim < 1 || nrow != dim[0] || nleaf != want_nleaf ||
		    val_type != (int) Rtype)
			error("the device-resident SVT handle does not match "
			      "the dimensions / type of the object");
		if (want_offs && nnz > 0 && !(flags & SVTGPU_HAS_OFFS))
			error("the device-resident SVT handle holds no row "
			      "offsets");
		in->m = m;
		in->resident = 1;
		in->shared = 1;
		in->t_ready = rglue_now_ms();
		return;
	}
	/* a call that presents the cached matrix again is recognised without
	   building the leaf index */
	if (may_share && cache_enabled() && g_cache.m != NULL && ndim == 2 &&
	    g_cache.nrow == dim[0] && g_cache.nleaf == dim[1] &&
	    g_cache.Rtype == (int) Rtype) {
		int64_t nnz = 0;
		const uint64_t fpm = svt_fingerprint_matrix(x_SVT, dim, Rtype,
							    &nnz);
		if (fpm != 0 && fpm == g_cache.fp && nnz == g_cache.nnz) {
			g_cache.hits++;
			in->m = g_cache.m;
			in->resident = 1;
			in->shared = 1;
			in->t_ready = rglue_now_ms();
			in->index_ms = in->t_ready - t0;
			const char *v = getenv("SVTGPU_TRACE");
			if (v != NULL && v[0] == '2')
				fprintf(stderr, "[svtgpu] cached matrix recognised "
					"in %.2f ms (%lld leaves)\n",
					in->index_ms, (long long) dim[1]);
			return;
		}
	}
	svt_leaf_index ix;
	svt_index_leaves(x_SVT, dim, ndim, Rtype, &ix);
	const int use_cache = may_share && cache_enabled() && ix.nnz > 0;
	uint64_t fp = 0;
	if (use_cache) {
		const double tf0 = rglue_now_ms();
		fp = svt_fingerprint(x_SVT, dim, ndim, Rtype, &ix);
		{
			const char *v = getenv("SVTGPU_TRACE");
			if (v != NULL && v[0] == '2') {
				fprintf(stderr, "[svtgpu] index %.2f ms, "
					"fingerprint %.2f ms (%lld leaves)\n",
					tf0 - t0, rglue_now_ms() - tf0,
					(long long) ix.nleaf);
				if (ndim == 2) {   /* self-check of the direct form */
					int64_t nnz2 = 0;
					const double td = rglue_now_ms();
					const uint64_t f2 = svt_fingerprint_matrix(
						x_SVT, dim, Rtype, &nnz2);
					fprintf(stderr, "[svtgpu] direct fingerprint "
						"%.2f ms, %s\n", rglue_now_ms() - td,
						f2 == fp && nnz2 == ix.nnz
						? "same value" : "DIFFERENT");
				}
			}
		}
		if (g_cache.m != NULL && g_cache.fp == fp &&
		    g_cache.nrow == ix.nrow && g_cache.nleaf == ix.nleaf &&
		    g_cache.nnz == ix.nnz && g_cache.Rtype == (int) Rtype) {
			g_cache.hits++;
			in->m = g_cache.m;
			in->resident = 1;      /* nothing moves for this call */
			in->shared = 1;
			in->t_ready = rglue_now_ms();
			in->index_ms = in->t_ready - t0;
			return;
		}
		g_cache.misses++;
		cache_drop();                  /* make room before uploading */
		want_offs = want_vals = 1;     /* serve every later operation */
	}
	double t1 = rglue_now_ms();
	int rc = svt_upload_leaves(&ix, Rtype, want_offs, want_vals, &in->m,
				   &in->flatten_ms);
	if (rc != SVTGPU_OK)
		rglue_fail(rc, "svt_upload_leaves");
	if (use_cache) {
		g_cache.m = in->m;
		g_cache.fp = fp;
		g_cache.nrow = ix.nrow;
		g_cache.nleaf = ix.nleaf;
		g_cache.nnz = ix.nnz;
		g_cache.Rtype = (int) Rtype;
		in->shared = 1;
	}
	in->t_ready = rglue_now_ms();
	in->index_ms = t1 - t0;
	in->upload_ms = in->t_ready - t1;
}

/* 1 when no leaf of x stores values (every leaf lacunar, or no leaf at all;
 * a handle: uploaded without a value array): countNAs / anyNA have nothing to
 * look at -- lacunar leaves stand for ones (the reference's
 * summarize_ones(), src/Rvector_summarization.c) -- and the caller answers
 * zeros without flattening or uploading anything.  A regular matrix is
 * recognised at its first leaf; only a leading lacunar leaf starts the full
 * (parallel, header-only) scan.  Anything that does not look like a leaf
 * answers 0: the normal path then raises the proper error. */
int rglue_svt_stores_no_values(SEXP x_SVT, const int *dim, int ndim)
{
	if (TYPEOF(x_SVT) == EXTPTRSXP) {
		svtgpu_matrix *m = (svtgpu_matrix *) R_ExternalPtrAddr(x_SVT);
		int64_t nrow = 0, nleaf = 0, nnz = 0;
		int val_type = 0, flags = 0;
		if (m == NULL || svtgpu_matrix_info(m, &nrow, &nleaf, &nnz,
						    &val_type, &flags) != SVTGPU_OK)
			return 0;
		return !(flags & SVTGPU_HAS_VALS);
	}
	if (x_SVT == R_NilValue)
		return 1;
	if (ndim != 2 || !isVectorList(x_SVT) || LENGTH(x_SVT) != dim[1])
		return 0;
	const int64_t nleaf = dim[1];
	int64_t first = 0;
	while (first < nleaf && VECTOR_ELT(x_SVT, first) == R_NilValue)
		first++;
	if (first == nleaf)
		return 1;
	{
		SEXP leaf = VECTOR_ELT(x_SVT, first);
		if (!isVectorList(leaf) || LENGTH(leaf) < 2 ||
		    VECTOR_ELT(leaf, 0) != R_NilValue)
			return 0;
	}
	int regular = 0;
	#pragma omp parallel for schedule(static) reduction(max:regular)
	for (int64_t l = first + 1; l < nleaf; l++) {
		SEXP leaf = VECTOR_ELT(x_SVT, l);
		if (leaf == R_NilValue)
			continue;
		if (!isVectorList(leaf) || LENGTH(leaf) < 2 ||
		    VECTOR_ELT(leaf, 0) != R_NilValue)
			regular = 1;
	}
	return !regular;
}

void rglue_acquire(SEXP x_SVT, const int *dim, int ndim, SEXPTYPE Rtype,
		   int want_offs, int want_vals, rglue_input *in)
{
	rglue_acquire2(x_SVT, dim, ndim, Rtype, want_offs, want_vals, 1, in);
}

void rglue_done(rglue_input *in, const char *fun)
{
	double t3 = rglue_now_ms();
	rglue_record_timings(in->m, in->flatten_ms);
	if (in->resident) {
		/* nothing moved to the device for this call */
		last_timings[0] = last_timings[1] = last_timings[4] = 0.0;
	}
	if (!in->shared)
		svtgpu_matrix_free(in->m);
	in->m = NULL;
	rglue_trace(fun, in->index_ms, in->upload_ms, t3 - in->t_ready,
		    rglue_now_ms() - t3);
}

/* --- .Call ENTRY POINT (extension) ---
 * Flatten + upload an SVT once; the result stands in for 'x_SVT' in every
 * entry point of the GPU path until it is garbage collected or released. */
SEXP C_svtgpu_resident_SVT(SEXP x_dim, SEXP x_type, SEXP x_SVT)
{
	SEXPTYPE Rtype = rglue_get_and_check_Rtype(x_type,
				"C_svtgpu_resident_SVT", "x_type");
	if (Rtype != LGLSXP && Rtype != INTSXP && Rtype != REALSXP)
		error("SparseArray objects of type() \"%s\" are not "
		      "supported by the SparseArray GPU path",
		      type2char(Rtype));
	if (!IS_INTEGER(x_dim) || LENGTH(x_dim) < 1)
		error("'x_dim' must be an integer vector of length >= 1");
	if (TYPEOF(x_SVT) == EXTPTRSXP)
		return x_SVT;
	/* the handle owns its matrix: never the cached one (the cache would
	   free or reuse it behind the handle's back) */
	rglue_input in;
	rglue_acquire2(x_SVT, INTEGER(x_dim), LENGTH(x_dim), Rtype, 1, 1, 0,
		       &in);
	int rc = svtgpu_matrix_finish_upload(in.m);
	rglue_record_timings(in.m, in.flatten_ms);
	rglue_trace("C_svtgpu_resident_SVT", in.index_ms, in.upload_ms, 0.0,
		    0.0);
	if (rc != SVTGPU_OK) {
		svtgpu_matrix_free(in.m);
		rglue_fail(rc, "svtgpu_matrix_finish_upload");
	}
	return rglue_make_handle(in.m);
}

/* the external pointer that stands for a device matrix in 'x_SVT' arguments;
   its finalizer frees the device memory */
SEXP rglue_make_handle(svtgpu_matrix *m)
{
	SEXP handle = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
	R_RegisterCFinalizerEx(handle, resident_finalizer, TRUE);
	UNPROTECT(1);
	return handle;
}

/* --- .Call ENTRY POINT (extension) --- free the device memory now */
SEXP C_svtgpu_release(SEXP handle)
{
	if (TYPEOF(handle) != EXTPTRSXP)
		error("'handle' must be a device-resident SVT handle");
	resident_finalizer(handle);
	return R_NilValue;
}
