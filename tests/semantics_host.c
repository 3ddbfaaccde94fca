/* TEST INFRASTRUCTURE: host build of sparsearray_b200/csrc/svt_semantics.h
 * (the result-composition rules the CUDA kernels apply to their partials) so
 * the NA/NaN/zero-background logic can be checked on a machine without a GPU.
 * Built by tests/test_semantics_host.py with gcc. */
#include "../sparsearray_b200/csrc/svt_semantics.h"

/* partial: nz, n_na, n_nan, n_zero, sum, sum2, prod, vmin, vmax */
void sem_col_finalize(int opcode, int is_double, int narm, int64_t in_length,
		      double center, const double *partial, double *out_d,
		      int32_t *out_i, int *warn)
{
	SvtColPartial p;
	p.nz = (int64_t) partial[0];
	p.n_na = (int64_t) partial[1];
	p.n_nan = (int64_t) partial[2];
	p.n_zero = (int64_t) partial[3];
	p.sum = partial[4];
	p.sum2 = partial[5];
	p.prod = partial[6];
	p.vmin = partial[7];
	p.vmax = partial[8];
	SvtScalar r = svt_col_finalize(opcode, is_double, narm, in_length,
				       center, &p);
	*out_d = r.d;
	*out_i = r.i;
	*warn = r.warn;
}

double sem_col_mean(int is_double, int narm, int64_t in_length,
		    const double *partial)
{
	SvtColPartial p;
	svt_col_partial_init(&p);
	p.nz = (int64_t) partial[0];
	p.n_na = (int64_t) partial[1];
	p.n_nan = (int64_t) partial[2];
	p.sum = partial[4];
	return svt_col_mean(is_double, narm, in_length, &p);
}

void sem_row_finalize(int opcode, int is_double, int narm, int64_t nstrata,
		      int have_center, double center, const double *state6,
		      double *out_d, int32_t *out_i, int *warn)
{
	SvtScalar r = svt_row_finalize(opcode, is_double, narm, nstrata,
				       have_center, center, state6, 1);
	*out_d = r.d;
	*out_i = r.i;
	*warn = r.warn;
}

void sem_row_moments(int narm, int64_t nstrata, const double *state6,
		     double *mean, double *var)
{
	svt_row_moments(narm, nstrata, state6, 1, mean, var);
}

double sem_dot_finalize(int is_double, double s, int leaf_flag,
			int64_t hits_nonfinite, int n_nonfinite, int n_na)
{
	SvtDenseColInfo ci;
	ci.n_nonfinite = n_nonfinite;
	ci.n_na = n_na;
	return svt_dot_finalize(is_double, s, leaf_flag, hits_nonfinite, ci);
}

int sem_col_out_is_int(int opcode, int val_type)
{
	return svt_col_out_is_int(opcode, val_type);
}
