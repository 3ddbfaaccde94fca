/* TEST INFRASTRUCTURE -- stand-in for the un-vendored S4Vectors C interface
 * (LinkingTo: S4Vectors in the reference's DESCRIPTION; no version pinned,
 * the headers are not in /root/reference), limited to what
 * src/rowsum_methods.c uses.  The four routines restate the published
 * algorithm of S4Vectors/src/safe_arithm.c ("safe" int arithmetic: NA in ->
 * NA out; a result outside [-INT_MAX, INT_MAX] -> NA and a sticky overflow
 * flag).  Parity of the integer-overflow corner of rowsum()/colsum() is
 * pinned on this restatement, not on a build of S4Vectors itself.
 *
 * GET_SLOT()/install() only occur in the dgCMatrix entry points, which the
 * oracle never calls (they take S4 objects the shim does not model). */
#ifndef S4VECTORS_INTERFACE_STUB_H
#define S4VECTORS_INTERFACE_STUB_H

#include <Rdefines.h>
#include <limits.h>

static int s4v_ovflow_flag;

static inline void reset_ovflow_flag(void) { s4v_ovflow_flag = 0; }
static inline int get_ovflow_flag(void) { return s4v_ovflow_flag; }

static inline int safe_int_add(int x, int y)
{
	if (x == NA_INTEGER || y == NA_INTEGER)
		return NA_INTEGER;
	if ((y > 0 && x > INT_MAX - y) || (y < 0 && x < -INT_MAX - y)) {
		s4v_ovflow_flag = 1;
		return NA_INTEGER;
	}
	return x + y;
}

static inline int safe_int_mult(int x, int y)
{
	if (x == NA_INTEGER || y == NA_INTEGER)
		return NA_INTEGER;
	long long z = (long long) x * (long long) y;
	if (z > INT_MAX || z < -INT_MAX) {
		s4v_ovflow_flag = 1;
		return NA_INTEGER;
	}
	return (int) z;
}

#define GET_SLOT(x, what) \
	(Rf_error("S4 slots are not modelled by the R shim"), R_NilValue)
#define install(s) R_NilValue

#endif  /* S4VECTORS_INTERFACE_STUB_H */
