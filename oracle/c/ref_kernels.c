/* CPU ORACLE, C / OpenMP part (TEST INFRASTRUCTURE - never linked into the product library).
 *
 * Restates the pystencils-generated loop nests of sopht/numeric/eulerian_grid_ops/stencil_ops_{2d,3d} as plain C with
 * one OpenMP loop nest per reference kernel (what pystencils 1.4 emits for `cpu_openmp`), so that bench.py's CPU arm
 * times the reference's own dataflow instead of numpy slice arithmetic. Every function cites the reference file:line it
 * follows; the Python side (oracle/cstencils.py) composes them in the reference's launch order and is pinned to the
 * same golden vectors as the numpy oracle (tests/test_oracle_golden.py).
 *
 * Build: oracle/Makefile  ->  oracle/_build/libsopht_ref_kernels.so
 */
#include <stdint.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#include "ref_kernels_body.inc"
#undef REAL
#undef FN

#define REAL double
#define FN(name) CAT(name, _f64)
#include "ref_kernels_body.inc"
#undef REAL
#undef FN

int sopht_ref_kernels_abi(void) { return 1; }
