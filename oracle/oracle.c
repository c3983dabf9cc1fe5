/* CPU oracle, C restatement of the reference's hot-path kernels (see oracle_impl.h).
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's CPU baseline.
 * Built by oracle/Makefile into oracle/liboracle.so with -ffp-contract=off (the Julia CPU path
 * does not contract a*b+c into FMA).  Parity pinning: see oracle_np.py header. */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD 8   /* max array rank */
#define ORC_MAXW 16  /* max degree+1 */

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#define T float
#define FN(name) name##_f32
#include "oracle_impl.h"
#undef T
#undef FN

#define T double
#define FN(name) name##_f64
#include "oracle_impl.h"
#undef T
#undef FN
