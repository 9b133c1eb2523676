/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU oracle for the hot path of DiffPointRasterisation.jl: a plain-C restatement of the reference's
 * algorithm for `raster` (src/raster.jl:5-108) and `raster_pullback!` (src/raster_pullback.jl:2-160,
 * corner order from src/util.jl:7-27).  The reference is Julia and cannot run in this image (no Julia),
 * so this restatement is the checker; it is pinned against the reference's own known-answer tests
 * (src/raster.jl:143-309, README.md:41-68, README.md:99-183, src/util.jl:29-46) by tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (libdpr.so) never links or calls it.
 *
 * Build (see oracle/Makefile): gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC.  -ffp-contract=off matters:
 * the reference never fuses multiply-add, and the cell a point lands in depends on the last bit of `coord`.
 *
 * Two accumulation modes:
 *   faithful        (f64_accumulate = 0): every sum sequential in the element type, like the reference loops.
 *   f64-accumulate  (f64_accumulate = 1): stencil (coord, ref, deltas) in the element type, sums in double;
 *                                         the yardstick for the Float32 parity gate.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define SUF _f32
#define CEIL ceilf
#include "dpr_oracle_impl.inc"
#undef T
#undef SUF
#undef CEIL

#define T double
#define SUF _f64
#define CEIL ceil
#include "dpr_oracle_impl.inc"
#undef T
#undef SUF
#undef CEIL

/* voxel_shifts(Val(N)), src/util.jl:26-27: corner k (0-based) has shift_d = (k >> d) & 1, dimension 0 fastest. */
void dpro_voxel_shifts(int n, int64_t* out /* (2^n, n) row-major */) {
    for (int k = 0; k < (1 << n); ++k)
        for (int d = 0; d < n; ++d) out[k * n + d] = (k >> d) & 1;
}

int dpro_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
