// 1-D basis-function recurrences shared by the K1 sampler (kernels_exact.cu) and the field evaluator (fields.cu).
// Evaluate N_n, N'_n, T_n, T'_n for n = 0..nmax at one point x, in the reference's operation order.
// INCLUDE ONLY FROM TRANSLATION UNITS COMPILED WITH -fmad=false.  Citations: /root/reference/src/fem_domain/basis/hierarchical_basis_fns.rs
#pragma once
#include <stdint.h>

#include "../../include/fem2d.h"

namespace fem2d {

// HierMaxOrtho constants, verbatim incl. apparent typos (:206-225).
static __constant__ double c_euc_norm[12] = {0.968246, 2.561738, 0.838525, 4.248161, 0.816397, 5.882766, 0.808509, 1.0, 1.0, 1.0, 1.0, 1.0};
static __constant__ int c_q_num[12][14] = {
    {-1, 0, 1}, {0, -3, 0, 3}, {-1, 0, -5, 0, 6}, {0, -3, 0, -7, 0, 10}, {-1, 0, -5, 0, -9, 0, 15},
    {0, -3, 0, -7, 0, -11, 0, 21}, {-1, 0, -5, 0, -9, 0, -13, 0, 28}, {0, -3, 0, -7, 0, -11, 0, -15, 0, 36},
    {-1, 0, -5, 0, -9, 0, -13, 0, -17, 0, 40}, {0, -3, 0, -7, 0, -11, 0, -15, 0, -19, 0, 55},
    {-1, 0, -5, 0, -9, 0, -13, 0, -17, 0, -21, 0, 66}, {0, -3, 0, -7, 0, -11, 0, -15, 0, -19, 0, -23, 0, 72}};
static __constant__ int c_q_den[12] = {1, 3, 6, 10, 15, 21, 28, 36, 40, 55, 66, 72};

// `store(arr, n, value)` receives arr: 0 N, 1 N', 2 T, 3 T'.
template <class Store>
__device__ __forceinline__ void basis_at_point(int basis, uint32_t nmax, double x, Store store) {
    if (basis == FEM2D_BASIS_HIER_POLY) {   // HierPoly::new_without_d2 (:102-162)
        double pw_prev = 1.0;
        for (uint32_t n = 0; n <= nmax; n++) {
            if (n == 0) { store(2, 0, 1.0 - x); store(3, 0, -1.0); store(0, 0, 1.0); store(1, 0, 0.0); pw_prev = 1.0; }
            else if (n == 1) { store(2, 1, 1.0 + x); store(3, 1, 1.0); store(0, 1, x); store(1, 1, 1.0); pw_prev = x; }
            else {
                const double pw = pw_prev * x;
                const double d1 = (double)n * pw_prev;
                store(0, n, pw); store(1, n, d1);
                if (n % 2 == 0) { store(2, n, pw - 1.0); store(3, n, d1); }
                else { store(2, n, pw - x); store(3, n, d1 - 1.0); }
                pw_prev = pw;
            }
        }
    } else {   // HierMaxOrtho: LegendrePoly (:425-463) + QFunction (:316-352, :593-621)
        double L[21], Ld[21];
        for (uint32_t i = 0; i <= nmax; i++) {
            const double i_f = (double)i;
            if (i == 0) { L[0] = 1.0; Ld[0] = 0.0; }
            else if (i == 1) { L[1] = x; Ld[1] = 1.0; }
            else {
                L[i] = ((2.0 * i_f - 1.0) * x * L[i - 1] - (i_f - 1.0) * L[i - 2]) / i_f;
                Ld[i] = i_f * L[i - 1] + x * Ld[i - 1];
            }
            store(0, i, L[i]); store(1, i, Ld[i]);
        }
        for (uint32_t i = 0; i <= nmax; i++) {
            if (i == 0) { store(2, 0, 1.0 - x); store(3, 0, -1.0); }
            else if (i == 1) { store(2, 1, 1.0 + x); store(3, 1, 1.0); }
            else {
                double sv = 0.0, sp = 0.0;
                for (uint32_t k = 0; k <= i; k++) {
                    const double w = ((double)c_q_num[i - 2][k]) / ((double)c_q_den[i - 2]);
                    sv += w * L[k];
                    sp += w * Ld[k];
                }
                store(2, i, sv * c_euc_norm[i - 2]);
                store(3, i, sp * c_euc_norm[i - 2]);
            }
        }
    }
}

}  // namespace fem2d
