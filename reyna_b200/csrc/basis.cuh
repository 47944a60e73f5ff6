// Bounding-box tensor-Legendre basis in registers.
//
// Replaces the numba helpers of the reference (all under reyna/DGFEM/two_dimensional/_auxilliaries/):
//   tensor_LegendreP        polygonal_basis_utils.py:29-78   -> jacobi recurrence constants (LegConst) + leg1d
//   tensor_shift_leg        assembly_aux.py:32-66            -> leg1d (shift, clamp, h^-1/2 and h^-3/2 scaling)
//   tensor_tensor_leg       assembly_aux.py:69-91            -> eval_basis (values)
//   tensor_gradtensor_leg   assembly_aux.py:94-133           -> eval_basis (gradients)
// Basis index k enumerates (i,j), i+j<=P, row-major in (i,j) (Basis_index2D, polygonal_basis_utils.py:7-26);
// all loops over k are written as the nested (i,j) loops so that every register index is a compile-time constant.
#pragma once
#include "common.cuh"
#include <math.h>

namespace rb {

struct LegConst {
    double c0[2];                     // P_0 = 1/sqrt(gamma0),           alpha = 0 (values), 1 (derivatives)
    double c1[2];                     // P_1 = c1 * y
    double aold[2][RB_MAX_P + 1];     // a_old used by recurrence step j
    double ianew[2][RB_MAX_P + 1];    // 1 / a_new of recurrence step j
    double corr[RB_MAX_P + 2];        // sqrt(n (n+1)), n = 1..P  (assembly_aux.py:114)
};

// Host: the constants of the three-term recurrence, evaluated with the reference's formulas.
inline LegConst make_leg_const() {
    LegConst lc;
    memset(&lc, 0, sizeof(lc));
    for (int al = 0; al < 2; ++al) {
        const double a = (double)al;
        const double gamma0 = pow(2.0, 2.0 * a + 1.0) / (2.0 * a + 1.0) * (tgamma(a + 1.0) * tgamma(a + 1.0)) /
                              tgamma(2.0 * a + 1.0);
        lc.c0[al] = 1.0 / sqrt(gamma0);
        const double gamma1 = ((a + 1.0) * (a + 1.0) / (2.0 * a + 3.0)) * gamma0;
        lc.c1[al] = (a + 1.0) / sqrt(gamma1);
        double a_old = 1.0 / (a + 1.0) * sqrt((a + 1.0) * (a + 1.0) / (2.0 * a + 3.0));
        for (int j = 1; j <= RB_MAX_P; ++j) {
            const double h1 = 2.0 * (j + a);
            const double a_new = 1.0 / (j + a + 1.0) *
                                 sqrt((j + 1.0) * (j + 1.0 + 2.0 * a) * (j + 1.0 + a) * (j + 1.0 + a) /
                                      ((h1 + 1.0) * (h1 + 3.0)));
            lc.aold[al][j] = a_old;
            lc.ianew[al][j] = 1.0 / a_new;
            a_old = a_new;
        }
    }
    for (int n = 1; n <= RB_MAX_P + 1; ++n) lc.corr[n] = sqrt((n + 1.0) * n);
    return lc;
}

// Per-element scaling derived from the Cartesian bounding box [xmin,xmax,ymin,ymax].
struct __align__(16) ElemBox {
    double mx, my;       // mid point
    double ihx, ihy;     // 1 / half-extent
    double s05x, s05y;   // h^-1/2
    double s15x, s15y;   // h^-3/2
};

__device__ __forceinline__ ElemBox make_box(const double* __restrict__ bbox, int e) {
    const double4 b = *reinterpret_cast<const double4*>(bbox + 4 * (size_t)e);
    ElemBox bx;
    const double hx = 0.5 * (b.y - b.x), hy = 0.5 * (b.w - b.z);
    bx.mx = 0.5 * (b.y + b.x);
    bx.my = 0.5 * (b.w + b.z);
    bx.ihx = 1.0 / hx;
    bx.ihy = 1.0 / hy;
    bx.s05x = rsqrt(hx);
    bx.s05y = rsqrt(hy);
    bx.s15x = bx.s05x * bx.ihx;
    bx.s15y = bx.s05y * bx.ihy;
    return bx;
}

constexpr double kClamp = 1.0 - 2.220446049250313e-16;   // assembly_aux.py:50-54

template <int P, bool GRAD>
__device__ __forceinline__ void leg1d(double x, double m, double ih, double s05, double s15, const LegConst& lc,
                                      double (&val)[P + 1], double (&der)[P + 1]) {
    double y = (x - m) * ih;
    if (fabs(y) > 1.0) y = copysign(kClamp, y);
    double v[P + 1];
    v[0] = lc.c0[0];
    if (P >= 1) v[1] = lc.c1[0] * y;
#pragma unroll
    for (int j = 1; j < P; ++j) v[j + 1] = (y * v[j] - lc.aold[0][j] * v[j - 1]) * lc.ianew[0][j];
#pragma unroll
    for (int n = 0; n <= P; ++n) val[n] = s05 * v[n];
    if (GRAD) {
        der[0] = 0.0;
        if (P >= 1) {
            double w[P];
            w[0] = lc.c0[1];
            if (P >= 2) w[1] = lc.c1[1] * y;
#pragma unroll
            for (int j = 1; j < P - 1; ++j) w[j + 1] = (y * w[j] - lc.aold[1][j] * w[j - 1]) * lc.ianew[1][j];
#pragma unroll
            for (int n = 1; n <= P; ++n) der[n] = (s15 * lc.corr[n]) * w[n - 1];
        }
    }
}

template <int P>
struct BasisDim {
    static constexpr int D = (P + 1) * (P + 2) / 2;
};

// (i,j) of basis index k, branch-free so that it folds to a constant inside unrolled loops (rows start at 0, P+1, 2P+1, 3P).
template <int P>
__host__ __device__ constexpr int basis_i(int k) {
    static_assert(P <= 3, "basis_i: row starts are written out for P <= 3");
    return (k >= P + 1) + (k >= 2 * P + 1) + (k >= 3 * P);
}
template <int P>
__host__ __device__ constexpr int basis_j(int k) {
    const int i = basis_i<P>(k);
    return k - (i * (P + 1) - i * (i - 1) / 2);
}
// d/dx phi_k vanishes identically for i = 0 and d/dy phi_k for j = 0 (P_0' = 0, assembly_aux.py:112-133 multiplies by that
// exact zero): the kernels skip those products and the FMAs that would add their exact zeros.
template <int P>
__host__ __device__ constexpr bool gx_nz(int k) { return basis_i<P>(k) > 0; }
template <int P>
__host__ __device__ constexpr bool gy_nz(int k) { return basis_j<P>(k) > 0; }

// phi[k], and (when GRAD) d/dx, d/dy of the D basis functions of the element with box `bx` at the point (x,y).
// gx[k] / gy[k] of the identically-zero entries are set to a literal 0 (dead when the caller honours gx_nz / gy_nz).
template <int P, bool GRAD>
__device__ __forceinline__ void eval_basis(double x, double y, const ElemBox& bx, const LegConst& lc,
                                           double (&phi)[BasisDim<P>::D], double (&gx)[BasisDim<P>::D],
                                           double (&gy)[BasisDim<P>::D]) {
    double vx[P + 1], dx[P + 1], vy[P + 1], dy[P + 1];
    leg1d<P, GRAD>(x, bx.mx, bx.ihx, bx.s05x, bx.s15x, lc, vx, dx);
    leg1d<P, GRAD>(y, bx.my, bx.ihy, bx.s05y, bx.s15y, lc, vy, dy);
    int k = 0;
#pragma unroll
    for (int i = 0; i <= P; ++i) {
#pragma unroll
        for (int j = 0; j <= P - i; ++j) {
            phi[k] = vx[i] * vy[j];
            if (GRAD) {
                gx[k] = i > 0 ? dx[i] * vy[j] : 0.0;
                gy[k] = j > 0 ? vx[i] * dy[j] : 0.0;
            }
            ++k;
        }
    }
}

}  // namespace rb
