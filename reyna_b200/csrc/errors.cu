// L2 / DG / H1 error norms of a DG solution, batched on the device.
//
// Replaces DGFEM.errors (reyna/DGFEM/two_dimensional/main.py:476-615) and its four per-item Python kernels
//   error_element    _auxilliaries/error_assembly/full_assembly.py:9-96
//   error_interface  _auxilliaries/error_assembly/full_assembly.py:99-187
//   error_d_face     _auxilliaries/error_assembly/boundary_assembly.py:8-76
//   error_a_face     _auxilliaries/error_assembly/boundary_assembly.py:79-129
// One thread per sub-triangle / interior face / boundary face; every CTA writes its three partial sums to a fixed slot
// and a single-CTA kernel adds the slots in index order: no floating-point atomics, bitwise reproducible.
#include "basis.cuh"

namespace rb {

constexpr int kErrThreads = 128;
constexpr int kErrBlocks = kNumSMs * 8;     // fixed grid -> fixed summation tree

struct ErrArgs {
    rb_geom g;
    rb_err_samples s;
    const double* u;          // solution [n_elem * d]
    double wv[RB_MAX_QV], rvx[RB_MAX_QV], rvy[RB_MAX_QV];
    double we[RB_MAX_QE], xe[RB_MAX_QE];
    LegConst lc;
    double sigma_pp, p2;
    double* partial;          // [3][3 * kErrBlocks]  (kernel, quantity, block)
};

__device__ __forceinline__ void block_reduce3(double a, double b, double c, double* __restrict__ out) {
    __shared__ double s_w[3][kErrThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { s_w[0][w] = a; s_w[1][w] = b; s_w[2][w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0, z = 0;
        for (int k = 0; k < kErrThreads / 32; ++k) { x += s_w[0][k]; y += s_w[1][k]; z += s_w[2][k]; }
        out[blockIdx.x] = x;
        out[kErrBlocks + blockIdx.x] = y;
        out[2 * kErrBlocks + blockIdx.x] = z;
    }
}

template <int P>
__device__ __forceinline__ double eval_uh(const double (&phi)[BasisDim<P>::D], const double* __restrict__ u) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < BasisDim<P>::D; ++k) s = fma(phi[k], u[k], s);
    return s;
}

// ---- volume: l2 += 1/2 sum W (u-uh)^2 ; dg += 1/2 sum W (e.a.e + c0 (u-uh)^2) ; h1 += 1/2 sum W |e|^2 -----------------
template <int P, bool DIFF>
__global__ void __launch_bounds__(kErrThreads) err_volume_kernel(const __grid_constant__ ErrArgs A) {
    constexpr int D = BasisDim<P>::D, QV = (P + 1) * (P + 1);
    double l2 = 0, dg = 0, h1 = 0;
    const size_t T = (size_t)A.g.n_tri;
    for (int t = blockIdx.x * kErrThreads + threadIdx.x; t < A.g.n_tri; t += kErrBlocks * kErrThreads) {
        const int e = A.g.tri_elem[t];
        const ElemBox bx = make_box(A.g.bbox, e);
        const double2 V0 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.tri[3 * (size_t)t]];
        const double2 V1 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.tri[3 * (size_t)t + 1]];
        const double2 V2 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.tri[3 * (size_t)t + 2]];
        const double e1x = 0.5 * (V1.x - V0.x), e1y = 0.5 * (V1.y - V0.y);
        const double e2x = 0.5 * (V2.x - V0.x), e2y = 0.5 * (V2.y - V0.y);
        const double hdet = 0.5 * fabs(e1x * e2y - e1y * e2x);
        double ue[D];
#pragma unroll
        for (int k = 0; k < D; ++k) ue[k] = A.u[(size_t)e * D + k];
        double tl2 = 0, tdg = 0, th1 = 0;
#pragma unroll 1
        for (int q = 0; q < QV; ++q) {
            const double r0 = A.rvx[q], r1 = A.rvy[q], l0 = 1.0 - r0 - r1;
            const double x = l0 * V0.x + r0 * V1.x + r1 * V2.x;
            const double y = l0 * V0.y + r0 * V1.y + r1 * V2.y;
            double phi[D], gx[D], gy[D];
            eval_basis<P, DIFF>(x, y, bx, A.lc, phi, gx, gy);
            const size_t m = (size_t)q * T + t;
            const double W = hdet * A.wv[q];
            const double du = A.s.vol_u[m] - eval_uh<P>(phi, ue);
            const double t1 = du * du;
            tl2 = fma(W, t1, tl2);
            double t2 = 0.0;
            if (DIFF) {
                const double2 gu = reinterpret_cast<const double2*>(A.s.vol_gu)[m];
                double ex = gu.x, ey = gu.y;
#pragma unroll
                for (int k = 0; k < D; ++k) { ex = fma(-gx[k], ue[k], ex); ey = fma(-gy[k], ue[k], ey); }
                const double2 a0 = reinterpret_cast<const double2*>(A.s.vol_a)[2 * m];
                const double2 a1 = reinterpret_cast<const double2*>(A.s.vol_a)[2 * m + 1];
                t2 = ex * (a0.x * ex + a0.y * ey) + ey * (a1.x * ex + a1.y * ey);
                th1 = fma(W, ex * ex + ey * ey, th1);
            }
            const double c0 = A.s.vol_c0 ? A.s.vol_c0[m] : 0.0;
            tdg = fma(W, t2 + c0 * t1, tdg);
        }
        l2 += tl2; dg += tdg; h1 += th1;
    }
    block_reduce3(l2, dg, h1, A.partial);
}

__device__ __forceinline__ double err_max_sub_area(const rb_geom& g, int e, double2 P0, double ex, double ey) {
    double kb = 0.0;
    for (int k = g.poly_off[e]; k < g.poly_off[e + 1]; ++k) {
        const double2 v = reinterpret_cast<const double2*>(g.nodes)[g.poly_vtx[k]];
        kb = fmax(kb, 0.5 * fabs(ex * (v.y - P0.y) - ey * (v.x - P0.x)));
    }
    return kb;
}

// ---- interior faces: dg += sum W (sigma + |b.n|/2) (u1 - u2)^2 ---------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kErrThreads) err_iface_kernel(const __grid_constant__ ErrArgs A) {
    constexpr int D = BasisDim<P>::D, QE = P + 1;
    double dg = 0;
    const size_t F = (size_t)A.g.n_iface;
    for (int f = blockIdx.x * kErrThreads + threadIdx.x; f < A.g.n_iface; f += kErrBlocks * kErrThreads) {
        const int e1 = A.g.ielem[2 * (size_t)f], e2 = A.g.ielem[2 * (size_t)f + 1];
        const ElemBox b1 = make_box(A.g.bbox, e1), b2 = make_box(A.g.bbox, e2);
        const double2 P0 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.iedge[2 * (size_t)f]];
        const double2 P1 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.iedge[2 * (size_t)f + 1]];
        const double2 mid = make_double2((P0.x + P1.x) * 0.5, (P0.y + P1.y) * 0.5);
        const double2 tan = make_double2(0.5 * (P1.x - P0.x), 0.5 * (P1.y - P0.y));
        const double De = sqrt(tan.x * tan.x + tan.y * tan.y);
        const double2 n = reinterpret_cast<const double2*>(A.g.inorm)[f];
        double sigma = 0.0;
        if (A.s.if_a_mid) {
            const double2 m0 = reinterpret_cast<const double2*>(A.s.if_a_mid)[2 * (size_t)f];
            const double2 m1 = reinterpret_cast<const double2*>(A.s.if_a_mid)[2 * (size_t)f + 1];
            const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;
            const double ex = P1.x - P0.x, ey = P1.y - P0.y;
            const double a1 = A.g.area[e1], a2 = A.g.area[e2];
            const double c1 = fmin(a1 / err_max_sub_area(A.g, e1, P0, ex, ey), A.p2);
            const double c2 = fmin(a2 / err_max_sub_area(A.g, e2, P0, ex, ey), A.p2);
            sigma = A.sigma_pp * lam * (2.0 * De) * fmax(c1 / a1, c2 / a2);
        }
        double u1[D], u2[D];
#pragma unroll
        for (int k = 0; k < D; ++k) { u1[k] = A.u[(size_t)e1 * D + k]; u2[k] = A.u[(size_t)e2 * D + k]; }
        double s = 0.0;
#pragma unroll 1
        for (int q = 0; q < QE; ++q) {
            const double x = mid.x + A.xe[q] * tan.x, y = mid.y + A.xe[q] * tan.y;
            double phi[D], gx[D], gy[D];
            eval_basis<P, false>(x, y, b1, A.lc, phi, gx, gy);
            double jump = eval_uh<P>(phi, u1);
            eval_basis<P, false>(x, y, b2, A.lc, phi, gx, gy);
            jump -= eval_uh<P>(phi, u2);
            double wgt = sigma;
            if (A.s.if_b) {
                const double2 bb = reinterpret_cast<const double2*>(A.s.if_b)[(size_t)q * F + f];
                wgt += 0.5 * fabs(bb.x * n.x + bb.y * n.y);
            }
            s = fma(De * A.we[q] * wgt, jump * jump, s);
        }
        dg += s;
    }
    block_reduce3(0.0, dg, 0.0, A.partial + 3 * kErrBlocks);
}

// ---- boundary faces: elliptic: dg += sum W sigma (u-uh)^2 ; with advection (all faces): dg += sum W |b.n|/2 (u-uh)^2 ----
template <int P>
__global__ void __launch_bounds__(kErrThreads) err_bface_kernel(const __grid_constant__ ErrArgs A) {
    constexpr int D = BasisDim<P>::D, QE = P + 1;
    double dg = 0;
    const size_t FB = (size_t)A.g.n_bface;
    for (int f = blockIdx.x * kErrThreads + threadIdx.x; f < A.g.n_bface; f += kErrBlocks * kErrThreads) {
        const int e = A.g.belem[f];
        const ElemBox bx = make_box(A.g.bbox, e);
        const double2 P0 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.bedge[2 * (size_t)f]];
        const double2 P1 = reinterpret_cast<const double2*>(A.g.nodes)[A.g.bedge[2 * (size_t)f + 1]];
        const double2 mid = make_double2((P0.x + P1.x) * 0.5, (P0.y + P1.y) * 0.5);
        const double2 tan = make_double2(0.5 * (P1.x - P0.x), 0.5 * (P1.y - P0.y));
        const double De = sqrt(tan.x * tan.x + tan.y * tan.y);
        const double2 n = reinterpret_cast<const double2*>(A.g.bnorm)[f];
        double sigma = 0.0;
        if (A.s.bf_a_mid && A.s.bf_flags && (A.s.bf_flags[f] & 1)) {
            const double2 m0 = reinterpret_cast<const double2*>(A.s.bf_a_mid)[2 * (size_t)f];
            const double2 m1 = reinterpret_cast<const double2*>(A.s.bf_a_mid)[2 * (size_t)f + 1];
            const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;
            const double ar = A.g.area[e];
            const double kb = err_max_sub_area(A.g, e, P0, P1.x - P0.x, P1.y - P0.y);
            sigma = A.sigma_pp * lam * (2.0 * De) * fmin(ar / kb, A.p2) / ar;
        }
        double ue[D];
#pragma unroll
        for (int k = 0; k < D; ++k) ue[k] = A.u[(size_t)e * D + k];
        double s = 0.0;
#pragma unroll 1
        for (int q = 0; q < QE; ++q) {
            const double x = mid.x + A.xe[q] * tan.x, y = mid.y + A.xe[q] * tan.y;
            double phi[D], gx[D], gy[D];
            eval_basis<P, false>(x, y, bx, A.lc, phi, gx, gy);
            const size_t m = (size_t)q * FB + f;
            const double du = A.s.bf_u[m] - eval_uh<P>(phi, ue);
            double wgt = sigma;
            if (A.s.bf_b) {
                const double2 bb = reinterpret_cast<const double2*>(A.s.bf_b)[m];
                wgt += 0.5 * fabs(bb.x * n.x + bb.y * n.y);
            }
            s = fma(De * A.we[q] * wgt, du * du, s);
        }
        dg += s;
    }
    block_reduce3(0.0, dg, 0.0, A.partial + 6 * kErrBlocks);
}

// out[0] = l2^2, out[1] = dg^2, out[2] = h1^2 ; slots added in index order (volume, interior, boundary)
__global__ void __launch_bounds__(32) err_final_kernel(const double* __restrict__ partial, double* __restrict__ out) {
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int kern = 0; kern < 3; ++kern)
            for (int b = 0; b < kErrBlocks; ++b) s += partial[(size_t)kern * 3 * kErrBlocks + threadIdx.x * kErrBlocks + b];
        out[threadIdx.x] = s;
    }
}

// ---- point evaluation of the DG solution (replaces tools.evaluate, reyna/DGFEM/two_dimensional/tools.py:117-153) ------
// out[i] = sum_k u[elem[i]*d + k] phi_k^{elem[i]}(x_i); bbox-scaled tensor-Legendre basis exactly as in the assembly
template <int P>
__global__ void __launch_bounds__(256) eval_points_kernel(int n_points, const double* __restrict__ bbox,
                                                          const double* __restrict__ xy, const int32_t* __restrict__ elem,
                                                          const double* __restrict__ u, LegConst lc, double* __restrict__ out) {
    constexpr int D = BasisDim<P>::D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += gridDim.x * blockDim.x) {
        const int e = elem[i];
        const ElemBox bx = make_box(bbox, e);
        const double2 pt = reinterpret_cast<const double2*>(xy)[i];
        double phi[D], gx[D], gy[D];
        eval_basis<P, false>(pt.x, pt.y, bx, lc, phi, gx, gy);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) s = fma(phi[k], u[(size_t)e * D + k], s);
        out[i] = s;
    }
}

template <int P>
static int launch_errors(const ErrArgs& A, bool diff, cudaStream_t st) {
    if (A.g.n_tri > 0) {
        if (diff) err_volume_kernel<P, true><<<kErrBlocks, kErrThreads, 0, st>>>(A);
        else err_volume_kernel<P, false><<<kErrBlocks, kErrThreads, 0, st>>>(A);
        RB_LAUNCH_CHECK("err_volume_kernel");
    }
    if (A.g.n_iface > 0 && (A.s.if_a_mid || A.s.if_b)) {
        err_iface_kernel<P><<<kErrBlocks, kErrThreads, 0, st>>>(A);
        RB_LAUNCH_CHECK("err_iface_kernel");
    }
    if (A.g.n_bface > 0 && (A.s.bf_a_mid || A.s.bf_b)) {
        err_bface_kernel<P><<<kErrBlocks, kErrThreads, 0, st>>>(A);
        RB_LAUNCH_CHECK("err_bface_kernel");
    }
    return 0;
}

}  // namespace rb

using namespace rb;

extern "C" size_t reyna_b200_errors_workspace(void) { return sizeof(double) * 9 * (size_t)kErrBlocks; }

extern "C" int reyna_b200_errors(const rb_geom* g, const rb_quad* q, const rb_err_samples* s, double sigma_D,
                                 const double* solution, double* out3, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    RB_REQUIRE(g && q && s && solution && out3 && workspace, RB_E_ARG, "reyna_b200_errors: null argument");
    RB_REQUIRE(workspace_bytes >= reyna_b200_errors_workspace(), RB_E_WORKSPACE, "reyna_b200_errors: workspace too small");
    const int p = q->p;
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_errors: polynomial degree must be 1..3");
    RB_REQUIRE(s->vol_u || g->n_tri == 0, RB_E_ARG, "reyna_b200_errors: exact-solution samples missing");
    const bool diff = s->vol_a != nullptr;
    RB_REQUIRE(!diff || s->vol_gu, RB_E_ARG, "reyna_b200_errors: diffusion given without gradient samples");
    RB_REQUIRE(!(s->bf_a_mid || s->bf_b) || s->bf_u || g->n_bface == 0, RB_E_ARG,
               "reyna_b200_errors: boundary exact-solution samples missing");
    ErrArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *g;
    A.s = *s;
    A.u = solution;
    const int qv = (p + 1) * (p + 1), qe = p + 1;
    for (int i = 0; i < qv; ++i) { A.wv[i] = q->wv[i]; A.rvx[i] = q->rv[i][0]; A.rvy[i] = q->rv[i][1]; }
    for (int i = 0; i < qe; ++i) { A.we[i] = q->we[i]; A.xe[i] = q->xe[i]; }
    A.lc = make_leg_const();
    A.sigma_pp = sigma_D * (double)(p * p);
    A.p2 = (double)(p * p);
    A.partial = (double*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    RB_CUDA(cudaMemsetAsync(workspace, 0, reyna_b200_errors_workspace(), st));
    int rc;
    switch (p) {
        case 1: rc = launch_errors<1>(A, diff, st); break;
        case 2: rc = launch_errors<2>(A, diff, st); break;
        default: rc = launch_errors<3>(A, diff, st); break;
    }
    if (rc) return rc;
    err_final_kernel<<<1, 32, 0, st>>>(A.partial, out3);
    RB_LAUNCH_CHECK("err_final_kernel");
    return 0;
}

extern "C" int reyna_b200_evaluate(int p, int32_t n_points, const double* bbox, const double* xy, const int32_t* elem,
                                   const double* solution, double* out, void* stream) {
    RB_REQUIRE(bbox && xy && elem && solution && out, RB_E_ARG, "reyna_b200_evaluate: null argument");
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_evaluate: polynomial degree must be 1..3");
    if (n_points <= 0) return 0;
    const LegConst lc = make_leg_const();
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = grid_for(n_points, 256, kNumSMs * 8);
    switch (p) {
        case 1: eval_points_kernel<1><<<blocks, 256, 0, st>>>(n_points, bbox, xy, elem, solution, lc, out); break;
        case 2: eval_points_kernel<2><<<blocks, 256, 0, st>>>(n_points, bbox, xy, elem, solution, lc, out); break;
        default: eval_points_kernel<3><<<blocks, 256, 0, st>>>(n_points, bbox, xy, elem, solution, lc, out); break;
    }
    RB_LAUNCH_CHECK("eval_points_kernel");
    return 0;
}
