// The fused DG assembly kernel: volume terms + interior-face terms -> CSR values, one launch, no atomics on doubles.
//
// Replaces, for all elements at once,
//   DGFEM._stiffness_matrix + localstiff            reyna/DGFEM/two_dimensional/main.py:275-327,
//                                                   _auxilliaries/assembly/full_assembly.py:9-77
//   DGFEM._interior_stiffness_matrix + int_localstiff   main.py:329-374, full_assembly.py:80-214
//   the scipy COO->CSR conversions and `B += ...`    main.py:319-320, 371-372, 211
//
// Design (B200): a CTA owns TE consecutive elements, i.e. TE consecutive block rows of B.
//   phase A  one thread per sub-triangle: Legendre basis + gradients in registers, the d x d local block (and the load)
//            accumulated in registers over the (p+1)^2 points, then summed per element IN TRIANGLE ORDER through a
//            shared-memory stage (the reference's `s[:, elem] += ...` order, main.py:315).
//   phase B  one thread per (element, interior face) item of the element->face adjacency: the diagonal part
//            B[K,K] of the SIPG/upwind face matrix goes through the same staged per-element sum, the off-diagonal part
//            B[K,K'] is complete after one face and is stored straight into its final CSR position.
//            Each face is therefore visited from both of its elements, each side producing only its own block rows:
//            no intermediate face-matrix buffer, no scatter-add, bitwise reproducible.
//   phase C  the summed diagonal blocks and load are written to the CSR value array / load vector.
// Coefficient samples are q-major ([q][item]) so that the per-point loads of a warp are contiguous.
// For d = 10 (p = 3) a thread owns half of the block rows (RS = 2) to keep the accumulators in registers; the halves are
// assigned per half-CTA so that the compile-time row offset stays warp-uniform.
#include "basis.cuh"

namespace rb {

constexpr int kBD = 128;   // threads per CTA
#ifndef RB_FUSE_LIMIT
#define RB_FUSE_LIMIT 36
#endif
constexpr int kFuseLimit = RB_FUSE_LIMIT;   // accumulators per block up to which a face item forms both blocks in one sweep

struct AsmArgs {
    rb_geom g;
    const int32_t* adj_off;
    const int32_t* adj_item;
    const int32_t* grp_off;
    rb_coefs c;
    double wv[RB_MAX_QV], rvx[RB_MAX_QV], rvy[RB_MAX_QV];
    double we[RB_MAX_QE], xe[RB_MAX_QE];
    LegConst lc;
    double sigma_pp;   // sigma_D * p^2
    double p2;         // p^2
    double* data;
    double* load;
    unsigned long long* zero_count;
    int te;            // elements per CTA
    int has_dups;      // adjacency contains merged (duplicate-neighbour) items
    const ElemBox* ebox;   // [n_elem]  bounding-box scalings, written by dg_prepare_kernel
    const double* fsigma;  // [n_iface] SIPG penalty per interior face (diffusion only), written by dg_prepare_kernel
};

// ---------------------------------------------------------------------------------------------------------------
// prepass: per-element basis scalings and per-face penalty.  Both are needed by every (element, face) item from either
// side of a face (and the polygon loops of the penalty are latency-bound), so they are computed once here.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double max_sub_area_g(const rb_geom& g, int e, double2 P0, double ex, double ey) {
    // |k_b|_K = max over the vertices v of K of 0.5*|(P1-P0) x (v-P0)|   (full_assembly.py:142-143)
    double kb = 0.0;
    const int lo = g.poly_off[e], hi = g.poly_off[e + 1];
    for (int k = lo; k < hi; ++k) {
        const double2 v = reinterpret_cast<const double2*>(g.nodes)[g.poly_vtx[k]];
        kb = fmax(kb, 0.5 * fabs(ex * (v.y - P0.y) - ey * (v.x - P0.x)));
    }
    return kb;
}

__global__ void __launch_bounds__(256) dg_prepare_kernel(rb_geom g, const double* __restrict__ if_a_mid, double sigma_pp,
                                                         double p2, ElemBox* __restrict__ ebox,
                                                         double* __restrict__ fsigma) {
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.n_elem; e += stride) ebox[e] = make_box(g.bbox, e);
    if (if_a_mid == nullptr) return;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.n_iface; f += stride) {
        const int e1 = g.ielem[2 * (size_t)f], e2 = g.ielem[2 * (size_t)f + 1];
        const double2 P0 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f]];
        const double2 P1 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f + 1]];
        const double2 n = reinterpret_cast<const double2*>(g.inorm)[f];
        const double2 m0 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f];
        const double2 m1 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f + 1];
        // lambda = n . a(mid) . n   (full_assembly.py:140)
        const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;
        const double ex = P1.x - P0.x, ey = P1.y - P0.y;
        const double tx = 0.5 * ex, ty = 0.5 * ey;
        const double De = sqrt(tx * tx + ty * ty);
        const double a1 = g.area[e1], a2 = g.area[e2];
        const double c1 = fmin(a1 / max_sub_area_g(g, e1, P0, ex, ey), p2);
        const double c2 = fmin(a2 / max_sub_area_g(g, e2, P0, ex, ey), p2);
        fsigma[f] = sigma_pp * lam * (2.0 * De) * fmax(c1 / a1, c2 / a2);   // full_assembly.py:146-148
    }
}


template <int P, int RS>
struct Dims {
    static constexpr int D = (P + 1) * (P + 2) / 2;
    static constexpr int DD = D * D;
    static constexpr int NV = DD + D;       // per-element accumulators in shared memory: block + load
    static constexpr int RJ = D / RS;       // block rows per thread
    static constexpr int QV = (P + 1) * (P + 1);
    static constexpr int QE = P + 1;
    static constexpr int IT = kBD / RS;     // items per pass
    static constexpr int CH = (RJ * D + RJ) > 24 ? ((RJ * D + RJ + 1) / 2 > 28 ? 22 : (RJ * D + RJ + 1) / 2)
                                                 : (RJ * D + RJ);   // staged values per round
    static_assert(D % RS == 0, "row split must divide the local dimension");
};

__host__ __device__ inline size_t asm_smem_bytes(int nv, int ch, int te) {
    // s_acc [te][nv] doubles | s_stage [ch][kBD+1] doubles | 3 x [te+1] ints + [te] ints
    return sizeof(double) * ((size_t)te * nv + (size_t)ch * (kBD + 1)) + sizeof(int) * (4 * ((size_t)te + 1));
}

// ---------------------------------------------------------------------------------------------------------------
// staged, ordered per-element sum of per-thread register vectors
// ---------------------------------------------------------------------------------------------------------------
// vals[0 .. NT) of the thread handling item (base + it) with row part `part`; items of element seg are
// [off[seg], off[seg+1]).  Value vv of part `part` lands at s_acc[seg*NV + dst0(part) + vv] for vv < NBLK and at
// s_acc[seg*NV + dst1(part) + (vv-NBLK)] for the tail (load entries).
template <int NT, int CH, int IT, int RS, int NV>
__device__ __forceinline__ void staged_sum(const double (&vals)[NT], bool active, int base, int ne,
                                           const int* __restrict__ s_off, double* __restrict__ s_acc,
                                           double* __restrict__ s_stage, int nblk, int dst0_stride, int dst1_base,
                                           int dst1_stride) {
    constexpr int NCH = (NT + CH - 1) / CH;
    const int tid = threadIdx.x;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        if (active) {
#pragma unroll
            for (int v = 0; v < CH; ++v)
                if (ch * CH + v < NT) s_stage[v * (kBD + 1) + tid] = vals[ch * CH + v];
        }
        __syncthreads();
        const int nv_here = (NT - ch * CH) < CH ? (NT - ch * CH) : CH;
        const int work = ne * RS * nv_here;
        for (int idx = tid; idx < work; idx += kBD) {
            const int v = idx % nv_here;
            const int sp = idx / nv_here;
            const int part = sp % RS;
            const int seg = sp / RS;
            int lo = s_off[seg] - base, hi = s_off[seg + 1] - base;
            lo = lo < 0 ? 0 : lo;
            hi = hi > IT ? IT : hi;
            if (lo < hi) {
                const int vv = ch * CH + v;
                const int dst = seg * NV + (vv < nblk ? part * dst0_stride + vv : dst1_base + part * dst1_stride + (vv - nblk));
                double s = s_acc[dst];
                const double* src = s_stage + v * (kBD + 1) + part * IT;
                for (int t = lo; t < hi; ++t) s += src[t];
                s_acc[dst] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// phase A: one sub-triangle
// ---------------------------------------------------------------------------------------------------------------
struct VolCoef {
    double a00, a01, a10, a11, b0, b1, c, f;
};

template <bool DIFF, bool ADV>
__device__ __forceinline__ VolCoef load_vol_coef(const AsmArgs& A, size_t m) {
    VolCoef k;
    k.a00 = k.a01 = k.a10 = k.a11 = k.b0 = k.b1 = 0.0;
    if (DIFF) {
        const double2 ar0 = __ldg(reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m);
        const double2 ar1 = __ldg(reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m + 1);
        k.a00 = ar0.x; k.a01 = ar0.y; k.a10 = ar1.x; k.a11 = ar1.y;
    }
    if (ADV) {
        const double2 bb = __ldg(reinterpret_cast<const double2*>(A.c.vol_b) + m);
        k.b0 = bb.x; k.b1 = bb.y;
    }
    k.c = A.c.vol_c ? __ldg(A.c.vol_c + m) : 0.0;
    k.f = A.c.vol_f ? __ldg(A.c.vol_f + m) : 0.0;
    return k;
}

template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void triangle_block(const AsmArgs& A, int t, double (&acc)[Dims<P, RS>::RJ * Dims<P, RS>::D + Dims<P, RS>::RJ]) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, J0 = PART * RJ;
    const size_t T = (size_t)A.g.n_tri;
    // the samples of the first point are requested before the geometry chain (tri -> nodes, tri_elem -> box) resolves
    VolCoef cur = load_vol_coef<DIFF, ADV>(A, (size_t)t);
    const int e = A.g.tri_elem[t];
    const int i0 = A.g.tri[3 * (size_t)t], i1 = A.g.tri[3 * (size_t)t + 1], i2 = A.g.tri[3 * (size_t)t + 2];
    const ElemBox bx = A.ebox[e];
    const double2 V0 = reinterpret_cast<const double2*>(A.g.nodes)[i0];
    const double2 V1 = reinterpret_cast<const double2*>(A.g.nodes)[i1];
    const double2 V2 = reinterpret_cast<const double2*>(A.g.nodes)[i2];
    // |det(0.5*[V1-V0; V2-V0])| (full_assembly.py:40-41); the extra 0.5 is the `0.5 * z` of :75-77
    const double e1x = 0.5 * (V1.x - V0.x), e1y = 0.5 * (V1.y - V0.y);
    const double e2x = 0.5 * (V2.x - V0.x), e2y = 0.5 * (V2.y - V0.y);
    const double hdet = 0.5 * fabs(e1x * e2y - e1y * e2x);
#pragma unroll 1
    for (int q = 0; q < DM::QV; ++q) {
        // software pipeline: the samples of point q+1 are in flight while point q is contracted (the kernel runs at
        // 8 warps per SM, so memory latency has to be hidden inside the thread)
        const int qn = q + 1 < DM::QV ? q + 1 : q;
        const VolCoef nxt = load_vol_coef<DIFF, ADV>(A, (size_t)qn * T + t);
        const double r0 = A.rvx[q], r1 = A.rvy[q];
        const double l0 = 1.0 - r0 - r1;
        const double x = l0 * V0.x + r0 * V1.x + r1 * V2.x;
        const double y = l0 * V0.y + r0 * V1.y + r1 * V2.y;
        double phi[D], gx[D], gy[D];
        eval_basis<P, true>(x, y, bx, A.lc, phi, gx, gy);
        const double W = hdet * A.wv[q];
        // W folded into the coefficients once per point
        const double a00 = W * cur.a00, a01 = W * cur.a01, a10 = W * cur.a10, a11 = W * cur.a11;
        const double b0 = W * cur.b0, b1 = W * cur.b1, cq = W * cur.c;
        // column (trial i) quantities u = W * grad(phi_i)^T a, s = W * (b.grad(phi_i) + c phi_i) are formed one column
        // at a time so that only the accumulators and the basis stay live
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double u0 = 0, u1 = 0;
            if (DIFF) {
                u0 = fma(gx[i], a00, gy[i] * a10);
                u1 = fma(gx[i], a01, gy[i] * a11);
            }
            double si = cq * phi[i];
            if (ADV) si = fma(b0, gx[i], fma(b1, gy[i], si));
#pragma unroll
            for (int jj = 0; jj < RJ; ++jj) {
                double v = acc[jj * D + i];
                if (DIFF) {
                    v = fma(u0, gx[J0 + jj], v);
                    v = fma(u1, gy[J0 + jj], v);
                }
                v = fma(si, phi[J0 + jj], v);
                acc[jj * D + i] = v;
            }
        }
        if (A.c.vol_f) {
            const double wf = W * cur.f;
#pragma unroll
            for (int jj = 0; jj < RJ; ++jj) acc[RJ * D + jj] = fma(wf, phi[J0 + jj], acc[RJ * D + jj]);
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// phase B: one (element, interior face) item
// ---------------------------------------------------------------------------------------------------------------
struct FaceCtx {
    double2 mid, tan, n;
    double De, sigma, sgn;
    int f, o;        // face, other element
    bool down;       // this side is the downwind side of the face
};

template <bool DIFF>
__device__ __forceinline__ FaceCtx face_setup(const AsmArgs& A, int item) {
    FaceCtx c;
    c.f = item >> 1;
    const int side = item & 1;
    c.o = A.g.ielem[2 * (size_t)c.f + (1 - side)];
    c.sgn = side ? -1.0 : 1.0;
    const int v0 = A.g.iedge[2 * (size_t)c.f], v1 = A.g.iedge[2 * (size_t)c.f + 1];
    const double2 P0 = reinterpret_cast<const double2*>(A.g.nodes)[v0];
    const double2 P1 = reinterpret_cast<const double2*>(A.g.nodes)[v1];
    c.mid = make_double2((P0.x + P1.x) * 0.5, (P0.y + P1.y) * 0.5);
    c.tan = make_double2(0.5 * (P1.x - P0.x), 0.5 * (P1.y - P0.y));
    c.De = sqrt(c.tan.x * c.tan.x + c.tan.y * c.tan.y);
    c.n = reinterpret_cast<const double2*>(A.g.inorm)[c.f];
    c.down = A.c.if_upwind ? ((A.c.if_upwind[c.f] != 0) == (side == 1)) : false;
    c.sigma = DIFF ? A.fsigma[c.f] : 0.0;
    return c;
}

struct FaceCoef {
    double a00, a01, a10, a11, b0, b1;
};

template <bool DIFF, bool ADV>
__device__ __forceinline__ FaceCoef load_face_coef(const AsmArgs& A, size_t m) {
    FaceCoef k;
    k.a00 = k.a01 = k.a10 = k.a11 = k.b0 = k.b1 = 0.0;
    if (DIFF) {
        const double2 ar0 = __ldg(reinterpret_cast<const double2*>(A.c.if_a) + 2 * m);
        const double2 ar1 = __ldg(reinterpret_cast<const double2*>(A.c.if_a) + 2 * m + 1);
        k.a00 = ar0.x; k.a01 = ar0.y; k.a10 = ar1.x; k.a11 = ar1.y;
    }
    if (ADV) {
        const double2 bb = __ldg(reinterpret_cast<const double2*>(A.c.if_b) + m);
        k.b0 = bb.x; k.b1 = bb.y;
    }
    return k;
}

// Accumulate, for rows [J0, J0+RJ) of element e, the own-element (diagonal) block into dacc and/or the other-element
// block into oacc.  With P_r = s_r phi_r, Q_r = -F_r/2, V_c = W s_c phi_c, U_c = sigma V_c - W F_c / 2 + gamma V_c the
// SIPG + upwind face matrix (full_assembly.py:174-212) is B[r,c] += P_r U_c + Q_r V_c.
template <int P, int RS, int PART, bool DIFF, bool ADV, bool DO_DIAG, bool DO_OFF>
__device__ __forceinline__ void face_blocks(const AsmArgs& A, const FaceCtx& c, const ElemBox& bxe, const ElemBox& bxo,
                                            double* __restrict__ dacc, double* __restrict__ oacc) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, J0 = PART * RJ;
    const size_t F = (size_t)A.g.n_iface;
    FaceCoef cur = load_face_coef<DIFF, ADV>(A, (size_t)c.f);
#pragma unroll 1
    for (int q = 0; q < DM::QE; ++q) {
        const int qn = q + 1 < DM::QE ? q + 1 : q;
        const FaceCoef nxt = load_face_coef<DIFF, ADV>(A, (size_t)qn * F + c.f);
        const double x = c.mid.x + A.xe[q] * c.tan.x;
        const double y = c.mid.y + A.xe[q] * c.tan.y;
        const double W = c.De * A.we[q];
        double atn0 = 0, atn1 = 0, gamma = 0;
        if (DIFF) {
            // F_k = n . (a grad phi_k) = (a^T n) . grad phi_k
            atn0 = c.n.x * cur.a00 + c.n.y * cur.a10;
            atn1 = c.n.x * cur.a01 + c.n.y * cur.a11;
        }
        if (ADV) {
            const double beta = cur.b0 * c.n.x + cur.b1 * c.n.y;
            gamma = c.down ? -c.sgn * beta : 0.0;
        }
        const double hw0 = -0.5 * W * atn0, hw1 = -0.5 * W * atn1;     // -W/2 * a^T n
        const double sg = c.sigma + gamma;
        double phi[D], gx[D], gy[D];
        eval_basis<P, DIFF>(x, y, bxe, A.lc, phi, gx, gy);
        double Pr[RJ], Qr[RJ];
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) {
            Pr[jj] = c.sgn * phi[J0 + jj];
            if (DIFF) Qr[jj] = -0.5 * (atn0 * gx[J0 + jj] + atn1 * gy[J0 + jj]);
        }
        const double Ws = W * c.sgn;
        if (DO_DIAG) {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double Vc = Ws * phi[i];
                double Uc = sg * Vc;
                if (DIFF) Uc = fma(hw0, gx[i], fma(hw1, gy[i], Uc));
#pragma unroll
                for (int jj = 0; jj < RJ; ++jj) {
                    double v = dacc[jj * D + i];
                    v = fma(Pr[jj], Uc, v);
                    if (DIFF) v = fma(Qr[jj], Vc, v);
                    dacc[jj * D + i] = v;
                }
            }
        }
        if (DO_OFF) {
            eval_basis<P, DIFF>(x, y, bxo, A.lc, phi, gx, gy);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double Vc = -Ws * phi[i];
                double Uc = sg * Vc;
                if (DIFF) Uc = fma(hw0, gx[i], fma(hw1, gy[i], Uc));
#pragma unroll
                for (int jj = 0; jj < RJ; ++jj) {
                    double v = oacc[jj * D + i];
                    v = fma(Pr[jj], Uc, v);
                    if (DIFF) v = fma(Qr[jj], Vc, v);
                    oacc[jj * D + i] = v;
                }
            }
        }
        cur = nxt;
    }
}

template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void item_pass(const AsmArgs& A, int e, int k, int k_end, int slot, int nb,
                                          int64_t row_base, double (&dacc)[Dims<P, RS>::RJ * Dims<P, RS>::D],
                                          unsigned& zeros) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, J0 = PART * RJ;
    constexpr bool FUSE = (RJ * D <= kFuseLimit);   // both blocks in one sweep when the accumulators fit in registers
    const ElemBox bxe = A.ebox[e];
    double oacc[RJ * D];
#pragma unroll
    for (int v = 0; v < RJ * D; ++v) oacc[v] = 0.0;
    // a group is one item plus any following items flagged as "same neighbour" (merged blocks; non-Voronoi meshes)
    int kk = k;
    do {
        const int item = A.adj_item[kk] & 0x7fffffff;
        const FaceCtx c = face_setup<DIFF>(A, item);
        const ElemBox bxo = A.ebox[c.o];
        if (FUSE) face_blocks<P, RS, PART, DIFF, ADV, true, true>(A, c, bxe, bxo, dacc, oacc);
        else face_blocks<P, RS, PART, DIFF, ADV, false, true>(A, c, bxe, bxo, dacc, oacc);
        ++kk;
    } while (A.has_dups && kk < k_end && A.adj_item[kk] < 0);
    // off-diagonal block: final after this group -> straight to its CSR slot
    const int rowlen = nb * D;
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) {
        double* dst = A.data + row_base + (int64_t)(J0 + jj) * rowlen + (int64_t)slot * D;
        if ((D & 1) == 0) {
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                reinterpret_cast<double2*>(dst)[i >> 1] = make_double2(oacc[jj * D + i], oacc[jj * D + i + 1]);
                zeros += (oacc[jj * D + i] == 0.0) + (oacc[jj * D + i + 1] == 0.0);
            }
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) {
                dst[i] = oacc[jj * D + i];
                zeros += (oacc[jj * D + i] == 0.0);
            }
        }
    }
    if (!FUSE) {
        // second sweep for the diagonal part, after the off-diagonal registers have been retired
        kk = k;
        do {
            const int item = A.adj_item[kk] & 0x7fffffff;
            const FaceCtx c = face_setup<DIFF>(A, item);
            face_blocks<P, RS, PART, DIFF, ADV, true, false>(A, c, bxe, bxe, dacc, dacc);
            ++kk;
        } while (A.has_dups && kk < k_end && A.adj_item[kk] < 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kBD) dg_assemble_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, DD = DM::DD, NV = DM::NV, RJ = DM::RJ, IT = DM::IT, CH = DM::CH;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int ea = blockIdx.x * A.te;
    const int ne = min(A.te, A.g.n_rows - ea);
    double* s_acc = smem;                                   // [te][NV]
    double* s_stage = s_acc + (size_t)A.te * NV;            // [CH][kBD+1]
    int* s_toff = reinterpret_cast<int*>(s_stage + (size_t)CH * (kBD + 1));   // [te+1]
    int* s_aoff = s_toff + (A.te + 1);
    int* s_goff = s_aoff + (A.te + 1);
    int* s_dslot = s_goff + (A.te + 1);                     // [te]

    for (int i = tid; i <= ne; i += kBD) {
        s_toff[i] = A.g.tri_off[ea + i];
        s_aoff[i] = A.adj_off[ea + i];
        s_goff[i] = A.grp_off[ea + i];
        if (i < ne) s_dslot[i] = 0;
    }
    for (int i = tid; i < ne * NV; i += kBD) s_acc[i] = 0.0;
    __syncthreads();

    const int part = tid / IT;          // warp-uniform row part
    const int lt = tid - part * IT;     // item within the pass

    // ---- phase A: volume terms ---------------------------------------------------------------------------------
    {
        const int ta = s_toff[0], tb = s_toff[ne];
        for (int base = ta; base < tb; base += IT) {
            const int t = base + lt;
            const bool active = t < tb;
            double acc[RJ * D + RJ];
#pragma unroll
            for (int v = 0; v < RJ * D + RJ; ++v) acc[v] = 0.0;
            if (active) {
                if (RS == 1 || part == 0) triangle_block<P, RS, 0, DIFF, ADV>(A, t, acc);
                else triangle_block<P, RS, RS - 1, DIFF, ADV>(A, t, acc);
            }
            staged_sum<RJ * D + RJ, CH, IT, RS, NV>(acc, active, base, ne, s_toff, s_acc, s_stage, RJ * D, RJ * D, DD, RJ);
        }
    }

    // ---- phase B: interior faces -------------------------------------------------------------------------------
    unsigned zeros = 0;
    {
        const int ka = s_aoff[0], kb = s_aoff[ne];
        for (int base = ka; base < kb; base += IT) {
            const int k = base + lt;
            bool active = k < kb;
            double dacc[RJ * D];
#pragma unroll
            for (int v = 0; v < RJ * D; ++v) dacc[v] = 0.0;
            if (active && A.adj_item[k] < 0) active = false;    // merged into the group leader
            if (active) {
                // element of this item: last seg with s_aoff[seg] <= k
                int lo = 0, hi = ne;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_aoff[mid] <= k) lo = mid; else hi = mid;
                }
                const int seg = lo, e = ea + seg;
                int gi = k - s_aoff[seg];
                if (A.has_dups)
                    for (int a = s_aoff[seg]; a < k; ++a) gi -= (A.adj_item[a] < 0);
                const int item = A.adj_item[k];
                const int o = A.g.ielem[2 * (size_t)(item >> 1) + (1 - (item & 1))];
                const int nb = s_goff[seg + 1] - s_goff[seg] + 1;
                const int slot = gi + (o > e ? 1 : 0);
                if (o < e && (RS == 1 || part == 0)) atomicAdd(&s_dslot[seg], 1);
                const int64_t row_base = (int64_t)DD * (s_goff[seg] + e);
                if (RS == 1 || part == 0) item_pass<P, RS, 0, DIFF, ADV>(A, e, k, s_aoff[seg + 1], slot, nb, row_base, dacc, zeros);
                else item_pass<P, RS, RS - 1, DIFF, ADV>(A, e, k, s_aoff[seg + 1], slot, nb, row_base, dacc, zeros);
            }
            // merged followers stay inactive: their stage column must read as zero -> stage zeros for them
            const bool in_range = k < kb;
            staged_sum<RJ * D, CH, IT, RS, NV>(dacc, in_range, base, ne, s_aoff, s_acc, s_stage, RJ * D, RJ * D, DD, RJ);
        }
    }
    __syncthreads();

    // ---- phase C: diagonal blocks and load -> global ---------------------------------------------------------------
    for (int idx = tid; idx < ne * DD; idx += kBD) {
        const int seg = idx / DD, rem = idx - seg * DD;
        const int j = rem / D, i = rem - j * D;
        const int e = ea + seg;
        const int nb = s_goff[seg + 1] - s_goff[seg] + 1;
        const int64_t pos = (int64_t)DD * (s_goff[seg] + e) + (int64_t)j * nb * D + (int64_t)s_dslot[seg] * D + i;
        const double v = s_acc[seg * NV + rem];
        A.data[pos] = v;
        zeros += (v == 0.0);
    }
    for (int idx = tid; idx < ne * D; idx += kBD) {
        const int seg = idx / D, i = idx - seg * D;
        A.load[(size_t)ea * D + idx] = s_acc[seg * NV + DD + i];
    }
    if (zeros) atomicAdd(A.zero_count, (unsigned long long)zeros);
}

template <int P, int RS, bool DIFF, bool ADV>
static int launch_assemble(const AsmArgs& A, cudaStream_t st) {
    using DM = Dims<P, RS>;
    const size_t smem = asm_smem_bytes(DM::NV, DM::CH, A.te);
    auto kern = dg_assemble_kernel<P, RS, DIFF, ADV>;
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)div_up(A.g.n_rows, A.te);
    kern<<<blocks, kBD, smem, st>>>(A);
    RB_LAUNCH_CHECK("dg_assemble_kernel");
    return 0;
}

template <int P, int RS>
static int dispatch_coefs(const AsmArgs& A, bool diff, bool adv, cudaStream_t st) {
    if (diff && adv) return launch_assemble<P, RS, true, true>(A, st);
    if (diff) return launch_assemble<P, RS, true, false>(A, st);
    return launch_assemble<P, RS, false, true>(A, st);
}

}  // namespace rb

using namespace rb;

// elements per CTA: the choice that keeps most lanes busy in both phases (a pass handles kBD/RS items; the last pass of a
// phase is partially filled)
static int choose_te(double tris_per_elem, double items_per_elem, int rs, double w_vol, double w_face) {
    const int it = kBD / rs;
    int best = 1;
    double best_u = -1.0;
    for (int te = 4; te <= 64; ++te) {
        const double na = te * tris_per_elem, nb = te * items_per_elem;
        const double ua = na > 0 ? na / (ceil(na / it) * it) : 1.0;
        const double ub = nb > 0 ? nb / (ceil(nb / it) * it) : 1.0;
        const double u = (w_vol * ua + w_face * ub) / (w_vol + w_face) - 0.002 * te / 64.0;   // ties -> smaller CTAs
        if (u > best_u) { best_u = u; best = te; }
    }
    return best;
}

extern "C" size_t reyna_b200_assemble_workspace(int32_t n_elem, int32_t n_iface) {
    return sizeof(ElemBox) * (size_t)(n_elem > 0 ? n_elem : 1) + sizeof(double) * (size_t)(n_iface > 0 ? n_iface : 1) + 256;
}

extern "C" int reyna_b200_assemble(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                   double sigma_D, double* data, double* load, long long* zero_count, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    RB_REQUIRE(g && s && c && q && data && load && zero_count && workspace, RB_E_ARG, "reyna_b200_assemble: null argument");
    RB_REQUIRE(workspace_bytes >= reyna_b200_assemble_workspace(g->n_elem, g->n_iface), RB_E_WORKSPACE,
               "reyna_b200_assemble: workspace too small");
    const int p = q->p;
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_assemble: polynomial degree must be 1..3");
    const bool diff = c->vol_a != nullptr, adv = c->vol_b != nullptr;
    RB_REQUIRE(diff || adv, RB_E_NOCOEF, "reyna_b200_assemble: no advection or diffusion specified");
    RB_REQUIRE(!diff || (c->if_a && c->if_a_mid) || g->n_iface == 0, RB_E_ARG,
               "reyna_b200_assemble: diffusion given without interior-face samples");
    RB_REQUIRE(!adv || (c->if_b && c->if_upwind) || g->n_iface == 0, RB_E_ARG,
               "reyna_b200_assemble: advection given without interior-face samples / upwind flags");
    if (g->n_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    AsmArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *g;
    A.adj_off = s->adj_off;
    A.adj_item = s->adj_item;
    A.grp_off = s->grp_off;
    A.c = *c;
    const int qv = (p + 1) * (p + 1), qe = p + 1;
    for (int i = 0; i < qv; ++i) {
        A.wv[i] = q->wv[i];
        A.rvx[i] = q->rv[i][0];
        A.rvy[i] = q->rv[i][1];
    }
    for (int i = 0; i < qe; ++i) {
        A.we[i] = q->we[i];
        A.xe[i] = q->xe[i];
    }
    A.lc = make_leg_const();
    A.sigma_pp = sigma_D * (double)(p * p);
    A.p2 = (double)(p * p);
    A.data = data;
    A.load = load;
    A.zero_count = reinterpret_cast<unsigned long long*>(zero_count);
    A.has_dups = s->n_dup_items > 0;
    // prepass tables (16-byte aligned slices of the workspace)
    uintptr_t w = ((uintptr_t)workspace + 15) & ~(uintptr_t)15;
    ElemBox* ebox = reinterpret_cast<ElemBox*>(w);
    double* fsigma = reinterpret_cast<double*>(w + sizeof(ElemBox) * (size_t)(g->n_elem > 0 ? g->n_elem : 1));
    A.ebox = ebox;
    A.fsigma = fsigma;
    {
        const int64_t work = g->n_elem > g->n_iface ? g->n_elem : g->n_iface;
        const int blocks = grid_for(work, 256, kNumSMs * 8);
        dg_prepare_kernel<<<blocks, 256, 0, st>>>(*g, (diff && g->n_iface > 0) ? c->if_a_mid : nullptr, A.sigma_pp, A.p2,
                                                 ebox, fsigma);
        RB_LAUNCH_CHECK("dg_prepare_kernel");
    }
    const int rs = p == 3 ? 2 : 1;
    const double items_per_elem = (double)s->n_items / g->n_rows;
    const double tris_per_elem = (double)g->n_tri / g->n_rows;
    const int d = (p + 1) * (p + 2) / 2;
    // relative cost of one item of each phase (FMA counts of the two contractions)
    const double w_vol = tris_per_elem * qv * (d * d * (diff ? 3.0 : 1.0) + 12.0 * d);
    const double w_face = items_per_elem * qe * (2.0 * d * d * (diff ? 2.0 : 1.0) + 30.0 * d);
    A.te = choose_te(tris_per_elem, items_per_elem, rs, w_vol, w_face);
    switch (p) {
        case 1: return dispatch_coefs<1, 1>(A, diff, adv, st);
        case 2: return dispatch_coefs<2, 1>(A, diff, adv, st);
        default: return dispatch_coefs<3, 2>(A, diff, adv, st);
    }
}
