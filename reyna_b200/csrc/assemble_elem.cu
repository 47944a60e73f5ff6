// Element-streaming DG assembly kernel: one thread (or RS column-split threads) walks ALL sub-triangles and ALL interior
// faces of one element, so the d x d diagonal block is accumulated in registers in the reference's order (triangles in
// order, main.py:315, then faces) without any cross-thread reduction, barrier or floating-point atomic.
//
// Replaces DGFEM._stiffness_matrix + localstiff (reyna/DGFEM/two_dimensional/main.py:275-327, full_assembly.py:9-77),
// DGFEM._interior_stiffness_matrix + int_localstiff (main.py:329-374, full_assembly.py:80-214) and the scipy COO->CSR
// summation (main.py:319-320, 371-372, 211) for all elements at once.
//
// Why this shape on B200: the FP64 accumulators (a whole 6 x 6 block at p = 2) cap the kernel at 8 warps per SM, so
// latency must be hidden inside the thread.  Every input of the NEXT item is therefore already on its way while the
// current one is contracted:
//   * coefficient samples (q-major arrays) stream through a per-thread cp.async ring in shared memory, RD points deep;
//   * sub-triangle vertices: indices are read one triangle ahead, the vertex coordinates follow by cp.async;
//   * faces: adjacency item + neighbour id are read two items ahead, the flattened face record (dg_prepare_kernel) and
//     the neighbour's basis scalings follow by cp.async one item ahead.
// The only dependent-load chains left are one per element and phase.  Elements of a CTA are counting-sorted by their
// number of faces so that the lanes of a warp run loops of (almost) equal length.
//   phase V  volume terms of all sub-triangles          -> acc (registers), load vector -> global
//   phase D  diagonal parts of all (element, face) items -> acc                          -> CSR values (diagonal block)
//   phase O  off-diagonal block of every item                                            -> CSR values (its final slot)
// For p = 1 phases D and O are fused (both blocks fit in registers).
#include "assemble_common.cuh"
#include <cuda_pipeline.h>

namespace rb {

namespace elem {

constexpr int kWarp = 32;          // elements per CTA: every warp is its own CTA, so a finished warp frees its slot at once
constexpr int kSortChunk = 128;    // elements sorted together by number of faces (lane balance vs locality)

template <int P, int RS>
struct Dims {
    static constexpr int D = (P + 1) * (P + 2) / 2;
    static constexpr int DD = D * D;
    static constexpr int CI = D / RS;          // block columns (and load rows) per thread
    static constexpr int NA = D * CI;          // block accumulators per thread
    static constexpr int QV = (P + 1) * (P + 1);
    static constexpr int QE = P + 1;
    static constexpr int NE = kWarp;           // elements per CTA
    static constexpr int BD = kWarp * RS;      // threads per CTA (RS column parts, one warp each)
    static constexpr int RD = QV < 9 ? QV : (QV == 9 ? 9 : 8);   // depth of the volume coefficient ring (points)
    static constexpr bool FUSE = NA <= 9;      // diagonal and off-diagonal block of a face in one sweep
    // 16-byte planes per thread.  volume: RD x {a0, a1, b, (c,f)} + 2 x 3 vertices ; faces: 2 x 4 record + 2 x 4 box +
    // QE x {a0, a1, b}
    static constexpr int PLV = RD * 4 + 6;
    static constexpr int PLF = 8 + 8 + QE * 3;
    static constexpr int PL = PLV > PLF ? PLV : PLF;
    static_assert(D % RS == 0, "column split must divide the local dimension");
};

// plane `pl` of thread `tid`: consecutive threads touch consecutive 16-byte words -> conflict-free
__device__ __forceinline__ double2* plane(double2* s, int pl, int tid) { return s + (size_t)pl * blockDim.x + tid; }

__device__ __forceinline__ int gid_of(const AsmArgs& A, int e) { return A.g.elem_gid ? A.g.elem_gid[e] : e; }

// ---- volume ring ------------------------------------------------------------------------------------------------------
template <bool DIFF, bool ADV>
__device__ __forceinline__ void issue_vol(const AsmArgs& A, double2* s, int tid, int slot, size_t m) {
    if (DIFF) {
        __pipeline_memcpy_async(plane(s, slot * 4 + 0, tid), reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m, 16);
        __pipeline_memcpy_async(plane(s, slot * 4 + 1, tid), reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m + 1, 16);
    }
    if (ADV) __pipeline_memcpy_async(plane(s, slot * 4 + 2, tid), reinterpret_cast<const double2*>(A.c.vol_b) + m, 16);
    double* cf = reinterpret_cast<double*>(plane(s, slot * 4 + 3, tid));
    if (A.c.vol_c) __pipeline_memcpy_async(cf, A.c.vol_c + m, 8);
    if (A.c.vol_f) __pipeline_memcpy_async(cf + 1, A.c.vol_f + m, 8);
}

__device__ __forceinline__ void issue_nodes(const AsmArgs& A, double2* s, int tid, int base_plane, int i0, int i1, int i2) {
    const double2* nd = reinterpret_cast<const double2*>(A.g.nodes);
    __pipeline_memcpy_async(plane(s, base_plane + 0, tid), nd + i0, 16);
    __pipeline_memcpy_async(plane(s, base_plane + 1, tid), nd + i1, 16);
    __pipeline_memcpy_async(plane(s, base_plane + 2, tid), nd + i2, 16);
}

// ---- face staging -----------------------------------------------------------------------------------------------------
template <bool DIFF, bool ADV>
__device__ __forceinline__ void issue_face_coef(const AsmArgs& A, double2* s, int tid, int base_plane, int q, size_t m) {
    if (DIFF) {
        __pipeline_memcpy_async(plane(s, base_plane + q * 3 + 0, tid), reinterpret_cast<const double2*>(A.c.if_a) + 2 * m, 16);
        __pipeline_memcpy_async(plane(s, base_plane + q * 3 + 1, tid), reinterpret_cast<const double2*>(A.c.if_a) + 2 * m + 1, 16);
    }
    if (ADV) __pipeline_memcpy_async(plane(s, base_plane + q * 3 + 2, tid), reinterpret_cast<const double2*>(A.c.if_b) + m, 16);
}

__device__ __forceinline__ void issue_rec(const AsmArgs& A, double2* s, int tid, int base_plane, int f) {
    const double2* src = reinterpret_cast<const double2*>(A.frec + f);
#pragma unroll
    for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(plane(s, base_plane + k, tid), src + k, 16);
}

__device__ __forceinline__ void issue_box(const AsmArgs& A, double2* s, int tid, int base_plane, int e) {
    const double2* src = reinterpret_cast<const double2*>(A.ebox + e);
#pragma unroll
    for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(plane(s, base_plane + k, tid), src + k, 16);
}

__device__ __forceinline__ ElemBox read_box(double2* s, int tid, int base_plane) {
    const double2 m = *plane(s, base_plane, tid), ih = *plane(s, base_plane + 1, tid);
    const double2 s05 = *plane(s, base_plane + 2, tid), s15 = *plane(s, base_plane + 3, tid);
    ElemBox b;
    b.mx = m.x; b.my = m.y; b.ihx = ih.x; b.ihy = ih.y; b.s05x = s05.x; b.s05y = s05.y; b.s15x = s15.x; b.s15y = s15.y;
    return b;
}


// One 6-double row of a block at a 16-byte aligned address: a 32-byte and a 16-byte store (STG.256 + STG.128), the 32-byte one
// on whichever half is sector aligned -- two global store operations per row instead of three (the kernel is bound by the
// number of per-lane global operations, profiles/r01_assemble_experiments.md).
__device__ __forceinline__ void store_row6(double* dst, double v0, double v1, double v2, double v3, double v4, double v5) {
    const bool al = (reinterpret_cast<size_t>(dst) & 31) == 0;
    double* p4 = al ? dst : dst + 2;
    double* p2 = al ? dst + 4 : dst;
    const double a0 = al ? v0 : v2, a1 = al ? v1 : v3, a2 = al ? v2 : v4, a3 = al ? v3 : v5;
    const double b0 = al ? v4 : v0, b1 = al ? v5 : v1;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p4), "d"(a0), "d"(a1), "d"(a2), "d"(a3) : "memory");
    *reinterpret_cast<double2*>(p2) = make_double2(b0, b1);
}

struct FaceGeo {
    double midx, midy, tanx, tany, nx, ny, De, sigma;
};

__device__ __forceinline__ FaceGeo read_rec(double2* s, int tid, int base_plane) {
    const double2 a = *plane(s, base_plane, tid), b = *plane(s, base_plane + 1, tid);
    const double2 c = *plane(s, base_plane + 2, tid), d = *plane(s, base_plane + 3, tid);
    FaceGeo g;
    g.midx = a.x; g.midy = a.y; g.tanx = b.x; g.tany = b.y; g.nx = c.x; g.ny = c.y; g.De = d.x; g.sigma = d.y;
    return g;
}

// One quadrature point of one face item: P_r = s phi_r, Q_r = -F_r/2, V_c = W s phi_c, U_c = (sigma + gamma) V_c - W F_c/2;
// B[r,c] += P_r U_c + Q_r V_c  (full_assembly.py:174-212; s = +1 on the K1 side, -1 on the K2 side).
template <int P, int RS, int PART, bool DIFF, bool ADV, bool DO_DIAG, bool DO_OFF>
__device__ __forceinline__ void face_point(const AsmArgs& A, const FaceGeo& g, double sgn, bool down, int q, double2* s,
                                           int tid, int coef_plane, const ElemBox& bxe, const ElemBox& bxo,
                                           double* __restrict__ dacc, double* __restrict__ oacc) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, CI = DM::CI, I0 = PART * CI;
    double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
    if (DIFF) {
        const double2 r0 = *plane(s, coef_plane + q * 3 + 0, tid), r1 = *plane(s, coef_plane + q * 3 + 1, tid);
        a00 = r0.x; a01 = r0.y; a10 = r1.x; a11 = r1.y;
    }
    if (ADV) {
        const double2 bb = *plane(s, coef_plane + q * 3 + 2, tid);
        b0 = bb.x; b1 = bb.y;
    }
    const double x = g.midx + A.xe[q] * g.tanx;
    const double y = g.midy + A.xe[q] * g.tany;
    const double W = g.De * A.we[q];
    double atn0 = 0, atn1 = 0, gamma = 0;
    if (DIFF) {
        // F_k = n . (a grad phi_k) = (a^T n) . grad phi_k
        atn0 = g.nx * a00 + g.ny * a10;
        atn1 = g.nx * a01 + g.ny * a11;
    }
    if (ADV) {
        const double beta = b0 * g.nx + b1 * g.ny;
        gamma = down ? -sgn * beta : 0.0;
    }
    const double hw0 = -0.5 * W * atn0, hw1 = -0.5 * W * atn1;     // -W/2 * a^T n
    const double sg = g.sigma + gamma;
    double phi[D], gx[D], gy[D];
    eval_basis<P, DIFF>(x, y, bxe, A.lc, phi, gx, gy);
    double Pr[D], Qr[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
        Pr[j] = sgn * phi[j];
        Qr[j] = 0.0;
        if (DIFF) {
            if (gx_nz<P>(j) && gy_nz<P>(j)) Qr[j] = -0.5 * (atn0 * gx[j] + atn1 * gy[j]);
            else if (gx_nz<P>(j)) Qr[j] = -0.5 * (atn0 * gx[j]);
            else if (gy_nz<P>(j)) Qr[j] = -0.5 * (atn1 * gy[j]);
        }
    }
    const double Ws = W * sgn;
    if (DO_DIAG) {
#pragma unroll
        for (int ii = 0; ii < CI; ++ii) {
            const int i = I0 + ii;
            const double Vc = Ws * phi[i];
            double Uc = sg * Vc;
            if (DIFF && gy_nz<P>(i)) Uc = fma(hw1, gy[i], Uc);
            if (DIFF && gx_nz<P>(i)) Uc = fma(hw0, gx[i], Uc);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double v = dacc[j * CI + ii];
                v = fma(Pr[j], Uc, v);
                if (DIFF && (gx_nz<P>(j) || gy_nz<P>(j))) v = fma(Qr[j], Vc, v);
                dacc[j * CI + ii] = v;
            }
        }
    }
    if (DO_OFF) {
        eval_basis<P, DIFF>(x, y, bxo, A.lc, phi, gx, gy);
#pragma unroll
        for (int ii = 0; ii < CI; ++ii) {
            const int i = I0 + ii;
            const double Vc = -Ws * phi[i];
            double Uc = sg * Vc;
            if (DIFF && gy_nz<P>(i)) Uc = fma(hw1, gy[i], Uc);
            if (DIFF && gx_nz<P>(i)) Uc = fma(hw0, gx[i], Uc);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double v = oacc[j * CI + ii];
                v = fma(Pr[j], Uc, v);
                if (DIFF && (gx_nz<P>(j) || gy_nz<P>(j))) v = fma(Qr[j], Vc, v);
                oacc[j * CI + ii] = v;
            }
        }
    }
}

// One pass over the (element, face) items of element e.  DO_DIAG: diagonal contributions into dacc (and the count of
// neighbours below e = slot of the diagonal block).  DO_OFF: the off-diagonal block of every neighbour group, stored to its
// CSR slot.  Staging planes: record double buffer [0,8), neighbour box double buffer [8,16), coefficient ring [16, 16+3 QE).
template <int P, int RS, int PART, bool DIFF, bool ADV, bool DO_DIAG, bool DO_OFF>
__device__ __forceinline__ void face_pass(const AsmArgs& A, int e, int eg, const ElemBox& bxe, int k0, int k1, double2* s,
                                          int tid, double* __restrict__ dacc, int& dslot, int64_t row_base, int rowlen,
                                          unsigned& zeros) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, CI = DM::CI, I0 = PART * CI, NA = DM::NA, QE = DM::QE;
    if (k0 >= k1) return;
    const size_t F = (size_t)A.g.n_iface;
    // prologue: items k0 (staged now) and k0+1 (registers)
    int item = A.adj_item[k0];
    int nbr = A.adj_nb[k0];
    int item1 = k0 + 1 < k1 ? A.adj_item[k0 + 1] : 0;
    int nbr1 = k0 + 1 < k1 ? A.adj_nb[k0 + 1] : 0;
    {
        const int f = (item & 0x7fffffff) >> 1;
        issue_rec(A, s, tid, 0, f);
        if (DO_OFF) issue_box(A, s, tid, 8, nbr);
        __pipeline_commit();
#pragma unroll
        for (int q = 0; q < QE; ++q) {
            issue_face_coef<DIFF, ADV>(A, s, tid, 16, q, (size_t)q * F + f);
            __pipeline_commit();
        }
    }
    unsigned char upw = A.c.if_upwind ? A.c.if_upwind[(item & 0x7fffffff) >> 1] : 0;
    double oacc[DO_OFF ? NA : 1];
    if (DO_OFF) {
#pragma unroll
        for (int v = 0; v < NA; ++v) oacc[v] = 0.0;
    }
    int gi = 0;                                   // neighbour groups completed so far
    for (int k = k0; k < k1; ++k) {
        const int buf = (k - k0) & 1;
        const int f = (item & 0x7fffffff) >> 1;
        const int side = item & 1;
        const bool has1 = k + 1 < k1;
        // stage item k+1 (record, neighbour box), read item k+2 (registers)
        const int f1 = (item1 & 0x7fffffff) >> 1;
        unsigned char upw1 = 0;
        if (has1) {
            issue_rec(A, s, tid, (buf ^ 1) * 4, f1);
            if (DO_OFF) issue_box(A, s, tid, 8 + (buf ^ 1) * 4, nbr1);
            if (A.c.if_upwind) upw1 = A.c.if_upwind[f1];
        }
        __pipeline_commit();
        int item2 = 0, nbr2 = 0;
        if (k + 2 < k1) {
            item2 = A.adj_item[k + 2];
            nbr2 = A.adj_nb[k + 2];
        }
        const double sgn = side ? -1.0 : 1.0;
        const bool down = A.c.if_upwind ? ((upw != 0) == (side == 1)) : false;
        FaceGeo g;
        ElemBox bxo;
#pragma unroll
        for (int q = 0; q < QE; ++q) {
            // groups newer than point q of this item: the staging group above, q refills, QE-1-q older refills  -> QE
            __pipeline_wait_prior(QE);
            if (q == 0) {
                g = read_rec(s, tid, buf * 4);
                if (DO_OFF) bxo = read_box(s, tid, 8 + buf * 4);
            }
            face_point<P, RS, PART, DIFF, ADV, DO_DIAG, DO_OFF>(A, g, sgn, down, q, s, tid, 16, bxe, DO_OFF ? bxo : bxe, dacc,
                                                                oacc);
            if (has1) issue_face_coef<DIFF, ADV>(A, s, tid, 16, q, (size_t)q * F + f1);   // slot q is free again
            __pipeline_commit();
        }
        const bool last_of_group = !has1 || item1 >= 0;     // the next item starts another neighbour
        if (last_of_group) {
            const int og = gid_of(A, nbr);
            if (DO_DIAG) dslot += (og < eg);
            if (DO_OFF) {
                const int slot = gi + (og > eg ? 1 : 0);
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double* dst = A.data + row_base + (int64_t)j * rowlen + (int64_t)slot * D + I0;
                    if (RS == 1 && D == 6) {
                        store_row6(dst, oacc[j * 6], oacc[j * 6 + 1], oacc[j * 6 + 2], oacc[j * 6 + 3], oacc[j * 6 + 4], oacc[j * 6 + 5]);
#pragma unroll
                        for (int i = 0; i < 6; ++i) zeros += (oacc[j * 6 + i] == 0.0);
                    } else if (RS == 1 && (D & 1) == 0) {
#pragma unroll
                        for (int i = 0; i < D; i += 2) {
                            reinterpret_cast<double2*>(dst)[i >> 1] = make_double2(oacc[j * CI + i], oacc[j * CI + i + 1]);
                            zeros += (oacc[j * CI + i] == 0.0) + (oacc[j * CI + i + 1] == 0.0);
                        }
                    } else {
#pragma unroll
                        for (int ii = 0; ii < CI; ++ii) {
                            dst[ii] = oacc[j * CI + ii];
                            zeros += (oacc[j * CI + ii] == 0.0);
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < NA; ++v) oacc[v] = 0.0;
            }
            ++gi;
        }
        item = item1; nbr = nbr1; upw = upw1;
        item1 = item2; nbr1 = nbr2;
    }
    __pipeline_wait_prior(0);
}

template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void element_work(const AsmArgs& A, int e, double2* s, int tid, unsigned& zeros) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, DD = DM::DD, CI = DM::CI, I0 = PART * CI, NA = DM::NA, QV = DM::QV, RD = DM::RD;
    const ElemBox bx = A.ebox[e];
    const int t0 = A.g.tri_off[e], t1 = A.g.tri_off[e + 1];
    const int k0 = A.adj_off[e], k1 = A.adj_off[e + 1];
    const int g0 = A.grp_off[e], g1 = A.grp_off[e + 1];
    const int eg = gid_of(A, e);
    const bool has_c = A.c.vol_c != nullptr, has_f = A.c.vol_f != nullptr;
    const size_t T = (size_t)A.g.n_tri;

    double acc[NA + CI];
#pragma unroll
    for (int v = 0; v < NA + CI; ++v) acc[v] = 0.0;

    // ---- phase V: volume terms, triangles in order ---------------------------------------------------------------------
    if (t0 < t1) {
        const int total = (t1 - t0) * QV;         // stream of (triangle, point) pairs
        int i0 = A.g.tri[3 * (size_t)t0], i1 = A.g.tri[3 * (size_t)t0 + 1], i2 = A.g.tri[3 * (size_t)t0 + 2];
        // prologue: the first RD points; the vertices of the first triangle ride in the first group
        issue_nodes(A, s, tid, RD * 4, i0, i1, i2);
#pragma unroll
        for (int r = 0; r < RD; ++r) {
            if (r < total) {
                const int tt = t0 + r / QV, qq = r % QV;
                issue_vol<DIFF, ADV>(A, s, tid, r, (size_t)qq * T + tt);
            }
            __pipeline_commit();
        }
        int sidx = 0;                             // stream index of the point being consumed
        for (int t = t0; t < t1; ++t) {
            const int nbuf = (t - t0) & 1;
            double2 V0, V1, V2;
            double hdet = 0.0;
#pragma unroll 1
            for (int q = 0; q < QV; ++q, ++sidx) {
                if (q == 0) {
                    // vertices of this triangle were committed >= QV - QV/2 - 1 groups ago (or in the prologue)
                    __pipeline_wait_prior(QV - QV / 2 - 1 < RD - 1 ? QV - QV / 2 - 1 : RD - 1);
                    V0 = *plane(s, RD * 4 + nbuf * 3 + 0, tid);
                    V1 = *plane(s, RD * 4 + nbuf * 3 + 1, tid);
                    V2 = *plane(s, RD * 4 + nbuf * 3 + 2, tid);
                    // |det(0.5*[V1-V0; V2-V0])| (full_assembly.py:40-41); the extra 0.5 is the `0.5 * z` of :75-77
                    const double e1x = 0.5 * (V1.x - V0.x), e1y = 0.5 * (V1.y - V0.y);
                    const double e2x = 0.5 * (V2.x - V0.x), e2y = 0.5 * (V2.y - V0.y);
                    hdet = 0.5 * fabs(e1x * e2y - e1y * e2x);
                    if (t + 1 < t1) {             // indices of the next triangle (vertices follow half a triangle later)
                        i0 = A.g.tri[3 * (size_t)(t + 1)];
                        i1 = A.g.tri[3 * (size_t)(t + 1) + 1];
                        i2 = A.g.tri[3 * (size_t)(t + 1) + 2];
                    }
                } else {
                    __pipeline_wait_prior(RD - 1);
                }
                const int slot = sidx % RD;
                double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
                const double r0 = A.rvx[q], r1 = A.rvy[q];
                const double l0 = 1.0 - r0 - r1;
                const double x = l0 * V0.x + r0 * V1.x + r1 * V2.x;
                const double y = l0 * V0.y + r0 * V1.y + r1 * V2.y;
                const double W = hdet * A.wv[q];
                if (DIFF) {
                    const double2 ra = *plane(s, slot * 4 + 0, tid), rb = *plane(s, slot * 4 + 1, tid);
                    a00 = W * ra.x; a01 = W * ra.y; a10 = W * rb.x; a11 = W * rb.y;
                }
                if (ADV) {
                    const double2 bb = *plane(s, slot * 4 + 2, tid);
                    b0 = W * bb.x; b1 = W * bb.y;
                }
                const double2 cf = *plane(s, slot * 4 + 3, tid);
                const double cq = has_c ? W * cf.x : 0.0;
                const double wf = has_f ? W * cf.y : 0.0;
                // refill this slot with the point RD ahead in the stream; the next triangle's vertices ride along
                {
                    const int nx = sidx + RD;
                    if (nx < total) {
                        const int tt = t0 + nx / QV, qq = nx % QV;
                        issue_vol<DIFF, ADV>(A, s, tid, slot, (size_t)qq * T + tt);
                    }
                    if (q == QV / 2 && t + 1 < t1) issue_nodes(A, s, tid, RD * 4 + (nbuf ^ 1) * 3, i0, i1, i2);
                    __pipeline_commit();
                }
                double phi[D], gx[D], gy[D];
                eval_basis<P, true>(x, y, bx, A.lc, phi, gx, gy);
                // column (trial i) quantities u = W * grad(phi_i)^T a, s = W * (b.grad(phi_i) + c phi_i)
#pragma unroll
                for (int ii = 0; ii < CI; ++ii) {
                    const int i = I0 + ii;
                    const bool nzx = gx_nz<P>(i), nzy = gy_nz<P>(i);     // compile-time after unrolling
                    double u0 = 0, u1 = 0;
                    if (DIFF) {
                        if (nzx && nzy) {
                            u0 = fma(gx[i], a00, gy[i] * a10);
                            u1 = fma(gx[i], a01, gy[i] * a11);
                        } else if (nzx) {
                            u0 = gx[i] * a00;
                            u1 = gx[i] * a01;
                        } else if (nzy) {
                            u0 = gy[i] * a10;
                            u1 = gy[i] * a11;
                        }
                    }
                    double si = cq * phi[i];
                    if (ADV && nzy) si = fma(b1, gy[i], si);
                    if (ADV && nzx) si = fma(b0, gx[i], si);
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        double v = acc[j * CI + ii];
                        if (DIFF && (nzx || nzy)) {
                            if (gx_nz<P>(j)) v = fma(u0, gx[j], v);
                            if (gy_nz<P>(j)) v = fma(u1, gy[j], v);
                        }
                        v = fma(si, phi[j], v);
                        acc[j * CI + ii] = v;
                    }
                }
                if (has_f) {
#pragma unroll
                    for (int ii = 0; ii < CI; ++ii) acc[NA + ii] = fma(wf, phi[I0 + ii], acc[NA + ii]);
                }
            }
        }
        __pipeline_wait_prior(0);
    }
    // load vector rows of this part
    if (RS == 1 && D == 6) {
        store_row6(A.load + (size_t)e * 6, acc[NA], acc[NA + 1], acc[NA + 2], acc[NA + 3], acc[NA + 4], acc[NA + 5]);
    } else {
#pragma unroll
        for (int ii = 0; ii < CI; ++ii) A.load[(size_t)e * D + I0 + ii] = acc[NA + ii];
    }

    // ---- phases D / O: interior faces ---------------------------------------------------------------------------------
    const int nb = (g1 - g0) + 1;
    const int rowlen = nb * D;
    const int64_t row_base = (int64_t)DD * (g0 + e);
    int dslot = 0;
    if (DM::FUSE) {
        face_pass<P, RS, PART, DIFF, ADV, true, true>(A, e, eg, bx, k0, k1, s, tid, acc, dslot, row_base, rowlen, zeros);
    } else {
        face_pass<P, RS, PART, DIFF, ADV, true, false>(A, e, eg, bx, k0, k1, s, tid, acc, dslot, row_base, rowlen, zeros);
    }
    // diagonal block: CI consecutive columns of every row
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double* dst = A.data + row_base + (int64_t)j * rowlen + (int64_t)dslot * D + I0;
        if (RS == 1 && D == 6) {
            store_row6(dst, acc[j * 6], acc[j * 6 + 1], acc[j * 6 + 2], acc[j * 6 + 3], acc[j * 6 + 4], acc[j * 6 + 5]);
#pragma unroll
            for (int i = 0; i < 6; ++i) zeros += (acc[j * 6 + i] == 0.0);
        } else if (RS == 1 && (D & 1) == 0) {
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                reinterpret_cast<double2*>(dst)[i >> 1] = make_double2(acc[j * CI + i], acc[j * CI + i + 1]);
                zeros += (acc[j * CI + i] == 0.0) + (acc[j * CI + i + 1] == 0.0);
            }
        } else {
#pragma unroll
            for (int ii = 0; ii < CI; ++ii) {
                dst[ii] = acc[j * CI + ii];
                zeros += (acc[j * CI + ii] == 0.0);
            }
        }
    }
    if (!DM::FUSE) {
        int unused = 0;
        face_pass<P, RS, PART, DIFF, ADV, false, true>(A, e, eg, bx, k0, k1, s, tid, acc, unused, row_base, rowlen, zeros);
    }
}

// lane balance: elements are counting-sorted by their number of face items inside chunks of kSortChunk consecutive
// elements (descending), so that the 32 elements of a warp run loops of (almost) equal length
__global__ void __launch_bounds__(kSortChunk) dg_sort_kernel(int n_rows, const int32_t* __restrict__ adj_off,
                                                             int32_t* __restrict__ perm) {
    __shared__ int s_hist[32], s_start[32];
    const int tid = threadIdx.x;
    const int e = blockIdx.x * kSortChunk + tid;
    if (tid < 32) s_hist[tid] = 0;
    __syncthreads();
    int key = 0, rank = 0;
    if (e < n_rows) {
        const int cnt = adj_off[e + 1] - adj_off[e];
        key = 31 - min(cnt, 31);
        rank = atomicAdd(&s_hist[key], 1);      // the rank inside a bin only decides WHICH lane computes an element
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int b = 0; b < 32; ++b) { s_start[b] = run; run += s_hist[b]; }
    }
    __syncthreads();
    if (e < n_rows) perm[blockIdx.x * kSortChunk + s_start[key] + rank] = e;
}

template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kWarp * RS, 8 / RS) dg_assemble_elem_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dims<P, RS>;
    extern __shared__ __align__(16) double2 smem2[];
    const int tid = threadIdx.x;
    const int part = tid / kWarp;       // warp-uniform column part
    const int le = tid - part * kWarp;
    const int idx = blockIdx.x * kWarp + le;
    if (idx >= A.g.n_rows) return;
    const int e = A.perm[idx];
    unsigned zeros = 0;
    if (RS == 1 || part == 0) element_work<P, RS, 0, DIFF, ADV>(A, e, smem2, tid, zeros);
    else element_work<P, RS, RS - 1, DIFF, ADV>(A, e, smem2, tid, zeros);
    if (zeros) atomicAdd(A.zero_count, (unsigned long long)zeros);
}

template <int P, int RS, bool DIFF, bool ADV>
static int launch(const AsmArgs& A, cudaStream_t st) {
    using DM = Dims<P, RS>;
    const size_t smem = sizeof(double2) * (size_t)DM::PL * DM::BD;
    auto kern = dg_assemble_elem_kernel<P, RS, DIFF, ADV>;
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dg_sort_kernel<<<(int)div_up(A.g.n_rows, kSortChunk), kSortChunk, 0, st>>>(A.g.n_rows, A.adj_off, A.perm);
    RB_LAUNCH_CHECK("dg_sort_kernel");
    const int blocks = (int)div_up(A.g.n_rows, kWarp);
    kern<<<blocks, DM::BD, smem, st>>>(A);
    RB_LAUNCH_CHECK("dg_assemble_elem_kernel");
    return 0;
}

template <int P, int RS>
static int dispatch(const AsmArgs& A, bool diff, bool adv, cudaStream_t st) {
    if (diff && adv) return launch<P, RS, true, true>(A, st);
    if (diff) return launch<P, RS, true, false>(A, st);
    return launch<P, RS, false, true>(A, st);
}

}  // namespace elem

int launch_assemble_elem(const AsmArgs& A, int p, bool diff, bool adv, cudaStream_t st) {
    switch (p) {
        case 1: return elem::dispatch<1, 1>(A, diff, adv, st);
        case 2: return elem::dispatch<2, 1>(A, diff, adv, st);
        default: return elem::dispatch<3, 2>(A, diff, adv, st);
    }
}

}  // namespace rb
