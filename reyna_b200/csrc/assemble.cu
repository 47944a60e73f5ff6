// The fused DG assembly kernel: volume terms + interior-face terms -> CSR values, one launch, no atomics on doubles.
//
// Replaces, for all elements at once,
//   DGFEM._stiffness_matrix + localstiff            reyna/DGFEM/two_dimensional/main.py:275-327,
//                                                   _auxilliaries/assembly/full_assembly.py:9-77
//   DGFEM._interior_stiffness_matrix + int_localstiff   main.py:329-374, full_assembly.py:80-214
//   the scipy COO->CSR conversions and `B += ...`    main.py:319-320, 371-372, 211
//
// Design (B200): a CTA owns TE consecutive elements, i.e. TE consecutive block rows of B.
//   phase A  sub-triangles: Legendre basis + gradients in registers, the d x d local block (and the load) accumulated in
//            registers over the (p+1)^2 points, then summed per element IN TRIANGLE ORDER through a shared-memory stage
//            (the reference's `s[:, elem] += ...` order, main.py:315).
//   phase B  (element, interior face) items of the element->face adjacency: the diagonal part B[K,K] of the SIPG/upwind
//            face matrix goes through the same staged per-element sum, the off-diagonal part B[K,K'] is complete after
//            one face and is stored straight into its final CSR position.  Each face is therefore visited from both of
//            its elements, each side producing only its own block rows: no intermediate face-matrix buffer, no
//            scatter-add, bitwise reproducible.
//   phase C  the summed diagonal blocks and load are written to the CSR value array / load vector.
// An item is processed by RS threads, each owning d/RS COLUMNS (trial functions) of the blocks: the FP64 accumulators are
// what limits occupancy (8 warps per SM with a whole 6 x 6 block per thread), and a column split only duplicates the
// basis evaluation -- the per-column quantities (a grad phi_i, b.grad phi_i, ...) are computed once.  The RS parts of an
// item sit in different warps (part = tid / IT), so the compile-time column offset stays warp-uniform.
// Coefficient samples are q-major ([q][item]); they travel global -> shared with per-thread cp.async so that all points
// of an item are in flight at once without costing registers.
#include "assemble_common.cuh"
#include <cuda_pipeline.h>
#include <stdlib.h>

namespace rb {

constexpr int kBD = 128;   // threads per CTA
#ifndef RB_MINB
#define RB_MINB 3
#endif
constexpr int kFuseLimit = 18;   // accumulators per block up to which a face item forms both of its blocks in one sweep

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
                                                         double* __restrict__ fsigma, FaceRec* __restrict__ frec) {
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.n_elem; e += stride) ebox[e] = make_box(g.bbox, e);
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.n_iface; f += stride) {
        const double2 P0 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f]];
        const double2 P1 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f + 1]];
        const double2 n = reinterpret_cast<const double2*>(g.inorm)[f];
        const double ex = P1.x - P0.x, ey = P1.y - P0.y;
        const double tx = 0.5 * ex, ty = 0.5 * ey;
        const double De = sqrt(tx * tx + ty * ty);
        double sigma = 0.0;
        if (if_a_mid != nullptr) {
            const int e1 = g.ielem[2 * (size_t)f], e2 = g.ielem[2 * (size_t)f + 1];
            const double2 m0 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f];
            const double2 m1 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f + 1];
            // lambda = n . a(mid) . n   (full_assembly.py:140)
            const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;
            const double a1 = g.area[e1], a2 = g.area[e2];
            const double c1 = fmin(a1 / max_sub_area_g(g, e1, P0, ex, ey), p2);
            const double c2 = fmin(a2 / max_sub_area_g(g, e2, P0, ex, ey), p2);
            sigma = sigma_pp * lam * (2.0 * De) * fmax(c1 / a1, c2 / a2);   // full_assembly.py:146-148
            fsigma[f] = sigma;
        }
        FaceRec r;
        r.midx = (P0.x + P1.x) * 0.5; r.midy = (P0.y + P1.y) * 0.5;     // same expressions as face_setup
        r.tanx = tx; r.tany = ty;
        r.nx = n.x; r.ny = n.y;
        r.De = De; r.sigma = sigma;
        frec[f] = r;
    }
}


template <int P, int RS>
struct Dims {
    static constexpr int D = (P + 1) * (P + 2) / 2;
    static constexpr int DD = D * D;
    static constexpr int NV = DD + D;       // per-element accumulators in shared memory: block + load
    static constexpr int CI = D / RS;       // block columns (and load rows) per thread
    static constexpr int NA = D * CI;       // block accumulators per thread
    static constexpr int QV = (P + 1) * (P + 1);
    static constexpr int QE = P + 1;
    static constexpr int IT = kBD / RS;     // items per pass
    static constexpr int NTV = NA + CI;     // values a thread stages after a volume pass
    static constexpr int CH = NTV > 24 ? ((NTV + 1) / 2 > 28 ? 22 : (NTV + 1) / 2) : NTV;   // staged values per round
    // doubles of one staged item: (a 4, b 2, c 1, f 1) per volume point, (a 4, b 2) per face point
    static constexpr int CV = QV * 8, CF = QE * 6;
    static constexpr int COEF = IT * (CV > CF ? CV : CF);            // doubles of the coefficient stage
    static constexpr int BOX = IT * 16;                              // + two ElemBox per item (own / other element)
    static constexpr int UN = (CH * (kBD + 1)) > (COEF + BOX) ? (CH * (kBD + 1)) : (COEF + BOX);   // union with the summation stage
    static_assert(D % RS == 0, "column split must divide the local dimension");
};

__host__ __device__ inline size_t asm_smem_bytes(int nv, int un, int te) {
    // s_acc [te][nv] doubles | union { s_stage [ch][kBD+1], s_coef [IT][max(CV,CF)] } doubles | 3 x [te+1] ints + [te] ints
    return sizeof(double) * ((size_t)te * nv + (size_t)un) + sizeof(int) * (4 * ((size_t)te + 1));
}

// ---------------------------------------------------------------------------------------------------------------
// staged, ordered per-element sum of per-thread register vectors
// ---------------------------------------------------------------------------------------------------------------
// vals[0 .. NT) of the thread handling item (base + it) with column part `part`; items of element seg are
// [off[seg], off[seg+1]).  Block value vv = j*CI + ii (< NBLK) lands at s_acc[seg*NV + j*D + part*CI + ii], tail value
// NBLK + ii (load entry) at s_acc[seg*NV + D*D + part*CI + ii].  The sum over an element's items runs in item order.
template <int NT, int NBLK, int CH, int IT, int RS, int D>
__device__ __forceinline__ void staged_sum(const double (&vals)[NT], bool active, int base, int ne,
                                           const int* __restrict__ s_off, double* __restrict__ s_acc,
                                           double* __restrict__ s_stage) {
    constexpr int NCH = (NT + CH - 1) / CH;
    constexpr int CI = D / RS, NV = D * D + D;
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
                int dst;
                if (vv < NBLK) {
                    const int j = vv / CI, ii = vv - j * CI;
                    dst = seg * NV + j * D + part * CI + ii;
                } else {
                    dst = seg * NV + D * D + part * CI + (vv - NBLK);
                }
                double s = s_acc[dst];
                const double* src = s_stage + v * (kBD + 1) + part * IT;
                // loads batched four at a time (independent of the running sum), additions strictly in item order
                int t = lo;
                for (; t + 4 <= hi; t += 4) {
                    const double x0 = src[t], x1 = src[t + 1], x2 = src[t + 2], x3 = src[t + 3];
                    s += x0; s += x1; s += x2; s += x3;
                }
                for (; t < hi; ++t) s += src[t];
                s_acc[dst] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// phase A: one sub-triangle
// ---------------------------------------------------------------------------------------------------------------
// Coefficient samples travel global -> shared with per-thread cp.async (LDGSTS): all (p+1)^2 points of an item are in
// flight at once and cost no registers.  Stage layout per point q:
// [a row0 (16 B) x IT][a row1 x IT][b x IT][c (8 B) x IT][f x IT] -- every shared access of a warp is contiguous.
template <int P, int RS, bool DIFF, bool ADV>
__device__ __forceinline__ void stage_vol(const AsmArgs& A, double* __restrict__ s_coef, int lt, int part, int t) {
    using DM = Dims<P, RS>;
    constexpr int IT = DM::IT;
    const size_t T = (size_t)A.g.n_tri;
#pragma unroll
    for (int q0 = 0; q0 < DM::QV; q0 += RS) {
        const int q = q0 + part;                 // the parts of an item stage alternating points
        if (q < DM::QV) {
            const size_t m = (size_t)q * T + t;
            double2* blk = reinterpret_cast<double2*>(s_coef + (size_t)q * IT * 8);
            if (DIFF) {
                __pipeline_memcpy_async(blk + lt, reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m, 16);
                __pipeline_memcpy_async(blk + IT + lt, reinterpret_cast<const double2*>(A.c.vol_a) + 2 * m + 1, 16);
            }
            if (ADV) __pipeline_memcpy_async(blk + 2 * IT + lt, reinterpret_cast<const double2*>(A.c.vol_b) + m, 16);
            double* sc = s_coef + (size_t)q * IT * 8 + 6 * IT;
            if (A.c.vol_c) __pipeline_memcpy_async(sc + lt, A.c.vol_c + m, 8);
            if (A.c.vol_f) __pipeline_memcpy_async(sc + IT + lt, A.c.vol_f + m, 8);
        }
    }
}

// The bounding-box scalings of an item's element(s) are staged next to the coefficients (8 x 16 B planes of IT entries:
// planes 0-3 own element, 4-7 other element) and re-read from shared memory at every point instead of occupying 16-32
// registers for the whole item.
template <int IT>
__device__ __forceinline__ void stage_box(const AsmArgs& A, double* __restrict__ s_box, int lt, int which, int e) {
    const double2* src = reinterpret_cast<const double2*>(A.ebox + e);
    double2* dst = reinterpret_cast<double2*>(s_box) + (size_t)(4 * which) * IT + lt;
#pragma unroll
    for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(dst + (size_t)k * IT, src + k, 16);
}

template <int IT, bool KEEP_IN_SMEM>
__device__ __forceinline__ ElemBox staged_box(const double* __restrict__ s_box, int lt, int which) {
    ElemBox b;
    if (KEEP_IN_SMEM) {
        // volatile: the reads stay inside the point loop (hoisted, the box would occupy 16 registers for the whole item)
        const volatile double* src = s_box + 2 * ((size_t)(4 * which) * IT + lt);
        b.mx = src[0]; b.my = src[1];
        b.ihx = src[2 * IT]; b.ihy = src[2 * IT + 1];
        b.s05x = src[4 * IT]; b.s05y = src[4 * IT + 1];
        b.s15x = src[6 * IT]; b.s15y = src[6 * IT + 1];
    } else {
        const double2* src = reinterpret_cast<const double2*>(s_box) + (size_t)(4 * which) * IT + lt;
        const double2 m = src[0], ih = src[IT], s05 = src[2 * IT], s15 = src[3 * IT];
        b.mx = m.x; b.my = m.y; b.ihx = ih.x; b.ihy = ih.y; b.s05x = s05.x; b.s05y = s05.y; b.s15x = s15.x; b.s15y = s15.y;
    }
    return b;
}

struct TriGeom {
    double2 V0, V1, V2;
    double hdet;
    ElemBox bx;      // register copy (whole-block threads, RS == 1); column-split parts read the staged copy instead
};

__device__ __forceinline__ TriGeom load_tri_geom(const AsmArgs& A, int t) {
    TriGeom g;
    const int i0 = A.g.tri[3 * (size_t)t], i1 = A.g.tri[3 * (size_t)t + 1], i2 = A.g.tri[3 * (size_t)t + 2];
    g.V0 = reinterpret_cast<const double2*>(A.g.nodes)[i0];
    g.V1 = reinterpret_cast<const double2*>(A.g.nodes)[i1];
    g.V2 = reinterpret_cast<const double2*>(A.g.nodes)[i2];
    // |det(0.5*[V1-V0; V2-V0])| (full_assembly.py:40-41); the extra 0.5 is the `0.5 * z` of :75-77
    const double e1x = 0.5 * (g.V1.x - g.V0.x), e1y = 0.5 * (g.V1.y - g.V0.y);
    const double e2x = 0.5 * (g.V2.x - g.V0.x), e2y = 0.5 * (g.V2.y - g.V0.y);
    g.hdet = 0.5 * fabs(e1x * e2y - e1y * e2x);
    return g;
}

// acc[j*CI + ii] += sum_q W [ grad(phi_i)^T a grad(phi_j) + (b.grad(phi_i) + c phi_i) phi_j ],  i = PART*CI + ii  (column,
// trial), j = row (test): B[(e,j),(e,i)] of full_assembly.py:56-70 with main.py:294-299.  acc[NA + ii] += sum_q W f phi_i.
template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void triangle_block(const AsmArgs& A, const TriGeom& g, const double* __restrict__ s_coef,
                                               const double* __restrict__ s_box, int lt, double (&acc)[Dims<P, RS>::NTV]) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, CI = DM::CI, I0 = PART * CI, NA = DM::NA, IT = DM::IT;
    const bool has_c = A.c.vol_c != nullptr, has_f = A.c.vol_f != nullptr;
#pragma unroll 1
    for (int q = 0; q < DM::QV; ++q) {
        const double r0 = A.rvx[q], r1 = A.rvy[q];
        const double l0 = 1.0 - r0 - r1;
        const double x = l0 * g.V0.x + r0 * g.V1.x + r1 * g.V2.x;
        const double y = l0 * g.V0.y + r0 * g.V1.y + r1 * g.V2.y;
        double phi[D], gx[D], gy[D];
        eval_basis<P, true>(x, y, RS > 1 ? staged_box<IT, true>(s_box, lt, 0) : g.bx, A.lc, phi, gx, gy);
        const double W = g.hdet * A.wv[q];
        // W folded into the coefficients once per point
        const double2* blk = reinterpret_cast<const double2*>(s_coef + (size_t)q * IT * 8);
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
        if (DIFF) {
            const double2 r0a = blk[lt], r1a = blk[IT + lt];
            a00 = W * r0a.x; a01 = W * r0a.y; a10 = W * r1a.x; a11 = W * r1a.y;
        }
        if (ADV) {
            const double2 bb = blk[2 * IT + lt];
            b0 = W * bb.x; b1 = W * bb.y;
        }
        const double* sc = s_coef + (size_t)q * IT * 8 + 6 * IT;
        const double cq = has_c ? W * sc[lt] : 0.0;
        // column (trial i) quantities u = W * grad(phi_i)^T a, s = W * (b.grad(phi_i) + c phi_i)
#pragma unroll
        for (int ii = 0; ii < CI; ++ii) {
            constexpr int dummy = 0; (void)dummy;
            const int i = I0 + ii;
            double u0 = 0, u1 = 0;
            if (DIFF) {
                u0 = fma(gx[i], a00, gy[i] * a10);
                u1 = fma(gx[i], a01, gy[i] * a11);
            }
            double si = cq * phi[i];
            if (ADV) si = fma(b0, gx[i], fma(b1, gy[i], si));
#pragma unroll
            for (int j = 0; j < D; ++j) {
                double v = acc[j * CI + ii];
                if (DIFF) {
                    v = fma(u0, gx[j], v);
                    v = fma(u1, gy[j], v);
                }
                v = fma(si, phi[j], v);
                acc[j * CI + ii] = v;
            }
        }
        if (has_f) {
            const double wf = W * sc[IT + lt];
#pragma unroll
            for (int ii = 0; ii < CI; ++ii) acc[NA + ii] = fma(wf, phi[I0 + ii], acc[NA + ii]);
        }
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

// straight from global memory (only for the merged "same neighbour" followers of a group, which are not staged)
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

// stage layout per face point q: [a row0 x IT][a row1 x IT][b x IT] (16 B each)
template <int P, int RS, bool DIFF, bool ADV>
__device__ __forceinline__ void stage_face(const AsmArgs& A, double* __restrict__ s_coef, int lt, int part, int f) {
    using DM = Dims<P, RS>;
    constexpr int IT = DM::IT;
    const size_t F = (size_t)A.g.n_iface;
#pragma unroll
    for (int q0 = 0; q0 < DM::QE; q0 += RS) {
        const int q = q0 + part;
        if (q < DM::QE) {
            const size_t m = (size_t)q * F + f;
            double2* blk = reinterpret_cast<double2*>(s_coef + (size_t)q * IT * 6);
            if (DIFF) {
                __pipeline_memcpy_async(blk + lt, reinterpret_cast<const double2*>(A.c.if_a) + 2 * m, 16);
                __pipeline_memcpy_async(blk + IT + lt, reinterpret_cast<const double2*>(A.c.if_a) + 2 * m + 1, 16);
            }
            if (ADV) __pipeline_memcpy_async(blk + 2 * IT + lt, reinterpret_cast<const double2*>(A.c.if_b) + m, 16);
        }
    }
}

template <int IT, bool DIFF, bool ADV>
__device__ __forceinline__ FaceCoef staged_face_coef(const double* __restrict__ s_coef, int lt, int q) {
    FaceCoef k;
    k.a00 = k.a01 = k.a10 = k.a11 = k.b0 = k.b1 = 0.0;
    const double2* blk = reinterpret_cast<const double2*>(s_coef + (size_t)q * IT * 6);
    if (DIFF) {
        const double2 ar0 = blk[lt], ar1 = blk[IT + lt];
        k.a00 = ar0.x; k.a01 = ar0.y; k.a10 = ar1.x; k.a11 = ar1.y;
    }
    if (ADV) {
        const double2 bb = blk[2 * IT + lt];
        k.b0 = bb.x; k.b1 = bb.y;
    }
    return k;
}

// Accumulate, for columns [I0, I0+CI) and all rows of element e, the own-element (diagonal) block into dacc and/or the
// other-element block into oacc.  With P_r = s_r phi_r, Q_r = -F_r/2, V_c = W s_c phi_c, U_c = (sigma + gamma) V_c - W F_c / 2
// the SIPG + upwind face matrix (full_assembly.py:174-212) is B[r,c] += P_r U_c + Q_r V_c.
template <int P, int RS, int PART, bool DIFF, bool ADV, bool DO_DIAG, bool DO_OFF, bool STAGED>
__device__ __forceinline__ void face_blocks(const AsmArgs& A, const FaceCtx& c, const ElemBox& bxe, const ElemBox& bxo,
                                            const double* __restrict__ s_coef, const double* __restrict__ s_box, int lt,
                                            double* __restrict__ dacc, double* __restrict__ oacc) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, CI = DM::CI, I0 = PART * CI;
    const size_t F = (size_t)A.g.n_iface;
#pragma unroll 1
    for (int q = 0; q < DM::QE; ++q) {
        const FaceCoef cur = STAGED ? staged_face_coef<DM::IT, DIFF, ADV>(s_coef, lt, q)
                                    : load_face_coef<DIFF, ADV>(A, (size_t)q * F + c.f);
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
        eval_basis<P, DIFF>(x, y, RS > 1 ? staged_box<DM::IT, true>(s_box, lt, 0) : bxe, A.lc, phi, gx, gy);
        double Pr[D], Qr[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            Pr[j] = c.sgn * phi[j];
            if (DIFF) Qr[j] = -0.5 * (atn0 * gx[j] + atn1 * gy[j]);
        }
        const double Ws = W * c.sgn;
        if (DO_DIAG) {
#pragma unroll
            for (int ii = 0; ii < CI; ++ii) {
                const int i = I0 + ii;
                const double Vc = Ws * phi[i];
                double Uc = sg * Vc;
                if (DIFF) Uc = fma(hw0, gx[i], fma(hw1, gy[i], Uc));
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double v = dacc[j * CI + ii];
                    v = fma(Pr[j], Uc, v);
                    if (DIFF) v = fma(Qr[j], Vc, v);
                    dacc[j * CI + ii] = v;
                }
            }
        }
        if (DO_OFF) {
            eval_basis<P, DIFF>(x, y, RS > 1 ? staged_box<DM::IT, true>(s_box, lt, 1) : bxo, A.lc, phi, gx, gy);
#pragma unroll
            for (int ii = 0; ii < CI; ++ii) {
                const int i = I0 + ii;
                const double Vc = -Ws * phi[i];
                double Uc = sg * Vc;
                if (DIFF) Uc = fma(hw0, gx[i], fma(hw1, gy[i], Uc));
#pragma unroll
                for (int j = 0; j < D; ++j) {
                    double v = oacc[j * CI + ii];
                    v = fma(Pr[j], Uc, v);
                    if (DIFF) v = fma(Qr[j], Vc, v);
                    oacc[j * CI + ii] = v;
                }
            }
        }
    }
}

template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void item_pass(const AsmArgs& A, const FaceCtx& c0, const ElemBox& bxe, const ElemBox& bxo,
                                          const double* __restrict__ s_coef,
                                          double* __restrict__ s_box, int lt, int k, int k_end, int slot, int nb,
                                          int64_t row_base, double (&dacc)[Dims<P, RS>::NA], unsigned& zeros) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, CI = DM::CI, I0 = PART * CI, NA = DM::NA, IT = DM::IT;
    constexpr bool FUSE = (NA <= kFuseLimit);   // both blocks in one sweep when the accumulators fit in registers
    double oacc[NA];
#pragma unroll
    for (int v = 0; v < NA; ++v) oacc[v] = 0.0;
    // the group leader's own face comes from the coefficient stage ...
    if (FUSE) face_blocks<P, RS, PART, DIFF, ADV, true, true, true>(A, c0, bxe, bxo, s_coef, s_box, lt, dacc, oacc);
    else face_blocks<P, RS, PART, DIFF, ADV, false, true, true>(A, c0, bxe, bxo, s_coef, s_box, lt, dacc, oacc);
    // ... any following items flagged "same neighbour" (merged blocks; non-Voronoi meshes) straight from global memory
    if (A.has_dups) {
        for (int kk = k + 1; kk < k_end && A.adj_item[kk] < 0; ++kk) {
            const FaceCtx c = face_setup<DIFF>(A, A.adj_item[kk] & 0x7fffffff);
            if (FUSE) face_blocks<P, RS, PART, DIFF, ADV, true, true, false>(A, c, bxe, bxo, s_coef, s_box, lt, dacc, oacc);
            else face_blocks<P, RS, PART, DIFF, ADV, false, true, false>(A, c, bxe, bxo, s_coef, s_box, lt, dacc, oacc);
        }
    }
    // off-diagonal block: final after this group -> straight to its CSR slot (CI consecutive columns of every row)
    const int rowlen = nb * D;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double* dst = A.data + row_base + (int64_t)j * rowlen + (int64_t)slot * D + I0;
        if (RS == 1 && (D & 1) == 0) {
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
    if (!FUSE) {
        // second sweep for the diagonal part, after the off-diagonal registers have been retired
        face_blocks<P, RS, PART, DIFF, ADV, true, false, true>(A, c0, bxe, bxe, s_coef, s_box, lt, dacc, dacc);
        if (A.has_dups) {
            for (int kk = k + 1; kk < k_end && A.adj_item[kk] < 0; ++kk) {
                const FaceCtx c = face_setup<DIFF>(A, A.adj_item[kk] & 0x7fffffff);
                face_blocks<P, RS, PART, DIFF, ADV, true, false, false>(A, c, bxe, bxe, s_coef, s_box, lt, dacc, dacc);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void volume_pass(const AsmArgs& A, bool active, int t, int lt, int part,
                                            double* __restrict__ s_coef, double* __restrict__ s_box,
                                            double (&acc)[Dims<P, RS>::NTV]) {
    if (active) {
        stage_vol<P, RS, DIFF, ADV>(A, s_coef, lt, part, t);
        if (RS > 1 && part == 0) stage_box<Dims<P, RS>::IT>(A, s_box, lt, 0, A.g.tri_elem[t]);
    }
    __pipeline_commit();
    TriGeom tg;
    if (active) {
        tg = load_tri_geom(A, t);                   // the dependent geometry gathers overlap the async copies
        if (RS == 1) tg.bx = A.ebox[A.g.tri_elem[t]];
    }
    __pipeline_wait_prior(0);
    if (RS > 1) __syncthreads();                    // the parts of an item staged different points
    if (active) triangle_block<P, RS, PART, DIFF, ADV>(A, tg, s_coef, s_box, lt, acc);
}

template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kBD, (RS > 1 && P < 3) ? RB_MINB : 2) dg_assemble_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dims<P, RS>;
    constexpr int D = DM::D, DD = DM::DD, NV = DM::NV, NA = DM::NA, NTV = DM::NTV, IT = DM::IT, CH = DM::CH;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int ea = blockIdx.x * A.te;
    const int ne = min(A.te, A.g.n_rows - ea);
    double* s_acc = smem;                                   // [te][NV]
    double* s_stage = s_acc + (size_t)A.te * NV;            // [CH][kBD+1]  summation stage, aliased with ...
    double* s_coef = s_stage;                               // [IT][CV | CF] the coefficient stage (never live together)
    double* s_box = s_coef + DM::COEF;                      // [8][IT] double2 planes: the items' element boxes
    int* s_toff = reinterpret_cast<int*>(s_stage + (size_t)DM::UN);   // [te+1]
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

    const int part = tid / IT;          // warp-uniform column part
    const int lt = tid - part * IT;     // item within the pass

    // ---- phase A: volume terms ---------------------------------------------------------------------------------
    {
        const int ta = s_toff[0], tb = s_toff[ne];
        for (int base = ta; base < tb; base += IT) {
            const int t = base + lt;
            const bool active = t < tb;
            double acc[NTV];
#pragma unroll
            for (int v = 0; v < NTV; ++v) acc[v] = 0.0;
            if (RS == 1 || part == 0) volume_pass<P, RS, 0, DIFF, ADV>(A, active, t, lt, part, s_coef, s_box, acc);
            else volume_pass<P, RS, RS - 1, DIFF, ADV>(A, active, t, lt, part, s_coef, s_box, acc);
            __syncthreads();                                // coefficient stage -> summation stage
            staged_sum<NTV, NA, CH, IT, RS, D>(acc, active, base, ne, s_toff, s_acc, s_stage);
        }
    }

    // ---- phase B: interior faces -------------------------------------------------------------------------------
    unsigned zeros = 0;
    {
        const int ka = s_aoff[0], kb = s_aoff[ne];
        for (int base = ka; base < kb; base += IT) {
            const int k = base + lt;
            const bool in_range = k < kb;
            const int item = in_range ? A.adj_item[k] : -1;
            const bool active = item >= 0;                  // followers (dup bit) are handled by their group leader
            if (active) stage_face<P, RS, DIFF, ADV>(A, s_coef, lt, part, item >> 1);
            __pipeline_commit();
            double dacc[NA];
#pragma unroll
            for (int v = 0; v < NA; ++v) dacc[v] = 0.0;
            FaceCtx c0;
            ElemBox bxe, bxo;
            int seg = 0, slot = 0, nb = 1;
            int64_t row_base = 0;
            if (active) {
                // element of this item: last seg with s_aoff[seg] <= k
                int lo = 0, hi = ne;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_aoff[mid] <= k) lo = mid; else hi = mid;
                }
                seg = lo;
                const int e = ea + seg;
                int gi = k - s_aoff[seg];
                if (A.has_dups)
                    for (int a = s_aoff[seg]; a < k; ++a) gi -= (A.adj_item[a] < 0);
                c0 = face_setup<DIFF>(A, item);
                if (RS == 1) {
                    bxe = A.ebox[e];
                    bxo = A.ebox[c0.o];
                } else if (part == 0) {
                    stage_box<IT>(A, s_box, lt, 0, e);
                    stage_box<IT>(A, s_box, lt, 1, c0.o);
                }
                nb = s_goff[seg + 1] - s_goff[seg] + 1;
                // block columns are ordered by GLOBAL element id (partitions number ghosts after the owned elements)
                const int og = A.g.elem_gid ? A.g.elem_gid[c0.o] : c0.o, eg = A.g.elem_gid ? A.g.elem_gid[e] : e;
                slot = gi + (og > eg ? 1 : 0);
                if (og < eg && part == 0) atomicAdd(&s_dslot[seg], 1);
                row_base = (int64_t)DD * (s_goff[seg] + e);
            }
            __pipeline_commit();
            __pipeline_wait_prior(0);
            if (RS > 1) __syncthreads();
            if (active) {
                if (RS == 1 || part == 0)
                    item_pass<P, RS, 0, DIFF, ADV>(A, c0, bxe, bxo, s_coef, s_box, lt, k, s_aoff[seg + 1], slot, nb, row_base, dacc, zeros);
                else
                    item_pass<P, RS, RS - 1, DIFF, ADV>(A, c0, bxe, bxo, s_coef, s_box, lt, k, s_aoff[seg + 1], slot, nb, row_base, dacc, zeros);
            }
            __syncthreads();                                // coefficient stage -> summation stage
            // merged followers stay inactive: their stage column must read as zero -> stage zeros for them
            staged_sum<NA, NA, CH, IT, RS, D>(dacc, in_range, base, ne, s_aoff, s_acc, s_stage);
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
    const size_t smem = asm_smem_bytes(DM::NV, DM::UN, A.te);
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
    const size_t nf = (size_t)(n_iface > 0 ? n_iface : 1);
    return sizeof(ElemBox) * (size_t)(n_elem > 0 ? n_elem : 1) + sizeof(FaceRec) * nf + sizeof(double) * nf +
           sizeof(int32_t) * (size_t)(n_elem > 0 ? n_elem : 1) + 256;
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
    FaceRec* frec = reinterpret_cast<FaceRec*>(w + sizeof(ElemBox) * (size_t)(g->n_elem > 0 ? g->n_elem : 1));
    double* fsigma = reinterpret_cast<double*>(frec + (size_t)(g->n_iface > 0 ? g->n_iface : 1));
    A.ebox = ebox;
    A.fsigma = fsigma;
    A.frec = frec;
    A.perm = reinterpret_cast<int32_t*>(fsigma + (size_t)(g->n_iface > 0 ? g->n_iface : 1));
    A.adj_nb = s->adj_nb;
    {
        const int64_t work = g->n_elem > g->n_iface ? g->n_elem : g->n_iface;
        const int blocks = grid_for(work, 256, kNumSMs * 8);
        dg_prepare_kernel<<<blocks, 256, 0, st>>>(*g, (diff && g->n_iface > 0) ? c->if_a_mid : nullptr, A.sigma_pp, A.p2,
                                                 ebox, fsigma, frec);
        RB_LAUNCH_CHECK("dg_prepare_kernel");
    }
    // threads per item: whole blocks per thread for p <= 2 (measured faster on B200 than column halves at 12 warps per
    // SM), column halves for p = 3; REYNA_B200_RS=2 selects halves at p = 2 (A/B aid)
    int rs = p >= 3 ? 2 : 1;
    if (const char* env = getenv("REYNA_B200_RS")) { if (atoi(env) == 2 && p == 2) rs = 2; }
    const double items_per_elem = (double)s->n_items / g->n_rows;
    const double tris_per_elem = (double)g->n_tri / g->n_rows;
    const int d = (p + 1) * (p + 2) / 2;
    // relative cost of one item of each phase (FMA counts of the two contractions)
    const double w_vol = tris_per_elem * qv * (d * d * (diff ? 3.0 : 1.0) + 12.0 * d);
    const double w_face = items_per_elem * qe * (2.0 * d * d * (diff ? 2.0 : 1.0) + 30.0 * d);
    A.te = choose_te(tris_per_elem, items_per_elem, rs, w_vol, w_face);
    // default: the element-streaming kernel (assemble_elem.cu) for p <= 2, the CTA-cooperative kernel of this file for p = 3
    // (measured on B200); REYNA_B200_KERNEL=cta|elem forces one of them (A/B aid)
    const char* kenv = getenv("REYNA_B200_KERNEL");
    const bool use_elem = kenv ? strcmp(kenv, "elem") == 0 : p <= 2;
    if (use_elem) return launch_assemble_elem(A, p, diff, adv, st);
    switch (p) {
        case 1: return dispatch_coefs<1, 1>(A, diff, adv, st);
        case 2: return rs == 2 ? dispatch_coefs<2, 2>(A, diff, adv, st) : dispatch_coefs<2, 1>(A, diff, adv, st);
        default: return dispatch_coefs<3, 2>(A, diff, adv, st);
    }
}
