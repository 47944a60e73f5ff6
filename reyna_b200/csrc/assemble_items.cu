// Item-per-lane DG assembly for B200: two kernels, both one warp per CTA, both fed and drained by the TMA / async-copy
// engines instead of per-lane global loads and stores.
//
// Replaces DGFEM._stiffness_matrix + localstiff (reyna/DGFEM/two_dimensional/main.py:275-327, full_assembly.py:9-77),
// DGFEM._interior_stiffness_matrix + int_localstiff (main.py:329-374, full_assembly.py:80-214) and the scipy COO->CSR
// summation (main.py:319-320, 371-372, 211) for all elements at once.
//
//   dg_volume_kernel   one lane per SUB-TRIANGLE.  A CTA is one "volume round": <= 32 consecutive sub-triangles of whole
//                      elements (the host packs rounds greedily, rb_geom.vround_tri0).  The coefficient samples of a round
//                      are ONE contiguous range of every sample array (layout: [round][q][lane]) and the vertex
//                      coordinates one contiguous range of rb_geom.tri_xy, so a single elected lane fetches everything
//                      with five cp.async.bulk (1-D TMA) copies completing on one mbarrier -- no per-lane LDGSTS.  Every
//                      lane contracts its triangle into a private d x d block (+ load) in registers, the blocks of an
//                      element's triangles are summed IN TRIANGLE ORDER through shared memory (the order of main.py:315)
//                      and the per-element sums go to a scratch array with fully coalesced 16-byte stores.
//   dg_face_kernel     one lane per (element, interior face) item.  A CTA walks a window of consecutive elements in
//                      rounds of whole elements with <= 32 items.  Face record, neighbour scalings and face samples are
//                      staged ONCE per item by cp.async (double buffered, one round ahead), both blocks of the item are
//                      formed from that one staging: the diagonal contributions are summed per element in item order
//                      (after the volume sums), the off-diagonal block goes to its final CSR slot.  All block rows of a
//                      round are CONTIGUOUS in the CSR value array, so they are collected in shared memory and written
//                      by ONE cp.async.bulk.global.shared::cta store -- no per-lane scattered STG.
// Bitwise reproducible: no floating-point atomics, every sum has a fixed order that does not depend on the partition.
#include "assemble_common.cuh"
#include <cuda_pipeline.h>

namespace rb {
namespace items {

constexpr int kWarp = 32;
constexpr int kMaxSeg = 8;      // volume segments staged per face round
constexpr int kWindow = 64;     // consecutive elements per face CTA at most (AsmArgs.window: 16..64, chosen by the launcher)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kVolWarps = kNumSMs * 11;   // persistent warps of the volume kernel (11 one-warp CTAs fit an SM)

// ---- PTX: mbarrier + bulk async copies ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// bounded spin: a copy that never lands (a bug) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (unsigned spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (ok) return;
        if (spin > (1u << 20)) __trap();
    }
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ int gid_of(const AsmArgs& A, int e) { return A.g.elem_gid ? A.g.elem_gid[e] : e; }
__device__ __forceinline__ int seg0_of(const AsmArgs& A, int e) { return A.g.elem_seg0 ? A.g.elem_seg0[e] : e; }

template <int P, int RS>
struct Dm {
    static constexpr int D = (P + 1) * (P + 2) / 2;
    static constexpr int DD = D * D;
    static constexpr int NV = DD + D;          // values per element sum: block + load
    static constexpr int QV = (P + 1) * (P + 1);
    static constexpr int QE = P + 1;
    static constexpr int RJ = D / RS;          // block rows per pass
    static constexpr int NB = RJ * D;          // block values per pass
    static_assert(D % RS == 0, "row split must divide the local dimension");
    static_assert(RS == 1 || (NB % 2 == 0 && D % 2 == 0), "row parts must be pair aligned");
    // volume pass `part`: NB block values (+ all D load values in the last part)
    static constexpr int NVAL_LAST = NB + D;
    static constexpr int NPAIR_MAX = (NVAL_LAST + 1) / 2;
};

constexpr int kPartStride = 33;   // double2 per pair row of the partial buffer (33: conflict-free column walks)


// destination (in doubles, inside an element's NV-value record) of value pair kk of pass PART: block rows of the part,
// then -- in the last part -- the load entries
template <int P, int RS, int PART>
__device__ __forceinline__ int pair_dst(int kk) {
    using DM = Dm<P, RS>;
    if (RS == 1) return 2 * kk;
    if (2 * kk < DM::NB) return PART * DM::NB + 2 * kk;
    return DM::DD + (2 * kk - DM::NB);
}

// =======================================================================================================================
// volume kernel
// =======================================================================================================================
// Persistent warps: CTA b (one warp) takes the volume rounds b, b + gridDim, ...  The samples of a round live in TWO halves of
// the shared-memory buffer (quadrature points [0, Q0) and [Q0, Qv)), each with its own mbarrier: as soon as the warp has
// consumed a half it refills it with the NEXT round's data, so that in steady state a warp never waits for memory and no
// second buffer is needed -- the copy of one half travels while the other half is being contracted.
template <int P, int RS, bool DIFF, bool ADV>
struct VolSmem {
    using DM = Dm<P, RS>;
    static constexpr int QV = DM::QV, Q0 = (QV + 1) / 2, Q1 = QV - Q0;
    static constexpr int a_bytes(int q) { return DIFF ? q * 32 * 32 : 0; }
    static constexpr int b_bytes(int q) { return ADV ? q * 32 * 16 : 0; }
    static constexpr int c_bytes(int q) { return q * 32 * 8 + 16; }          // 8-byte samples: 16 bytes of alignment slack
    // half 0 | half 1, each: a, b, c, f
    static constexpr int OA0 = 0, OB0 = OA0 + a_bytes(Q0), OC0 = OB0 + b_bytes(Q0), OF0 = OC0 + c_bytes(Q0), END0 = OF0 + c_bytes(Q0);
    static constexpr int OA1 = END0, OB1 = OA1 + a_bytes(Q1), OC1 = OB1 + b_bytes(Q1), OF1 = OC1 + c_bytes(Q1), END1 = OF1 + c_bytes(Q1);
    static constexpr int SAMPLES = END1, H1B = END1 - END0;
    static constexpr int PARTB = DM::NPAIR_MAX * kPartStride * 16;
    // one part (RS == 1): the partial blocks are reduced through the consumed half 1, in NPASS passes of PP value pairs;
    // several parts: the samples have to survive, the partials get a buffer of their own -- a small one, walked in five
    // passes, so that six one-warp CTAs fit an SM at p = 3 (a buffer for all pairs at once: 49 KB per CTA, four CTAs)
    static constexpr int NPASS = RS == 1 ? (PARTB + H1B - 1) / H1B : 5;
    static constexpr int PP = (DM::NPAIR_MAX + NPASS - 1) / NPASS;
    static constexpr int PBUF = PP * kPartStride * 16;
    static constexpr int OPART = RS == 1 ? END0 : SAMPLES;
    static constexpr int OMETA = (RS == 1 ? SAMPLES : SAMPLES + PBUF);        // int s_start[36], s_seg[36]
    static constexpr int OBAR = OMETA + 4 * 72;
    static constexpr int TOTAL = OBAR + 16;
    static_assert(PP * kPartStride * 16 <= (RS == 1 ? H1B : PBUF), "partial buffer");
    static_assert(END0 % 16 == 0 && OMETA % 16 == 0, "alignment");
};

// Quadrature points [QL, QH) of the lane's triangle from one half of the sample buffer: rows [J0, J0+RJ) of the local block
// (+ the load in the last part).  O*: byte offsets of the half's arrays, hc / hf: where the 8-byte samples start in theirs.
template <int P, int RS, int PART, bool DIFF, bool ADV, int QL, int QH, int OA, int OB, int OC, int OF>
__device__ __forceinline__ void volume_points(const AsmArgs& A, const unsigned char* __restrict__ sm, int lane, int n, int hc, int hf,
                                              const ElemBox& bx, double2 V0, double2 V1, double2 V2, double hdet, bool has_c,
                                              bool has_f, double* __restrict__ acc) {
    using DM = Dm<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, NB = DM::NB, J0 = PART * RJ;
    constexpr bool LAST = PART == RS - 1;
#pragma unroll 1
    for (int q = QL; q < QH; ++q) {
        const int idx = (q - QL) * n + lane;
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
        const double r0 = A.rvx[q], r1 = A.rvy[q];
        const double l0 = 1.0 - r0 - r1;
        const double x = l0 * V0.x + r0 * V1.x + r1 * V2.x;      // assembly_aux.py:7-29
        const double y = l0 * V0.y + r0 * V1.y + r1 * V2.y;
        const double W = hdet * A.wv[q];
        if (DIFF) {
            const double2 ra = *reinterpret_cast<const double2*>(sm + OA + (size_t)idx * 32);
            const double2 rb = *reinterpret_cast<const double2*>(sm + OA + (size_t)idx * 32 + 16);
            a00 = W * ra.x; a01 = W * ra.y; a10 = W * rb.x; a11 = W * rb.y;
        }
        if (ADV) {
            const double2 bb = *reinterpret_cast<const double2*>(sm + OB + (size_t)idx * 16);
            b0 = W * bb.x; b1 = W * bb.y;
        }
        const double cq = has_c ? W * *reinterpret_cast<const double*>(sm + OC + hc + (size_t)idx * 8) : 0.0;
        const double wf = (LAST && has_f) ? W * *reinterpret_cast<const double*>(sm + OF + hf + (size_t)idx * 8) : 0.0;
        double phi[D], gx[D], gy[D];
        eval_basis<P, true>(x, y, bx, A.lc, phi, gx, gy);
        // column (trial i) quantities u = W * grad(phi_i)^T a, s = W * (b.grad(phi_i) + c phi_i)   (full_assembly.py:57-77)
#pragma unroll
        for (int i = 0; i < D; ++i) {
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
            for (int jj = 0; jj < RJ; ++jj) {
                const int j = J0 + jj;
                double v = acc[jj * D + i];
                if (DIFF && (nzx || nzy)) {
                    if (gx_nz<P>(j)) v = fma(u0, gx[j], v);
                    if (gy_nz<P>(j)) v = fma(u1, gy[j], v);
                }
                v = fma(si, phi[j], v);
                acc[jj * D + i] = v;
            }
        }
        if (LAST && has_f) {
#pragma unroll
            for (int j = 0; j < D; ++j) acc[NB + j] = fma(wf, phi[j], acc[NB + j]);
        }
    }
}

// Ordered per-segment sums of the lanes' partial blocks of one part -> vdiag, through the partial buffer in NPASS passes.
template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void volume_reduce(const AsmArgs& A, unsigned char* __restrict__ sm, int lane, bool active, int nseg,
                                              const double* __restrict__ acc) {
    using DM = Dm<P, RS>;
    using S = VolSmem<P, RS, DIFF, ADV>;
    constexpr bool LAST = PART == RS - 1;
    constexpr int NVAL = DM::NB + (LAST ? DM::D : 0);
    constexpr int NP = (NVAL + 1) / 2;
    double2* part = reinterpret_cast<double2*>(sm + S::OPART);
    const int* s_start = reinterpret_cast<const int*>(sm + S::OMETA);
    const int* s_seg = s_start + 36;
#pragma unroll
    for (int pass = 0; pass < S::NPASS; ++pass) {
        constexpr int PPc = S::PP;
        const int k0 = pass * PPc;
        const int np = NP - k0 < PPc ? NP - k0 : PPc;      // pairs of this pass
        if (np <= 0) break;
        if (pass > 0) __syncwarp();                        // the previous pass has been read
        if (active) {
#pragma unroll
            for (int kk = 0; kk < PPc; ++kk)
                if (k0 + kk < NP) part[kk * kPartStride + lane] = make_double2(acc[2 * (k0 + kk)], acc[2 * (k0 + kk) + 1]);
        }
        __syncwarp();
        const int total = nseg * np;
        for (int w = lane; w < total; w += kWarp) {
            const int s = w / np, kk = w - s * np;
            const int l0 = s_start[s], l1 = s_start[s + 1];
            double2 v = part[kk * kPartStride + l0];
            for (int l = l0 + 1; l < l1; l += 4) {      // triangles in order (main.py:315); four loads in flight (x + 0.0 = x)
                double2 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = l + u < l1 ? part[kk * kPartStride + l + u] : make_double2(0.0, 0.0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v.x += t[u].x;
                    v.y += t[u].y;
                }
            }
            double* dst = A.vdiag + (size_t)s_seg[s] * DM::NV + pair_dst<P, RS, PART>(k0 + kk);
            if (2 * (k0 + kk) + 1 < NVAL) *reinterpret_cast<double2*>(dst) = v;
            else *dst = v.x;
        }
    }
}

// one lane: arm the half's mbarrier and start the bulk copies of quadrature points [QL, QH) of the round starting at triangle t0
template <int P, int RS, bool DIFF, bool ADV, int QL, int QH, int OA, int OB, int OC, int OF>
__device__ __forceinline__ void issue_half(const AsmArgs& A, unsigned char* __restrict__ sm, uint64_t* bar, int t0, int n) {
    constexpr int QV = Dm<P, RS>::QV;
    const bool has_c = A.c.vol_c != nullptr, has_f = A.c.vol_f != nullptr;
    const size_t first = (size_t)QV * t0 + (size_t)QL * n;
    const uint32_t cnt = (uint32_t)((QH - QL) * n);
    const uintptr_t pc = reinterpret_cast<uintptr_t>(A.c.vol_c + first), pf = reinterpret_cast<uintptr_t>(A.c.vol_f + first);
    const uint32_t hc = has_c ? (uint32_t)(pc & 15) : 0, hf = has_f ? (uint32_t)(pf & 15) : 0;
    const uint32_t ba = DIFF ? cnt * 32 : 0, bb = ADV ? cnt * 16 : 0;
    const uint32_t bc = has_c ? (hc + cnt * 8 + 15) & ~15u : 0, bf = has_f ? (hf + cnt * 8 + 15) & ~15u : 0;
    mbar_expect_tx(bar, ba + bb + bc + bf);
    if (DIFF) bulk_g2s(sm + OA, A.c.vol_a + 4 * first, ba, bar);
    if (ADV) bulk_g2s(sm + OB, A.c.vol_b + 2 * first, bb, bar);
    if (has_c) bulk_g2s(sm + OC, reinterpret_cast<const void*>(pc - hc), bc, bar);
    if (has_f) bulk_g2s(sm + OF, reinterpret_cast<const void*>(pf - hf), bf, bar);
}

template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kWarp, P <= 2 ? 12 : 1) dg_volume_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dm<P, RS>;
    using S = VolSmem<P, RS, DIFF, ADV>;
    constexpr int QV = DM::QV, Q0 = S::Q0;
    constexpr bool LASTP = true;
    (void)LASTP;
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x;
    // rounds are handed out dynamically (A.vol_next, zeroed by the launcher): the warps of an SM do not run at the same speed
    int r = A.vround_lo + blockIdx.x;
    if (r >= A.vround_hi) return;
    const bool has_c = A.c.vol_c != nullptr, has_f = A.c.vol_f != nullptr;
    uint64_t* bar0 = reinterpret_cast<uint64_t*>(sm + S::OBAR);
    uint64_t* bar1 = bar0 + 1;
    int t0 = A.g.vround_tri0[r], t1 = A.g.vround_tri0[r + 1];
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        issue_half<P, RS, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, bar0, t0, t1 - t0);
        issue_half<P, RS, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, bar1, t0, t1 - t0);
    }
    __syncwarp();
    for (unsigned k = 0; r < A.vround_hi; ++k) {
        const int n = t1 - t0;
        // the next round of this warp (its triangle range is needed when the halves are refilled)
        int rn = 0;
        if (lane == 0) rn = A.vround_lo + (int)gridDim.x + atomicAdd(A.vol_next, 1);
        rn = __shfl_sync(kFull, rn, 0);
        const bool has_next = rn < A.vround_hi;
        int tn0 = 0, tn1 = 0;
        if (has_next) {
            tn0 = A.g.vround_tri0[rn];
            tn1 = A.g.vround_tri0[rn + 1];
        }
        const bool active = lane < n;
        const int t = t0 + (active ? lane : 0);
        const int e = A.g.tri_elem[t];
        const int seg = A.g.tri_seg ? A.g.tri_seg[t] : e;
        const ElemBox bx = A.ebox[e];
        const double2* xy = reinterpret_cast<const double2*>(A.g.tri_xy) + 3 * (size_t)t;
        const double2 V0 = xy[0], V1 = xy[1], V2 = xy[2];
        // segments (= elements, or pieces of an element with more than 32 triangles) of the round
        const int prev = __shfl_up_sync(kFull, seg, 1);
        const bool head = active && (lane == 0 || seg != prev);
        const unsigned hm = __ballot_sync(kFull, head);
        const int nseg = __popc(hm);
        {
            int* s_start = reinterpret_cast<int*>(sm + S::OMETA);
            int* s_seg = s_start + 36;
            if (head) {
                const int rank = __popc(hm & lanemask_lt());
                s_start[rank] = lane;
                s_seg[rank] = seg;
            }
            if (lane == 0) s_start[nseg] = n;
        }
        // where the 8-byte samples start inside their (16-byte aligned) copies
        const size_t f0 = (size_t)QV * t0, f1 = f0 + (size_t)Q0 * n;
        const int hc0 = has_c ? (int)(reinterpret_cast<uintptr_t>(A.c.vol_c + f0) & 15) : 0;
        const int hc1 = has_c ? (int)(reinterpret_cast<uintptr_t>(A.c.vol_c + f1) & 15) : 0;
        const int hf0 = has_f ? (int)(reinterpret_cast<uintptr_t>(A.c.vol_f + f0) & 15) : 0;
        const int hf1 = has_f ? (int)(reinterpret_cast<uintptr_t>(A.c.vol_f + f1) & 15) : 0;
        // |det(0.5*[V1-V0; V2-V0])| (full_assembly.py:40-41); the extra 0.5 is the `0.5 * z` of :75-77
        const double e1x = 0.5 * (V1.x - V0.x), e1y = 0.5 * (V1.y - V0.y);
        const double e2x = 0.5 * (V2.x - V0.x), e2y = 0.5 * (V2.y - V0.y);
        const double hdet = 0.5 * fabs(e1x * e2y - e1y * e2x);
        const uint32_t parity = k & 1u;

        if (RS == 1) {
            double acc[2 * DM::NPAIR_MAX];
#pragma unroll
            for (int v = 0; v < 2 * DM::NPAIR_MAX; ++v) acc[v] = 0.0;
            mbar_wait(bar0, parity);
            if (active)
                volume_points<P, RS, 0, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, lane, n, hc0, hf0, bx, V0, V1, V2, hdet,
                                                                                       has_c, has_f, acc);
            __syncwarp();                           // half 0 is consumed: it takes the next round's points [0, Q0)
            if (has_next && lane == 0)
                issue_half<P, RS, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, bar0, tn0, tn1 - tn0);
            mbar_wait(bar1, parity);
            if (active)
                volume_points<P, RS, 0, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, lane, n, hc1, hf1, bx, V0, V1, V2, hdet,
                                                                                        has_c, has_f, acc);
            __syncwarp();                           // half 1 is consumed: the partial blocks are reduced through it
            volume_reduce<P, RS, 0, DIFF, ADV>(A, sm, lane, active, nseg, acc);
            fence_async_smem();                     // generic writes of the partial buffer before the async refill
            __syncwarp();
            if (has_next && lane == 0)
                issue_half<P, RS, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, bar1, tn0, tn1 - tn0);
        } else {
            // several row parts: every part sweeps all points, the buffer is refilled when the round is done
            mbar_wait(bar0, parity);
            mbar_wait(bar1, parity);
            {
                double acc[2 * ((DM::NB + 1) / 2)];
#pragma unroll
                for (int v = 0; v < 2 * ((DM::NB + 1) / 2); ++v) acc[v] = 0.0;
                if (active) {
                    volume_points<P, RS, 0, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, lane, n, hc0, hf0, bx, V0, V1, V2,
                                                                                           hdet, has_c, has_f, acc);
                    volume_points<P, RS, 0, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, lane, n, hc1, hf1, bx, V0, V1, V2,
                                                                                            hdet, has_c, has_f, acc);
                }
                __syncwarp();
                volume_reduce<P, RS, 0, DIFF, ADV>(A, sm, lane, active, nseg, acc);
            }
            {
                double acc[2 * DM::NPAIR_MAX];
#pragma unroll
                for (int v = 0; v < 2 * DM::NPAIR_MAX; ++v) acc[v] = 0.0;
                if (active) {
                    volume_points<P, RS, RS - 1, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, lane, n, hc0, hf0, bx, V0, V1, V2,
                                                                                                hdet, has_c, has_f, acc);
                    volume_points<P, RS, RS - 1, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, lane, n, hc1, hf1, bx, V0, V1,
                                                                                                 V2, hdet, has_c, has_f, acc);
                }
                __syncwarp();
                volume_reduce<P, RS, RS - 1, DIFF, ADV>(A, sm, lane, active, nseg, acc);
            }
            __syncwarp();
            if (has_next && lane == 0) {
                issue_half<P, RS, DIFF, ADV, 0, Q0, S::OA0, S::OB0, S::OC0, S::OF0>(A, sm, bar0, tn0, tn1 - tn0);
                issue_half<P, RS, DIFF, ADV, Q0, QV, S::OA1, S::OB1, S::OC1, S::OF1>(A, sm, bar1, tn0, tn1 - tn0);
            }
        }
        __syncwarp();                               // the round's segment table may be overwritten
        r = rn;
        t0 = tn0;
        t1 = tn1;
    }
}

// =======================================================================================================================
// face kernel
// =======================================================================================================================
// One warp per CTA, one lane per (element, face) item, ONE sweep over the face points of a round:
//   item lanes    evaluate the element's basis once per point: P, Q (rows) and U, V (columns of the diagonal part) go to a
//                 shared-memory record; the neighbour's basis gives the columns of the off-diagonal block, which the lane
//                 accumulates in registers;
//   row owners    lane (i, j) owns row j of the diagonal block of element i of the round: it starts from the volume sums and
//                 adds, item by item in adjacency order, P_j U + Q_j V from the records -- no cross-lane reduction, no
//                 partial blocks, a fixed order that does not depend on the partition.
// Latency is hidden by occupancy (single staging buffer, like the volume kernel); only the item indices are fetched one
// round ahead so that the dependent chain adjacency -> item -> face data never stalls a round.
constexpr int kTab = 72;           // ints per window table (kWindow + 1 entries used)

template <int P, int RS, bool DIFF, bool ADV>
struct FaceSmem {
    using DM = Dm<P, RS>;
    static constexpr int NT = P <= 2 ? 1 : 2;                                   // row-owner tasks per lane
    static constexpr int MAXE = (kWarp * NT) / DM::D < 8 ? (kWarp * NT) / DM::D : 8;   // elements per round
    static constexpr int NPL = 8 + 3 * DM::QE;                  // 16-byte planes per item: record 4, neighbour box 4, a 2/point, b 1/point
    static constexpr int PLANES = NPL * 512 + 256;               // + one 8-byte plane: penalties of the geometry-table path
    static constexpr int OOWN = PLANES;                          // own boxes of the round's elements [MAXE][64 bytes]
    static constexpr int OVD = OOWN + MAXE * 64;                 // volume sums of the round's segments [kMaxSeg][NV]
    static constexpr int STAGE = OVD + kMaxSeg * DM::NV * 8;     // the staging buffer
    static constexpr int OUTB = (32 + MAXE) * DM::DD * 8 + 16;
    static constexpr int REC = 4 * DM::D * 8 + 16;               // P, Q, U, V of one item at one point (+16: conflict-free lane stride)
    static constexpr int PQUVB = 32 * REC;
    static constexpr int OOUT = STAGE;                           // block rows of the round (+ 16 bytes of alignment shift)
    static constexpr int OPQ = OOUT;                             // the records live in the row buffer until the rows are written
    static constexpr int OWIN = OOUT + (OUTB > PQUVB ? OUTB : PQUVB);
    static constexpr int TOTAL = OWIN + 6 * kTab * 4;            // window tables: adj_off, grp_off, first segment, gid, round start / size
    static_assert(STAGE % 16 == 0 && OUTB % 16 == 0 && REC % 16 == 0, "alignment");
};

struct FaceGeo {
    double midx, midy, tanx, tany, nx, ny, De, sigma;
};

// a round: whole elements [e0, e0+ne) of the window (window-relative), their items [k0,k1), neighbour groups [g0,g1) and
// volume segments [sg0, sg0+nsg); lane i <= ne also holds the table entries of element e0+i
struct Round {
    int e0, ne, k0, k1, g0, g1, sg0, nsg;
    int off, goff, sg;
};

__device__ __forceinline__ Round load_round(const int* __restrict__ tab, int r, int lane) {
    const int *w_off = tab, *w_goff = tab + kTab, *w_sg = tab + 2 * kTab, *w_round = tab + 4 * kTab, *w_rne = tab + 5 * kTab;
    Round x;
    x.e0 = w_round[r];
    x.ne = w_rne[r];
    const int i = x.e0 + (lane <= x.ne ? lane : 0);
    x.off = w_off[i];
    x.goff = w_goff[i];
    x.sg = w_sg[i];
    x.k0 = w_off[x.e0]; x.k1 = w_off[x.e0 + x.ne];
    x.g0 = w_goff[x.e0]; x.g1 = w_goff[x.e0 + x.ne];
    x.sg0 = w_sg[x.e0]; x.nsg = w_sg[x.e0 + x.ne] - x.sg0;
    return x;
}

struct Items {
    int item, nbr;
};

__device__ __forceinline__ Items load_items(const AsmArgs& A, const int* __restrict__ tab, int r, int lane) {
    const int *w_off = tab, *w_round = tab + 4 * kTab, *w_rne = tab + 5 * kTab;
    const int k0 = w_off[w_round[r]], k1 = w_off[w_round[r] + w_rne[r]];
    Items it;
    it.item = 0; it.nbr = 0;
    if (lane < k1 - k0) {
        it.item = A.adj_item[k0 + lane];
        it.nbr = A.adj_nb[k0 + lane];
    }
    return it;
}

// cp.async staging of one round (everything an item needs, copied once).  (Fetching the per-item data into registers with
// vector loads and writing the planes with 128-bit shared stores instead costs fewer shared-memory wavefronts -- a per-lane
// cp.async gather lands lane by lane -- but was slower: profiles/r02_assemble_experiments.md.)
template <int P, int RS, bool DIFF, bool ADV>
__device__ __forceinline__ void stage_round(const AsmArgs& A, unsigned char* __restrict__ sb, const Round& x, const Items& it,
                                            int e_beg, int lane) {
    using DM = Dm<P, RS>;
    using S = FaceSmem<P, RS, DIFF, ADV>;
    constexpr int QE = DM::QE;
    const size_t F = (size_t)A.g.n_iface;
    if (lane < x.k1 - x.k0) {
        const int f = (it.item & 0x7fffffff) >> 1;
        unsigned char* dst = sb + lane * 16;
        const double2* rec = reinterpret_cast<const double2*>(A.frec + f);
        const double2* box = reinterpret_cast<const double2*>(A.ebox + it.nbr);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __pipeline_memcpy_async(dst + k * 512, rec + k, 16);
            __pipeline_memcpy_async(dst + (4 + k) * 512, box + k, 16);
        }
        if (A.fsigma) __pipeline_memcpy_async(sb + S::NPL * 512 + lane * 8, A.fsigma + f, 8);
#pragma unroll
        for (int q = 0; q < QE; ++q) {
            const size_t m = (size_t)q * F + f;
            if (DIFF) {
                __pipeline_memcpy_async(dst + (8 + 2 * q) * 512, reinterpret_cast<const double2*>(A.c.if_a) + 2 * m, 16);
                __pipeline_memcpy_async(dst + (9 + 2 * q) * 512, reinterpret_cast<const double2*>(A.c.if_a) + 2 * m + 1, 16);
            }
            if (ADV) __pipeline_memcpy_async(dst + (8 + 2 * QE + q) * 512, reinterpret_cast<const double2*>(A.c.if_b) + m, 16);
        }
    }
    if (lane < x.ne) {
        const double2* box = reinterpret_cast<const double2*>(A.ebox + e_beg + x.e0 + lane);
#pragma unroll
        for (int k = 0; k < 4; ++k) __pipeline_memcpy_async(sb + S::OOWN + lane * 64 + k * 16, box + k, 16);
    }
    {
        const int pieces = x.nsg * (DM::NV / 2);
        const double2* src = reinterpret_cast<const double2*>(A.vdiag + (size_t)x.sg0 * DM::NV);
        for (int w = lane; w < pieces; w += kWarp) __pipeline_memcpy_async(sb + S::OVD + w * 16, src + w, 16);
    }
}

__device__ __forceinline__ ElemBox lds_box(const unsigned char* p, int stride) {
    const double2 m = *reinterpret_cast<const double2*>(p), ih = *reinterpret_cast<const double2*>(p + stride);
    const double2 s05 = *reinterpret_cast<const double2*>(p + 2 * stride), s15 = *reinterpret_cast<const double2*>(p + 3 * stride);
    ElemBox b;
    b.mx = m.x; b.my = m.y; b.ihx = ih.x; b.ihy = ih.y; b.s05x = s05.x; b.s05y = s05.y; b.s15x = s15.x; b.s15y = s15.y;
    return b;
}

// Exact zeros are rare: one FMNMX per value on the high word (v == +-0 <=> its high word, read as a float, is +-0 and the
// low word is 0; a denormal with a zero high word only sends the caller to the exact count).
template <int N>
__device__ __forceinline__ unsigned count_zeros(const double (&v)[N]) {
    float m = 1.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) m = fminf(m, fabsf(__int_as_float(__double2hiint(v[i]))));
    unsigned z = 0;
    if (m == 0.0f) {
#pragma unroll
        for (int i = 0; i < N; ++i) z += (v[i] == 0.0);
    }
    return z;
}

__device__ __forceinline__ unsigned lowmask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }

// per-lane view of the round
struct LaneCtx {
    bool active;
    int el;          // element of the item inside the round
    int slot;        // block column slot of the off-diagonal block in its row
    int ro, rowlen;  // first double of the element's rows in the row buffer, doubles per scalar row
    int depth;       // 0 = first item of its neighbour group, k = k-th repeat of the same neighbour
    double sgn;
    bool down;
};

// row-owner task: row `row` of the diagonal block of element `el` of the round, fed by the items of lanes [l0, l1)
struct RowTask {
    bool on;
    int el, row, l0, l1, ro, rowlen, dslot;
};

// One sweep over the face points for row part PART of the off-diagonal blocks (and, in part 0, all diagonal rows):
//   P_r = s phi_r, Q_r = -F_r/2, V_c = W s phi_c, U_c = (sigma + gamma) V_c - W F_c/2;  B[r,c] += P_r U_c + Q_r V_c
// (full_assembly.py:174-212; s = +1 on the K1 side, -1 on the K2 side; diagonal part: columns from the element's own basis,
// off-diagonal part: columns from the neighbour's basis with the opposite sign).
template <int P, int RS, int PART, bool DIFF, bool ADV, int NT>
__device__ __forceinline__ void face_sweep(const AsmArgs& A, const unsigned char* __restrict__ sb, unsigned char* __restrict__ pquv,
                                           int lane, const LaneCtx& c, const RowTask (&task)[NT], double* __restrict__ oacc,
                                           double (&racc)[NT][Dm<P, RS>::D]) {
    using DM = Dm<P, RS>;
    using S = FaceSmem<P, RS, DIFF, ADV>;
    constexpr int D = DM::D, RJ = DM::RJ, QE = DM::QE, J0 = PART * RJ;
#pragma unroll 1
    for (int q = 0; q < QE; ++q) {
        if (c.active) {
            // face record and boxes are re-read from the staging buffer at every point: 48 registers less across the sweep
            FaceGeo g;
            {
                const unsigned char* p = sb + lane * 16;
                const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 512);
                const double2 cc = *reinterpret_cast<const double2*>(p + 1024), d = *reinterpret_cast<const double2*>(p + 1536);
                g.midx = a.x; g.midy = a.y; g.tanx = b.x; g.tany = b.y; g.nx = cc.x; g.ny = cc.y; g.De = d.x; g.sigma = d.y;
                if (A.fsigma) g.sigma = *reinterpret_cast<const double*>(sb + S::NPL * 512 + lane * 8);
            }
            const ElemBox bxe = lds_box(sb + S::OOWN + c.el * 64, 16);
            const ElemBox bxo = lds_box(sb + 4 * 512 + lane * 16, 512);
            double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
            if (DIFF) {
                const double2 r0 = *reinterpret_cast<const double2*>(sb + (8 + 2 * q) * 512 + lane * 16);
                const double2 r1 = *reinterpret_cast<const double2*>(sb + (9 + 2 * q) * 512 + lane * 16);
                a00 = r0.x; a01 = r0.y; a10 = r1.x; a11 = r1.y;
            }
            if (ADV) {
                const double2 bb = *reinterpret_cast<const double2*>(sb + (8 + 2 * QE + q) * 512 + lane * 16);
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
                gamma = c.down ? -c.sgn * beta : 0.0;
            }
            const double hw0 = -0.5 * W * atn0, hw1 = -0.5 * W * atn1;     // -W/2 * a^T n
            const double sg = g.sigma + gamma;
            const double Ws = W * c.sgn;
            double phi[D], gx[D], gy[D];
            eval_basis<P, DIFF>(x, y, bxe, A.lc, phi, gx, gy);
            double Pr[D], Qr[D];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                Pr[j] = c.sgn * phi[j];
                Qr[j] = 0.0;
                if (DIFF) {
                    if (gx_nz<P>(j) && gy_nz<P>(j)) Qr[j] = -0.5 * (atn0 * gx[j] + atn1 * gy[j]);
                    else if (gx_nz<P>(j)) Qr[j] = -0.5 * (atn0 * gx[j]);
                    else if (gy_nz<P>(j)) Qr[j] = -0.5 * (atn1 * gy[j]);
                }
            }
            if (PART == 0) {
                // the item's record for the row owners: (P_j, Q_j) pairs, then (U_i, V_i) pairs of the diagonal part
                double* rec = reinterpret_cast<double*>(pquv + lane * S::REC);
                double Ue[D], Ve[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    Ve[i] = Ws * phi[i];
                    Ue[i] = sg * Ve[i];
                    if (DIFF && gy_nz<P>(i)) Ue[i] = fma(hw1, gy[i], Ue[i]);
                    if (DIFF && gx_nz<P>(i)) Ue[i] = fma(hw0, gx[i], Ue[i]);
                }
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    reinterpret_cast<double2*>(rec)[i] = make_double2(Pr[i], Qr[i]);
                    reinterpret_cast<double2*>(rec)[D + i] = make_double2(Ue[i], Ve[i]);
                }
            }
            eval_basis<P, DIFF>(x, y, bxo, A.lc, phi, gx, gy);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const double Vc = -Ws * phi[i];
                double Uc = sg * Vc;
                if (DIFF && gy_nz<P>(i)) Uc = fma(hw1, gy[i], Uc);
                if (DIFF && gx_nz<P>(i)) Uc = fma(hw0, gx[i], Uc);
#pragma unroll
                for (int jj = 0; jj < RJ; ++jj) {
                    const int j = J0 + jj;
                    double v = oacc[jj * D + i];
                    v = fma(Pr[j], Uc, v);
                    if (DIFF && (gx_nz<P>(j) || gy_nz<P>(j))) v = fma(Qr[j], Vc, v);
                    oacc[jj * D + i] = v;
                }
            }
        }
        if (PART == 0) {
            __syncwarp();                           // the records of point q are written
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (task[t].on) {
                    for (int l = task[t].l0; l < task[t].l1; ++l) {      // items in adjacency order
                        const double2* rec = reinterpret_cast<const double2*>(pquv + l * S::REC);
                        const double2 pq = rec[task[t].row];
                        double2 uv[D];
#pragma unroll
                        for (int i = 0; i < D; ++i) uv[i] = rec[D + i];
#pragma unroll
                        for (int i = 0; i < D; ++i) racc[t][i] = fma(pq.y, uv[i].y, fma(pq.x, uv[i].x, racc[t][i]));
                    }
                }
            }
            __syncwarp();                           // the records may be overwritten by the next point
        }
    }
}

template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kWarp) dg_face_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dm<P, RS>;
    using S = FaceSmem<P, RS, DIFF, ADV>;
    constexpr int D = DM::D, DD = DM::DD, NV = DM::NV, RJ = DM::RJ, NB = DM::NB, NT = S::NT, MAXE = S::MAXE;
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x;
    const int e_beg = A.elem_lo + blockIdx.x * A.window;
    const int e_end = min(e_beg + A.window, A.elem_hi);
    const int nw = e_end - e_beg;
    int* tab = reinterpret_cast<int*>(sm + S::OWIN);
    int *w_off = tab, *w_goff = tab + kTab, *w_sg = tab + 2 * kTab, *w_eg = tab + 3 * kTab, *w_round = tab + 4 * kTab;
    int* w_rne = tab + 5 * kTab;
    // ---- window tables + the rounds of the window ---------------------------------------------------------------------------
    for (int i = lane; i <= nw; i += kWarp) {
        w_off[i] = A.adj_off[e_beg + i];
        w_goff[i] = A.grp_off[e_beg + i];
        w_sg[i] = A.g.elem_seg0 ? A.g.elem_seg0[e_beg + i] : e_beg + i;
        if (i < nw) w_eg[i] = gid_of(A, e_beg + i);
    }
    __syncwarp();
    int nr = 0;
    for (int e = 0; e < nw;) {
        const int idx = e + lane;
        const bool in = idx <= nw && lane <= MAXE;
        const int off = in ? w_off[idx] : 0, sg = in ? w_sg[idx] : 0;
        const int k0 = __shfl_sync(kFull, off, 0), sg0 = __shfl_sync(kFull, sg, 0);
        // an element is taken while the items fit one warp and the segments fit the staging buffer
        const bool ok = in && lane >= 1 && (off - k0 <= kWarp) && (sg - sg0 <= kMaxSeg);
        const unsigned m = __ballot_sync(kFull, ok) >> 1;
        const int ne = __ffs(~m) - 1;              // length of the run of set bits (both conditions are monotone)
        if (ne == 0) {                             // an element with more than 32 items: assembled by dg_face_giant_kernel
            ++e;
            continue;
        }
        if (lane == 0) {
            w_round[nr] = e;
            w_rne[nr] = ne;
        }
        e += ne;
        ++nr;
    }
    __syncwarp();
    if (nr == 0) return;
    unsigned zeros = 0;
    unsigned char* sb = sm;
    unsigned char* pquv = sm + S::OPQ;

    // items of round r are fetched two rounds ahead (registers), its face data one round ahead (cp.async, issued as soon as
    // the sweep of round r-1 has consumed the staging buffer)
    Items cur = load_items(A, tab, 0, lane), nxt = cur;
    if (nr > 1) nxt = load_items(A, tab, 1, lane);
    unsigned upw = 0;
    int og = 0;
    {
        const Round x0 = load_round(tab, 0, lane);
        stage_round<P, RS, DIFF, ADV>(A, sb, x0, cur, e_beg, lane);
        __pipeline_commit();
        if (lane < x0.k1 - x0.k0) {
            if (A.c.if_upwind) upw = A.c.if_upwind[(cur.item & 0x7fffffff) >> 1];
            og = A.g.elem_gid ? A.g.elem_gid[cur.nbr] : cur.nbr;
        }
    }
    int t_el[NT], t_row[NT];                        // row-owner tasks of this lane: (element of the round, row), in order
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        t_el[t] = (lane + t * kWarp) / D;
        t_row[t] = (lane + t * kWarp) - t_el[t] * D;
    }
    for (int r = 0; r < nr; ++r) {
        const Round x = load_round(tab, r, lane);
        const int nit = x.k1 - x.k0;
        LaneCtx c;
        c.active = lane < nit;

        // ---- decode the round ----------------------------------------------------------------------------------------------------
        const int elem_ls = x.off - x.k0;          // lane i <= ne: first lane of element i
        c.el = 0;
#pragma unroll
        for (int i = 1; i < MAXE; ++i) {
            const int o = __shfl_sync(kFull, elem_ls, i);
            if (i < x.ne && lane >= o) ++c.el;
        }
        if (!c.active) c.el = 0;
        const int my_ls = __shfl_sync(kFull, elem_ls, c.el);
        const int eg = w_eg[x.e0 + c.el];
        const bool headf = c.active && cur.item >= 0;
        const unsigned hm = __ballot_sync(kFull, headf);
        const unsigned below = (2u << lane) - 1u;
        c.slot = __popc(hm & below & ~lowmask(my_ls)) - 1 + (og > eg ? 1 : 0);
        c.depth = c.active ? lane - (31 - __clz(hm & below)) : 0;
        const int maxdepth = A.has_dups ? __reduce_max_sync(kFull, c.depth) : 0;
        const unsigned lm = __ballot_sync(kFull, headf && og < eg);
        const int ls_next = __shfl_down_sync(kFull, elem_ls, 1);
        const int goff_next = __shfl_down_sync(kFull, x.goff, 1);
        const int elem_dslot = __popc(lm & lowmask(ls_next) & ~lowmask(elem_ls));
        const int elem_ro = DD * ((x.goff - x.g0) + lane);
        const int elem_rowlen = (goff_next - x.goff + 1) * D;
        c.ro = __shfl_sync(kFull, elem_ro, c.el);
        c.rowlen = __shfl_sync(kFull, elem_rowlen, c.el);
        const int side = cur.item & 1;
        c.sgn = side ? -1.0 : 1.0;
        c.down = A.c.if_upwind ? ((upw != 0) == (side == 1)) : false;
        // row-owner tasks of this lane: (element, row) pairs in order
        RowTask task[NT];
        int t_s0[NT], t_s1[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            task[t].on = t_el[t] < x.ne;
            task[t].el = task[t].on ? t_el[t] : 0;
            task[t].row = t_row[t];
            task[t].l0 = __shfl_sync(kFull, elem_ls, task[t].el);
            task[t].l1 = __shfl_sync(kFull, elem_ls, task[t].el + 1);
            task[t].ro = __shfl_sync(kFull, elem_ro, task[t].el);
            task[t].rowlen = __shfl_sync(kFull, elem_rowlen, task[t].el);
            task[t].dslot = __shfl_sync(kFull, elem_dslot, task[t].el);
            t_s0[t] = __shfl_sync(kFull, x.sg, task[t].el) - x.sg0;
            t_s1[t] = __shfl_sync(kFull, x.sg, task[t].el + 1) - x.sg0;
        }
        // first byte of the round's rows in the value array; the row buffer starts at the same 16-byte phase
        double* gdst = A.data + (size_t)DD * ((size_t)x.g0 + e_beg + x.e0);
        const int oshift = (int)(reinterpret_cast<uintptr_t>(gdst) & 15);
        double* outd = reinterpret_cast<double*>(sm + S::OOUT + oshift);
        const int ndbl = DD * ((x.g1 - x.g0) + x.ne);      // doubles of the round's rows

        __pipeline_wait_prior(0);
        if (lane == 0) bulk_wait_read();            // the previous round's bulk store has read the row buffer (= record buffer)
        __syncwarp();

        // diagonal rows start from the volume sums (triangles first, main.py:315), the load is the volume sum itself
        double racc[NT][D];
        double lval[NT];
        const double* vd = reinterpret_cast<const double*>(sb + S::OVD);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int i = 0; i < D; ++i) racc[t][i] = 0.0;
            lval[t] = 0.0;
            if (task[t].on) {
                for (int s = t_s0[t]; s < t_s1[t]; ++s) {
#pragma unroll
                    for (int i = 0; i < D; ++i) racc[t][i] += vd[(size_t)s * NV + task[t].row * D + i];
                    lval[t] += vd[(size_t)s * NV + DD + task[t].row];
                }
            }
        }
        double oacc[NB];
#pragma unroll
        for (int v = 0; v < NB; ++v) oacc[v] = 0.0;
        face_sweep<P, RS, 0, DIFF, ADV, NT>(A, sb, pquv, lane, c, task, oacc, racc);

        // ---- rows of the round: off-diagonal blocks, diagonal rows, load -----------------------------------------------------------
#pragma unroll
        for (int part = 0; part < RS; ++part) {
            if (part > 0) {
#pragma unroll
                for (int v = 0; v < NB; ++v) oacc[v] = 0.0;
                face_sweep<P, RS, RS - 1, DIFF, ADV, NT>(A, sb, pquv, lane, c, task, oacc, racc);
                __syncwarp();
            }
            if (part == RS - 1 && r + 1 < nr) {
                // the staging buffer is consumed: the next round's face data starts travelling while the rows are written
                const Round xn = load_round(tab, r + 1, lane);
                stage_round<P, RS, DIFF, ADV>(A, sb, xn, nxt, e_beg, lane);
                if (lane < xn.k1 - xn.k0) {
                    if (A.c.if_upwind) upw = A.c.if_upwind[(nxt.item & 0x7fffffff) >> 1];
                    og = A.g.elem_gid ? A.g.elem_gid[nxt.nbr] : nxt.nbr;
                }
                cur = nxt;
                if (r + 2 < nr) nxt = load_items(A, tab, r + 2, lane);
            }
            if (part == RS - 1) __pipeline_commit();
            double* dst0 = outd + c.ro + c.slot * D;
            if (c.active && c.depth == 0) {
#pragma unroll
                for (int jj = 0; jj < RJ; ++jj) {
                    double* dst = dst0 + (part * RJ + jj) * c.rowlen;
                    if ((D & 1) == 0) {
#pragma unroll
                        for (int i = 0; i < D; i += 2) reinterpret_cast<double2*>(dst)[i >> 1] = make_double2(oacc[jj * D + i], oacc[jj * D + i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < D; ++i) dst[i] = oacc[jj * D + i];
                    }
                }
                if (maxdepth == 0) zeros += count_zeros<NB>(reinterpret_cast<const double(&)[NB]>(*oacc));
            }
            // faces that repeat the previous neighbour add to its block, in item order
            for (int dd = 1; dd <= maxdepth; ++dd) {
                __syncwarp();
                if (c.active && c.depth == dd) {
#pragma unroll
                    for (int jj = 0; jj < RJ; ++jj) {
                        double* dst = dst0 + (part * RJ + jj) * c.rowlen;
#pragma unroll
                        for (int i = 0; i < D; ++i) dst[i] += oacc[jj * D + i];
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (task[t].on) {
                double* dst = outd + task[t].ro + task[t].row * task[t].rowlen + task[t].dslot * D;
                if ((D & 1) == 0) {
#pragma unroll
                    for (int i = 0; i < D; i += 2) reinterpret_cast<double2*>(dst)[i >> 1] = make_double2(racc[t][i], racc[t][i + 1]);
                } else {
#pragma unroll
                    for (int i = 0; i < D; ++i) dst[i] = racc[t][i];
                }
                if (maxdepth == 0) zeros += count_zeros<D>(racc[t]);
                A.load[(size_t)(e_beg + x.e0 + task[t].el) * D + task[t].row] = lval[t];
            }
        }
        __syncwarp();
        if (maxdepth > 0) {
            // merged blocks: count the exact zeros of the finished rows (the per-lane counts above were skipped)
            for (int w = lane; w < ndbl; w += kWarp) zeros += (outd[w] == 0.0);
        }
        // ---- one bulk store for all rows of the round ------------------------------------------------------------------------------
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            int lo = 0, n = ndbl;
            if (oshift) { gdst[0] = outd[0]; lo = 1; --n; }
            const int core = n & ~1;
            if (core) bulk_s2g(gdst + lo, outd + lo, (uint32_t)core * 8);
            if (n & 1) gdst[lo + core] = outd[lo + core];
        }
    }
    if (lane == 0) bulk_wait_read();
    if (zeros) atomicAdd(A.zero_count, (unsigned long long)zeros);
}

// =======================================================================================================================
// elements with more than 32 (element, face) items -- agglomerated / many-sided polygons.  Rare, so simple: one warp per such
// element, items in chunks of 32 lanes read straight from global memory, two passes per item (diagonal contribution,
// off-diagonal block); the diagonal partials are summed in item order through shared memory into a running sum (volume sums
// first), the off-diagonal blocks go straight to their CSR slots.  dg_face_kernel leaves these elements out.
// =======================================================================================================================
template <int P, int RS, int PART, bool DIFF, bool ADV, bool OFF>
__device__ __forceinline__ void giant_pass(const AsmArgs& A, int f, const FaceGeo& g, double sgn, bool down, const ElemBox& bxe,
                                           const ElemBox& bxo, double* __restrict__ acc) {
    using DM = Dm<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, QE = DM::QE, J0 = PART * RJ;
    const size_t F = (size_t)A.g.n_iface;
#pragma unroll 1
    for (int q = 0; q < QE; ++q) {
        const size_t m = (size_t)q * F + f;
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0, b0 = 0, b1 = 0;
        if (DIFF) {
            const double2 r0 = reinterpret_cast<const double2*>(A.c.if_a)[2 * m], r1 = reinterpret_cast<const double2*>(A.c.if_a)[2 * m + 1];
            a00 = r0.x; a01 = r0.y; a10 = r1.x; a11 = r1.y;
        }
        if (ADV) {
            const double2 bb = reinterpret_cast<const double2*>(A.c.if_b)[m];
            b0 = bb.x; b1 = bb.y;
        }
        const double x = g.midx + A.xe[q] * g.tanx;
        const double y = g.midy + A.xe[q] * g.tany;
        const double W = g.De * A.we[q];
        double atn0 = 0, atn1 = 0, gamma = 0;
        if (DIFF) {
            atn0 = g.nx * a00 + g.ny * a10;
            atn1 = g.nx * a01 + g.ny * a11;
        }
        if (ADV) {
            const double beta = b0 * g.nx + b1 * g.ny;
            gamma = down ? -sgn * beta : 0.0;
        }
        const double hw0 = -0.5 * W * atn0, hw1 = -0.5 * W * atn1;
        const double sg = g.sigma + gamma;
        double phi[D], gx[D], gy[D];
        eval_basis<P, DIFF>(x, y, bxe, A.lc, phi, gx, gy);
        double Pr[RJ], Qr[RJ];
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) {
            const int j = J0 + jj;
            Pr[jj] = sgn * phi[j];
            Qr[jj] = 0.0;
            if (DIFF) {
                if (gx_nz<P>(j) && gy_nz<P>(j)) Qr[jj] = -0.5 * (atn0 * gx[j] + atn1 * gy[j]);
                else if (gx_nz<P>(j)) Qr[jj] = -0.5 * (atn0 * gx[j]);
                else if (gy_nz<P>(j)) Qr[jj] = -0.5 * (atn1 * gy[j]);
            }
        }
        const double Ws = OFF ? -(W * sgn) : W * sgn;
        if (OFF) eval_basis<P, DIFF>(x, y, bxo, A.lc, phi, gx, gy);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double Vc = Ws * phi[i];
            double Uc = sg * Vc;
            if (DIFF && gy_nz<P>(i)) Uc = fma(hw1, gy[i], Uc);
            if (DIFF && gx_nz<P>(i)) Uc = fma(hw0, gx[i], Uc);
#pragma unroll
            for (int jj = 0; jj < RJ; ++jj) {
                const int j = J0 + jj;
                double v = acc[jj * D + i];
                v = fma(Pr[jj], Uc, v);
                if (DIFF && (gx_nz<P>(j) || gy_nz<P>(j))) v = fma(Qr[jj], Vc, v);
                acc[jj * D + i] = v;
            }
        }
    }
}

// one row part of one chunk of items of a giant element
template <int P, int RS, int PART, bool DIFF, bool ADV>
__device__ __forceinline__ void giant_part(const AsmArgs& A, double2* __restrict__ part, int lane, bool active, int nit, int f,
                                           const FaceGeo& g, double sgn, bool down, const ElemBox& bxe, const ElemBox& bxo,
                                           double* __restrict__ row0, int rowlen, int slot, int depth, int maxdepth,
                                           double2& drun) {
    using DM = Dm<P, RS>;
    constexpr int D = DM::D, RJ = DM::RJ, NB = DM::NB, J0 = PART * RJ;
    constexpr int NPQ = (NB + 1) / 2;
    {
        double acc[2 * NPQ];
#pragma unroll
        for (int v = 0; v < 2 * NPQ; ++v) acc[v] = 0.0;
        if (active) giant_pass<P, RS, PART, DIFF, ADV, false>(A, f, g, sgn, down, bxe, bxe, acc);
        __syncwarp();
        if (active) {
#pragma unroll
            for (int kk = 0; kk < NPQ; ++kk) part[kk * kPartStride + lane] = make_double2(acc[2 * kk], acc[2 * kk + 1]);
        }
        __syncwarp();
        if (lane < NPQ) {
            for (int l = 0; l < nit; ++l) {         // faces in item order, after the volume sums
                const double2 t = part[lane * kPartStride + l];
                drun.x += t.x;
                drun.y += t.y;
            }
        }
    }
    {
        double acc[NB];
#pragma unroll
        for (int v = 0; v < NB; ++v) acc[v] = 0.0;
        if (active) giant_pass<P, RS, PART, DIFF, ADV, true>(A, f, g, sgn, down, bxe, bxo, acc);
        double* dst0 = row0 + slot * D;
        if (active && depth == 0) {
#pragma unroll
            for (int jj = 0; jj < RJ; ++jj)
#pragma unroll
                for (int i = 0; i < D; ++i) dst0[(J0 + jj) * rowlen + i] = acc[jj * D + i];
        }
        for (int dd = 1; dd <= maxdepth; ++dd) {    // repeats of a neighbour add to its block, in item order
            __threadfence_block();
            __syncwarp();
            if (active && depth == dd) {
#pragma unroll
                for (int jj = 0; jj < RJ; ++jj)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        double* q = dst0 + (J0 + jj) * rowlen + i;
                        *q = __ldcg(q) + acc[jj * D + i];
                    }
            }
        }
    }
}

template <int P, int RS, bool DIFF, bool ADV>
__global__ void __launch_bounds__(kWarp) dg_face_giant_kernel(const __grid_constant__ AsmArgs A) {
    using DM = Dm<P, RS>;
    constexpr int D = DM::D, DD = DM::DD, NV = DM::NV, NB = DM::NB;
    constexpr int NPQ = (NB + 1) / 2;
    __shared__ double2 part[NPQ * kPartStride];
    const int lane = threadIdx.x;
    const int e_lane = A.elem_lo + blockIdx.x * kWarp + lane;
    const bool big = e_lane < A.elem_hi && A.adj_off[e_lane + 1] - A.adj_off[e_lane] > kWarp;
    unsigned todo = __ballot_sync(kFull, big);
    unsigned zeros = 0;
    while (todo) {
        const int e = A.elem_lo + blockIdx.x * kWarp + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int k0 = A.adj_off[e], k1 = A.adj_off[e + 1], g0 = A.grp_off[e];
        const int rowlen = (A.grp_off[e + 1] - g0 + 1) * D;
        double* row0 = A.data + (size_t)DD * ((size_t)g0 + e);
        const int eg = gid_of(A, e);
        const ElemBox bxe = A.ebox[e];
        const int s0 = seg0_of(A, e), s1 = seg0_of(A, e + 1);
        // running sums of the diagonal block (+ load): lane kk owns value pair kk of every part; volume segments first
        double2 drun[RS];
#pragma unroll
        for (int part_i = 0; part_i < RS; ++part_i) {
            drun[part_i] = make_double2(0.0, 0.0);
            const int vi = part_i == 0 ? pair_dst<P, RS, 0>(lane) : pair_dst<P, RS, RS - 1>(lane);
            if (lane < NPQ || (part_i == RS - 1 && vi < NV)) {
                for (int s = s0; s < s1; ++s) {
                    if (vi < NV) drun[part_i].x += A.vdiag[(size_t)s * NV + vi];
                    if (vi + 1 < NV) drun[part_i].y += A.vdiag[(size_t)s * NV + vi + 1];
                }
            }
        }
        int heads_before = 0, dslot = 0;
        for (int kc = k0; kc < k1; kc += kWarp) {
            const int nit = min(kWarp, k1 - kc);
            const bool active = lane < nit;
            const int item = active ? A.adj_item[kc + lane] : 0;
            const int nbr = active ? A.adj_nb[kc + lane] : 0;
            const int f = (item & 0x7fffffff) >> 1;
            const int side = item & 1;
            const int og = active ? gid_of(A, nbr) : 0;
            const bool headf = active && item >= 0;
            const unsigned hm = __ballot_sync(kFull, headf);
            const unsigned below = (2u << lane) - 1u;
            const int slot = heads_before + __popc(hm & below) - 1 + (og > eg ? 1 : 0);
            const int depth = active ? lane - (31 - __clz(hm & below)) : 0;      // a repeat of the previous chunk's neighbour: lane + 1
            const int maxdepth = A.has_dups ? __reduce_max_sync(kFull, depth) : 0;
            dslot += __popc(__ballot_sync(kFull, headf && og < eg));
            heads_before += __popc(hm);
            FaceGeo g;
            ElemBox bxo = bxe;
            double sgn = side ? -1.0 : 1.0;
            bool down = false;
            if (active) {
                const FaceRec r = A.frec[f];
                g.midx = r.midx; g.midy = r.midy; g.tanx = r.tanx; g.tany = r.tany; g.nx = r.nx; g.ny = r.ny; g.De = r.De;
                g.sigma = A.fsigma ? A.fsigma[f] : r.sigma;
                bxo = A.ebox[nbr];
                down = A.c.if_upwind ? ((A.c.if_upwind[f] != 0) == (side == 1)) : false;
            } else {
                g.midx = g.midy = g.tanx = g.tany = g.nx = g.ny = g.De = g.sigma = 0.0;
            }
            giant_part<P, RS, 0, DIFF, ADV>(A, part, lane, active, nit, f, g, sgn, down, bxe, bxo, row0, rowlen, slot, depth, maxdepth,
                                            drun[0]);
            if (RS > 1)
                giant_part<P, RS, RS - 1, DIFF, ADV>(A, part, lane, active, nit, f, g, sgn, down, bxe, bxo, row0, rowlen, slot, depth,
                                                     maxdepth, drun[RS - 1]);
            __syncwarp();
        }
        // diagonal block and load
#pragma unroll
        for (int part_i = 0; part_i < RS; ++part_i) {
            const int vi = part_i == 0 ? pair_dst<P, RS, 0>(lane) : pair_dst<P, RS, RS - 1>(lane);
            const bool mine = part_i == 0 ? (RS == 1 ? lane < NV / 2 : lane < NPQ) : (lane < NPQ + D / 2);
            if (mine) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int v = vi + h;
                    const double val = h ? drun[part_i].y : drun[part_i].x;
                    if (v < DD) row0[(v / D) * rowlen + dslot * D + (v % D)] = val;
                    else if (v < NV) A.load[(size_t)e * D + (v - DD)] = val;
                }
            }
        }
        // exact zeros of the finished rows (they were written by this warp: read them back past L1)
        __threadfence_block();
        __syncwarp();
        for (int w = lane; w < D * rowlen; w += kWarp) zeros += (__ldcg(row0 + w) == 0.0);
    }
    if (zeros) atomicAdd(A.zero_count, (unsigned long long)zeros);
}

// =======================================================================================================================
template <int P, int RS, bool DIFF, bool ADV>
static int launch(const AsmArgs& A, cudaStream_t st, int phases) {
    using VS = VolSmem<P, RS, DIFF, ADV>;
    using FS = FaceSmem<P, RS, DIFF, ADV>;
    auto vk = dg_volume_kernel<P, RS, DIFF, ADV>;
    auto fk = dg_face_kernel<P, RS, DIFF, ADV>;
    RB_CUDA(cudaFuncSetAttribute(vk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VS::TOTAL));
    RB_CUDA(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS::TOTAL));
    if ((phases & 1) && A.vround_hi > A.vround_lo) {
        const int nrounds = A.vround_hi - A.vround_lo;
        RB_CUDA(cudaMemsetAsync(A.vol_next, 0, sizeof(int), st));
        // persistent warps: exactly as many one-warp CTAs as are resident at once (11 per SM at p <= 2, 6 at p = 3)
        static int resident = 0;
        if (resident == 0) {
            int per_sm = 0;
            RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vk, kWarp, VS::TOTAL));
            resident = kNumSMs * (per_sm > 0 ? per_sm : 1);
            if (resident > kVolWarps) resident = kVolWarps;
        }
        vk<<<nrounds < resident ? nrounds : resident, kWarp, VS::TOTAL, st>>>(A);
        RB_LAUNCH_CHECK("dg_volume_kernel");
    }
    if (!(phases & 2)) return 0;
    const int blocks = (int)div_up(A.elem_hi - A.elem_lo, A.window);
    fk<<<blocks, kWarp, FS::TOTAL, st>>>(A);
    RB_LAUNCH_CHECK("dg_face_kernel");
    if (A.has_giants) {
        dg_face_giant_kernel<P, RS, DIFF, ADV><<<(int)div_up(A.elem_hi - A.elem_lo, kWarp), kWarp, 0, st>>>(A);
        RB_LAUNCH_CHECK("dg_face_giant_kernel");
    }
    return 0;
}

template <int P, int RS>
static int dispatch(const AsmArgs& A, bool diff, bool adv, cudaStream_t st, int phases) {
    if (diff && adv) return launch<P, RS, true, true>(A, st, phases);
    if (diff) return launch<P, RS, true, false>(A, st, phases);
    return launch<P, RS, false, true>(A, st, phases);
}

}  // namespace items

int items_max_face_items() { return items::kWarp; }
int items_max_segments() { return items::kMaxSeg; }

int launch_assemble_items(const AsmArgs& A0, int p, bool diff, bool adv, cudaStream_t st, int phases) {
    // elements per face CTA: at least ~3 waves of CTAs on the GPU (8 one-warp CTAs per SM) so that the last wave does not
    // dominate small partitions, 16..64 elements (131 072 elements: 0.297 / 0.287 / 0.295 ms per call with 16 / 32 / 64)
    AsmArgs A = A0;
    const int n = A.elem_hi - A.elem_lo;
    int w = items::kWindow;
    while (w > 16 && (int64_t)n < (int64_t)w * 3 * 8 * kNumSMs) w >>= 1;
    if (const char* env = getenv("REYNA_B200_FACE_WINDOW")) {      // tuning aid: 16, 32 or 64
        const int v = atoi(env);
        if (v == 16 || v == 32 || v == 64) w = v;
    }
    A.window = w;
    switch (p) {
        case 1: return items::dispatch<1, 1>(A, diff, adv, st, phases);
        case 2: return items::dispatch<2, 1>(A, diff, adv, st, phases);
        default: return items::dispatch<3, 2>(A, diff, adv, st, phases);
    }
}

}  // namespace rb
