// p-multigrid + aggregation multigrid preconditioner for the DG system (used inside the Krylov solver that replaces
// scipy.sparse.linalg.spsolve, reyna/DGFEM/two_dimensional/main.py:231).
//
// Block-Jacobi alone needs O(1/h) Krylov iterations (44 000 at 262 144 elements, p = 2).  The tensor-Legendre basis is
// hierarchical, so coarse spaces are sub-blocks of B:
//   fine (p, d x d blocks)  --select modes (0,0),(0,1),(1,0)-->  P1 (3 x 3 blocks)  --select mode (0,0)-->  P0 (scalar,
//   one unknown per element: a two-point-flux finite-volume matrix)  --aggregate 4 consecutive elements-->  ...  --> dense
// and all coarse operators are Galerkin (R A R^T with R a selection / piecewise-constant aggregation).  Elements are
// assumed to be numbered with spatial locality (the benchmark meshes follow a Morton curve), so consecutive elements
// form compact aggregates.
// Preconditioner application z = M^-1 r (a fixed linear operator, symmetric when A is):
//   z  = D_F^-1 r                                  block Jacobi on the fine level      (additive: no fine SpMV)
//   z += P1^T V1(P1 r)                             V1: block-Jacobi sweep, P0 correction, block-Jacobi sweep  (3 x 3 blocks)
//   P0 correction: V-cycle over the aggregation levels with `nu` damped-Jacobi sweeps, coarse corrections scaled by
//   `alpha` (unsmoothed aggregation under-corrects), exact dense solve at the coarsest level.
// With element bounding boxes (rb_mg_opts.elem_bbox) the levels below P1 are AGGLOMERATED P1 spaces instead (groups of 4
// consecutive elements, linear polynomials on the bounding box of the group, 3 x 3 block Jacobi, V(nu,nu)): linear functions
// are represented on every level, which the piecewise-constant chain cannot do (2-3 times fewer Krylov iterations).
// Iteration counts become mesh independent (p = 2: 250 ... 490 BiCGStab / 400 ... 730 CG iterations down to the rounding floor
// of the true residual between 65k and 1M elements; block Jacobi alone: 21 500 / 44 050 at 65k / 262k).
#include "common.cuh"
#include <math.h>
#include <vector>

namespace rb {

constexpr int kMgThreads = 256;
constexpr int kMaxLevels = 16;
constexpr int kDenseMax = 256;     // coarsest level: dense inverse (Gauss-Jordan in one CTA)

struct MgLevel {
    int n = 0;
    int64_t nnz = 0;
    int32_t* indptr = nullptr;
    int32_t* indices = nullptr;
    double* data = nullptr;
    double* dinv = nullptr;
    double *b = nullptr, *x = nullptr, *xt = nullptr;   // right-hand side, iterate, ping-pong iterate
    // agglomerated-P1 chain (3 x 3 blocks; n = 3 * n_blk scalar rows)
    int n_blk = 0;
    double* minv = nullptr;      // [n_blk][3][3]  omega * inverse of the diagonal block
    double* bbox = nullptr;      // [n_blk][4]     bounding box of the (agglomerated) element
    double* tr = nullptr;        // [n_blk][8]     its P1 basis expressed in the basis of its aggregate (next level)
    float* dataf = nullptr;      // [nnz] single-precision copy of `data` for the smoothing sweeps (rb_mg_opts.fp32_levels)
};

}  // namespace rb

struct rb_dist;

struct rb_mg {
    int n_f = 0, d = 0, p = 0, n_elem = 0;
    const int32_t *f_indptr = nullptr, *f_indices = nullptr;
    const double* f_data = nullptr;
    // P1 level (3 x 3 blocks), present when d > 3
    int has_p1 = 0;
    int32_t *p1_indptr = nullptr, *p1_indices = nullptr;
    double *p1_data = nullptr, *p1_minv = nullptr, *p1_r = nullptr, *p1_x = nullptr, *p1_t = nullptr;
    int32_t* bcol = nullptr;       // block columns (element ids) of the fine / P1 block pattern
    double* p1_t0 = nullptr;       // distributed mode: this rank's P0 right-hand side before the all-gather
    int sel1[3] = {0, 1, 2};
    // P0 chain
    int nlev = 0;
    rb::MgLevel lev[rb::kMaxLevels];
    double* dense_inv = nullptr;
    int nu = 3;
    double omega = 0.7, alpha = 1.8;
    // distributed mode: fine and P1 levels are partitioned (halo exchange), the P0 chain is global and replicated
    rb_dist* dist = nullptr;
    int n_ghost_elem = 0, lo_elem = 0;
    std::vector<int32_t> bounds, bounds_l1;
    std::vector<void*> owned;      // arena chunks (cudaMalloc), released by reyna_b200_mg_destroy
    char* arena_cur = nullptr;     // the hierarchy sub-allocates its ~100 arrays from a few large chunks: cudaMalloc costs
    size_t arena_left = 0;         // 0.1 ... 10 ms per call on a busy driver, which used to dominate the set-up of small solves
    size_t arena_chunk = 8u << 20;
    // agglomerated-P1 chain (rb_mg_opts.elem_bbox given): blev[0] is the P1 level itself
    int agg_p1 = 0, nblev = 0;
    rb::MgLevel blev[rb::kMaxLevels];
};

namespace rb {

static int dev_alloc(rb_mg* mg, void** ptr, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;        // cudaMalloc's alignment
    if (bytes == 0) bytes = 256;
    if (bytes > mg->arena_left) {
        const size_t chunk = bytes > mg->arena_chunk ? bytes : mg->arena_chunk;
        void* p = nullptr;
        RB_CUDA(cudaMalloc(&p, chunk));
        mg->owned.push_back(p);
        mg->arena_cur = static_cast<char*>(p);
        mg->arena_left = chunk;
    }
    *ptr = mg->arena_cur;
    mg->arena_cur += bytes;
    mg->arena_left -= bytes;
    return 0;
}
// ---------------------------------------------------------------------------------------------------------------
// setup kernels
// ---------------------------------------------------------------------------------------------------------------
// Sub-block extraction: coarse row (e, jc) takes, from every d x d block of block row e, the entries (sel[jc], sel[ic]).
// The fine matrix must carry the structural block pattern (every scalar row of a block row has the same columns, d
// consecutive columns per neighbouring element).
__global__ void mg_extract_modes(int n_elem, int d, int dc, int s0, int s1, int s2, const int32_t* __restrict__ indptr,
                                 const int32_t* __restrict__ indices, const double* __restrict__ data,
                                 int32_t* __restrict__ c_indptr, int32_t* __restrict__ c_indices,
                                 double* __restrict__ c_data) {
    const int sel[3] = {s0, s1, s2};
    const int total = n_elem * dc;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        const int e = r / dc, jc = r - e * dc;
        const int f0 = indptr[e * d];
        const int rowlen = indptr[e * d + 1] - f0;
        const int nb = rowlen / d;
        const int64_t blocks_before = (int64_t)f0 / (d * d);
        const int c0 = (int)(blocks_before * dc * dc) + jc * nb * dc;
        c_indptr[r] = c0;
        const int frow = indptr[e * d + sel[jc]];
        for (int s = 0; s < nb; ++s) {
            const int celem = indices[f0 + s * d] / d;
            for (int ic = 0; ic < dc; ++ic) {
                c_indices[c0 + s * dc + ic] = celem * dc + ic;
                c_data[c0 + s * dc + ic] = data[frow + s * d + sel[ic]];
            }
        }
        if (r == total - 1) c_indptr[total] = c0 + nb * dc;
    }
}

// Galerkin aggregation A_c = P^T A P with P piecewise constant over groups of `agg` consecutive rows.  Each fine row has
// ascending columns, so the mapped columns c / agg are non-decreasing: a k-way merge yields the coarse row sorted.
template <bool FILL>
__global__ void mg_aggregate(int nc, int n, int agg, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                             const double* __restrict__ data, int32_t* __restrict__ c_cnt_or_indptr,
                             int32_t* __restrict__ c_indices, double* __restrict__ c_data, double* __restrict__ c_dinv) {
    for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += gridDim.x * blockDim.x) {
        int pos[8], end[8];
        const int r0 = I * agg;
        int nr = n - r0 < agg ? n - r0 : agg;
        // Ghost columns (>= n; a partitioned matrix orders its blocks by GLOBAL element id, so they can sit anywhere in a row)
        // are skipped; the remaining owned columns are ascending.
        for (int k = 0; k < nr; ++k) {
            pos[k] = indptr[r0 + k];
            end[k] = indptr[r0 + k + 1];
            while (pos[k] < end[k] && indices[pos[k]] >= n) ++pos[k];
        }
        int out = FILL ? c_cnt_or_indptr[I] : 0;
        int count = 0;
        for (;;) {
            int cmin = 0x7fffffff;
            for (int k = 0; k < nr; ++k)
                if (pos[k] < end[k]) { const int c = indices[pos[k]] / agg; cmin = c < cmin ? c : cmin; }
            if (cmin == 0x7fffffff) break;
            double v = 0.0;
            for (int k = 0; k < nr; ++k) {        // fixed order: rows ascending, entries ascending -> deterministic
                while (pos[k] < end[k] && indices[pos[k]] / agg == cmin) {
                    if (FILL) v += data[pos[k]];
                    ++pos[k];
                    while (pos[k] < end[k] && indices[pos[k]] >= n) ++pos[k];
                }
            }
            if (FILL) {
                c_indices[out] = cmin;
                c_data[out] = v;
                if (cmin == I) c_dinv[I] = v != 0.0 ? 1.0 / v : 1.0;
                ++out;
            }
            ++count;
        }
        if (!FILL) c_cnt_or_indptr[I] = count;
    }
}

__global__ void mg_diag_inv(int n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const double* __restrict__ data, double* __restrict__ dinv) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        double dv = 1.0;
        for (int k = indptr[r]; k < indptr[r + 1]; ++k)
            if (indices[k] == r) dv = data[k];
        dinv[r] = dv != 0.0 ? 1.0 / dv : 1.0;
    }
}

// dense inverse of the coarsest operator: one CTA of 1024 threads, Gauss-Jordan with partial pivoting on [A | I].  Both
// matrices are held COLUMN-major (at[k*n + r] = A[r][k]): thread (r, slice) updates the columns k = slice, slice + 4, ... of
// row r, coalesced over r.  The pivot is the first row with the largest magnitude (block arg-max, ties to the smaller row), so
// the operations on every entry and their order do not depend on the thread layout.
constexpr int kInvThreads = 1024;
constexpr int kInvSlices = kInvThreads / kDenseMax;
__global__ void __launch_bounds__(kInvThreads) mg_dense_inverse(int n, const int32_t* __restrict__ indptr,
                                                                 const int32_t* __restrict__ indices, const double* __restrict__ data,
                                                                 double* __restrict__ at, double* __restrict__ invt) {
    __shared__ double s_best[kInvThreads / 32];
    __shared__ int s_row[kInvThreads / 32];
    __shared__ int s_piv;
    __shared__ double s_ip;
    __shared__ double s_prow[2 * kDenseMax];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n * n; i += nt) { at[i] = 0.0; invt[i] = (i / n == i % n) ? 1.0 : 0.0; }
    __syncthreads();
    for (int r = tid; r < n; r += nt)
        for (int k = indptr[r]; k < indptr[r + 1]; ++k)
            if (indices[k] < n) at[(size_t)indices[k] * n + r] = data[k];
    __syncthreads();
    const int row = tid % kDenseMax, slice = tid / kDenseMax;
    for (int c = 0; c < n; ++c) {
        // pivot: first row r >= c with the largest |A[r][c]|
        double best = -1.0;
        int br = n;
        for (int r = c + tid; r < n; r += nt) {
            const double v = fabs(at[(size_t)c * n + r]);
            if (v > best) { best = v; br = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int orow = __shfl_xor_sync(0xffffffffu, br, o);
            if (ob > best || (ob == best && orow < br)) { best = ob; br = orow; }
        }
        if (lane == 0) { s_best[warp] = best; s_row[warp] = br; }
        __syncthreads();
        if (warp == 0) {
            best = lane < nt / 32 ? s_best[lane] : -1.0;
            br = lane < nt / 32 ? s_row[lane] : n;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int orow = __shfl_xor_sync(0xffffffffu, br, o);
                if (ob > best || (ob == best && orow < br)) { best = ob; br = orow; }
            }
            if (lane == 0) {
                s_piv = br;
                const double pv = at[(size_t)c * n + br];
                s_ip = pv != 0.0 ? 1.0 / pv : 1.0;
            }
        }
        __syncthreads();
        const int piv = s_piv;
        const double ip = s_ip;
        // swap rows c <-> piv and scale the pivot row (thread k handles column k)
        for (int k = tid; k < n; k += nt) {
            double ac = at[(size_t)k * n + c], ap = at[(size_t)k * n + piv];
            double ic = invt[(size_t)k * n + c], ipv = invt[(size_t)k * n + piv];
            if (piv != c) { at[(size_t)k * n + piv] = ac; invt[(size_t)k * n + piv] = ic; ac = ap; ic = ipv; }
            at[(size_t)k * n + c] = ac * ip;
            invt[(size_t)k * n + c] = ic * ip;
            if (k < kDenseMax) { s_prow[k] = ac * ip; s_prow[kDenseMax + k] = ic * ip; }      // the scaled pivot row, for the elimination
        }
        __syncthreads();
        // eliminate column c from all other rows: the multipliers are read before any slice touches column c
        for (int r0 = 0; r0 < n; r0 += kDenseMax) {         // one pass when n <= kDenseMax (always, in practice)
            const int r = r0 + row;
            double f = 0.0;
            if (r < n && r != c) f = at[(size_t)c * n + r];
            __syncthreads();
            if (f != 0.0) {
                if (n <= kDenseMax) {
#pragma unroll 4
                    for (int k = slice; k < n; k += kInvSlices) {
                        at[(size_t)k * n + r] -= f * s_prow[k];
                        invt[(size_t)k * n + r] -= f * s_prow[kDenseMax + k];
                    }
                } else {
                    for (int k = slice; k < n; k += kInvSlices) {
                        at[(size_t)k * n + r] -= f * at[(size_t)k * n + c];
                        invt[(size_t)k * n + r] -= f * invt[(size_t)k * n + c];
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// cycle kernels
// ---------------------------------------------------------------------------------------------------------------
// x = omega * dinv * b   (first sweep from a zero guess)
__global__ void mg_jacobi0(int n, double omega, const double* __restrict__ dinv, const double* __restrict__ b,
                           double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = omega * dinv[i] * b[i];
}

// G lanes cooperate on one row (contiguous, coalesced reads of its entries); returns (A x)_row in every lane of the group
template <int G>
__device__ __forceinline__ double row_dot(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                          const double* __restrict__ data, const double* __restrict__ x, int row, int lane,
                                          unsigned gmask, int ncols) {
    // columns >= ncols are ghost couplings of a partitioned matrix: the (rank-local) multigrid levels ignore them
    double s = 0.0;
    const int hi = indptr[row + 1];
    for (int k = indptr[row] + lane; k < hi; k += G) {
        const int c = indices[k];
        if (c < ncols) s = fma(data[k], x[c], s);
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) s += __shfl_xor_sync(gmask, s, o, G);
    return s;
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
}

// xo = x + omega * dinv * (b - A x)
template <int G>
__global__ void __launch_bounds__(kMgThreads) mg_jacobi(int n, int ncols, double omega, const int32_t* __restrict__ indptr,
                                                        const int32_t* __restrict__ indices, const double* __restrict__ data,
                                                        const double* __restrict__ dinv, const double* __restrict__ b,
                                                        const double* __restrict__ x, double* __restrict__ xo) {
    const int lane = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += stride) {
        const double ax = row_dot<G>(indptr, indices, data, x, (int)i, lane, gm, ncols);
        if (lane == 0) xo[i] = x[i] + omega * dinv[i] * (b[i] - ax);
    }
}

// bc[I] = sum over the 4 rows i of aggregate I of (b - A x)_i : 4 lanes per fine row, 16 lanes per coarse row
__global__ void __launch_bounds__(kMgThreads) mg_restrict_residual(int nc, int n, const int32_t* __restrict__ indptr,
                                                                   const int32_t* __restrict__ indices,
                                                                   const double* __restrict__ data, const double* __restrict__ b,
                                                                   const double* __restrict__ x, double* __restrict__ bc) {
    const int lane16 = threadIdx.x & 15, sub = lane16 >> 2, lane = lane16 & 3;
    const unsigned gm16 = 0xffffu << ((threadIdx.x & 31) & 16);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / 16;
    for (int64_t I = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 16; I < nc; I += stride) {
        const int i = (int)I * 4 + sub;
        double s = 0.0;
        if (i < n) {
            const int hi = indptr[i + 1];
            for (int k = indptr[i] + lane; k < hi; k += 4) {
                const int c = indices[k];
                if (c < n) s = fma(-data[k], x[c], s);
            }
            if (lane == 0) s += b[i];
        }
        // fixed-shape tree over the 16 lanes: deterministic
#pragma unroll
        for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(gm16, s, o, 16);
        if (lane16 == 0) bc[I] = s;
    }
}

// x[i] += alpha * xc[i / agg]
__global__ void mg_prolong_add(int n, int agg, double alpha, const double* __restrict__ xc, double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += alpha * xc[i / agg];
}

// y = A^-1 x with the column-major inverse of mg_dense_inverse: a CTA of 256 threads = 8 k-slices x 32 rows, the partial
// sums of the slices are combined in a fixed order through shared memory (n <= 256: a handful of CTAs, ~2 us)
__global__ void __launch_bounds__(256) mg_dense_apply(int n, const double* __restrict__ invt, const double* __restrict__ x,
                                                      double* __restrict__ y) {
    __shared__ double s_part[8][33];
    const int rl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int r = blockIdx.x * 32 + rl;
    double s = 0.0;
    if (r < n)
        for (int k = sl; k < n; k += 8) s = fma(invt[(size_t)k * n + r], x[k], s);
    s_part[sl][rl] = s;
    __syncthreads();
    if (sl == 0 && r < n) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_part[q][rl];
        y[r] = t;
    }
}

// block level helpers (dc x dc diagonal-block inverse `minv`, rows r = e*dc + j)
// out = Minv v                      (ADD == false)   |   out += Minv v   (ADD == true)
template <bool ADD>
__global__ void mg_block_apply(int n, int dc, const double* __restrict__ minv, const double* __restrict__ v,
                               double* __restrict__ out) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int e = r / dc, j = r - e * dc;
        const double* m = minv + ((size_t)e * dc + j) * dc;
        const double* vv = v + (size_t)e * dc;
        double z = 0.0;
        for (int i = 0; i < dc; ++i) z = fma(m[i], vv[i], z);
        out[r] = ADD ? out[r] + z : z;
    }
}

// t = b - A x
template <int G>
__global__ void __launch_bounds__(kMgThreads) mg_residual(int n, int ncols, const int32_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices, const double* __restrict__ data,
                                                          const double* __restrict__ b, const double* __restrict__ x,
                                                          double* __restrict__ t) {
    const int lane = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += stride) {
        const double ax = row_dot<G>(indptr, indices, data, x, (int)i, lane, gm, ncols);
        if (lane == 0) t[i] = b[i] - ax;
    }
}

// t = b - A x for the 3 x 3-block P1 matrix without reading its column indices (block columns shared with the fine level)
template <int G>
__global__ void __launch_bounds__(kMgThreads) mg_residual_blocked3(int n, int ncols, const int32_t* __restrict__ indptr,
                                                                   const int32_t* __restrict__ bcol,
                                                                   const double* __restrict__ data, const double* __restrict__ b,
                                                                   const double* __restrict__ x, double* __restrict__ t) {
    const int lane = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / G;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += stride) {
        const int e = (int)(i / 3);
        const int lo = indptr[i], len = indptr[i + 1] - lo;
        const int32_t* bc = bcol + indptr[e * 3] / 9;
        const double* dr = data + lo;
        double s = 0.0;
        for (int k = lane; k < len; k += G) {
            const int sb = k / 3, c = bc[sb] * 3 + (k - sb * 3);
            if (c < ncols) s = fma(dr[k], x[c], s);
        }
#pragma unroll
        for (int o = G / 2; o; o >>= 1) s += __shfl_xor_sync(gm, s, o, G);
        if (lane == 0) t[i] = b[i] - s;
    }
}

// coarse[e*dc + jc] = fine[e*d + sel[jc]]
__global__ void mg_select(int n_elem, int d, int dc, int s0, int s1, int s2, const double* __restrict__ fine,
                          double* __restrict__ coarse) {
    const int sel[3] = {s0, s1, s2};
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_elem * dc; r += gridDim.x * blockDim.x) {
        const int e = r / dc, jc = r - e * dc;
        coarse[r] = fine[(size_t)e * d + sel[jc]];
    }
}

// fine[e*d + sel[jc]] += coarse[e*dc + jc]
__global__ void mg_pad_add(int n_elem, int d, int dc, int s0, int s1, int s2, const double* __restrict__ coarse,
                           double* __restrict__ fine) {
    const int sel[3] = {s0, s1, s2};
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_elem * dc; r += gridDim.x * blockDim.x) {
        const int e = r / dc, jc = r - e * dc;
        fine[(size_t)e * d + sel[jc]] += coarse[r];
    }
}


// ---------------------------------------------------------------------------------------------------------------
// agglomerated-P1 chain: every level keeps the three modes (0,0), (0,1), (1,0) per (agglomerated) element.  An aggregate
// is 4 consecutive elements, its basis the scaled Legendre polynomials of degree <= 1 on the bounding box of the group
// (the same construction as the element basis, assembly_aux.py:32-66), so that P1 functions are represented exactly
// on every level.  With k00 = c0^2 / sqrt(hx hy), ky = c0 c1 / (sqrt(hx hy) hy), kx likewise:
//   Phi_00 = (K00/k00) phi_00,   Phi_01 = (Ky/ky) phi_01 + Ky (my - My)/k00 phi_00,   Phi_10 = (Kx/kx) phi_10 + ... phi_00
// (capitals: aggregate, lower case: member element).  tr = {s00, syy, sy0, sxx, sx0}.
// ---------------------------------------------------------------------------------------------------------------
__global__ void mgb_coarsen_bbox(int nc, int n, const double* __restrict__ bf, double* __restrict__ bc) {
    for (int G = blockIdx.x * blockDim.x + threadIdx.x; G < nc; G += gridDim.x * blockDim.x) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int e = 4 * G; e < 4 * G + 4 && e < n; ++e) {
            x0 = fmin(x0, bf[4 * (size_t)e]); x1 = fmax(x1, bf[4 * (size_t)e + 1]);
            y0 = fmin(y0, bf[4 * (size_t)e + 2]); y1 = fmax(y1, bf[4 * (size_t)e + 3]);
        }
        bc[4 * (size_t)G] = x0; bc[4 * (size_t)G + 1] = x1; bc[4 * (size_t)G + 2] = y0; bc[4 * (size_t)G + 3] = y1;
    }
}

struct BoxK { double mx, my, k00, kx, ky; };
__device__ __forceinline__ BoxK box_k(const double* __restrict__ b) {
    const double hx = 0.5 * (b[1] - b[0]), hy = 0.5 * (b[3] - b[2]);
    BoxK k;
    k.mx = 0.5 * (b[1] + b[0]); k.my = 0.5 * (b[3] + b[2]);
    const double r = rsqrt(hx * hy);
    k.k00 = 0.5 * r;                                   // c0^2,  c0 = 1/sqrt(2)
    const double c01 = 0.8660254037844386 * r;         // c0 c1, c1 = sqrt(3/2)
    k.kx = c01 / hx; k.ky = c01 / hy;
    return k;
}

__global__ void mgb_transfer(int n, const double* __restrict__ bf, const double* __restrict__ bc, double* __restrict__ tr) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const BoxK f = box_k(bf + 4 * (size_t)e), c = box_k(bc + 4 * (size_t)(e >> 2));
        double* t = tr + 8 * (size_t)e;
        t[0] = c.k00 / f.k00;
        t[1] = c.ky / f.ky; t[2] = c.ky * (f.my - c.my) / f.k00;
        t[3] = c.kx / f.kx; t[4] = c.kx * (f.mx - c.mx) / f.k00;
        t[5] = t[6] = t[7] = 0.0;
    }
}

// M (3 x 3, row-major) <- Pe^T A Pf, with P = [[s00, sy0, sx0], [0, syy, 0], [0, 0, sxx]] (rows: element modes)
__device__ __forceinline__ void mgb_transform(const double* __restrict__ te, const double* __restrict__ tf, const double (&a)[9],
                                              double (&m)[9]) {
    const double pe[9] = {te[0], te[2], te[4], 0.0, te[1], 0.0, 0.0, 0.0, te[3]};
    const double pf[9] = {tf[0], tf[2], tf[4], 0.0, tf[1], 0.0, 0.0, 0.0, tf[3]};
    double ap[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) ap[i * 3 + j] = a[i * 3 + 0] * pf[0 * 3 + j] + a[i * 3 + 1] * pf[1 * 3 + j] + a[i * 3 + 2] * pf[2 * 3 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m[i * 3 + j] = pe[0 * 3 + i] * ap[0 * 3 + j] + pe[1 * 3 + i] * ap[1 * 3 + j] + pe[2 * 3 + i] * ap[2 * 3 + j];
}

// Galerkin coarse operator of one agglomeration step, one thread per coarse block row.  The fine matrix is a scalar CSR
// matrix with complete 3 x 3 blocks (block columns ascending); FILL == false counts the coarse blocks per row.
template <bool FILL>
__global__ void mgb_aggregate(int nc, int n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                              const double* __restrict__ data, const double* __restrict__ tr,
                              int32_t* __restrict__ blk_cnt_or_off, int32_t* __restrict__ c_indptr,
                              int32_t* __restrict__ c_indices, double* __restrict__ c_data) {
    for (int G = blockIdx.x * blockDim.x + threadIdx.x; G < nc; G += gridDim.x * blockDim.x) {
        int pos[4], end[4], lo[4];
        const int e0 = 4 * G;
        const int nr = n - e0 < 4 ? n - e0 : 4;
        for (int k = 0; k < nr; ++k) {
            lo[k] = indptr[3 * (e0 + k)];
            pos[k] = 0;
            end[k] = (indptr[3 * (e0 + k) + 1] - lo[k]) / 3;       // blocks of fine block row e0 + k
            while (pos[k] < end[k] && indices[lo[k] + 3 * pos[k]] / 3 >= n) ++pos[k];   // ghost couplings are skipped
        }
        int count = 0;
        if (!FILL) {
            for (;;) {
                int hmin = 0x7fffffff;
                for (int k = 0; k < nr; ++k)
                    if (pos[k] < end[k]) { const int h = indices[lo[k] + 3 * pos[k]] / 12; hmin = h < hmin ? h : hmin; }
                if (hmin == 0x7fffffff) break;
                for (int k = 0; k < nr; ++k)
                    while (pos[k] < end[k] && indices[lo[k] + 3 * pos[k]] / 12 == hmin) {
                        ++pos[k];
                        while (pos[k] < end[k] && indices[lo[k] + 3 * pos[k]] / 3 >= n) ++pos[k];
                    }
                ++count;
            }
            blk_cnt_or_off[G] = count;
            continue;
        }
        const int b0 = blk_cnt_or_off[G], nbG = blk_cnt_or_off[G + 1] - b0;
        const int row0 = 9 * b0;                                 // scalar row 3G + i starts at row0 + i * 3 * nbG
        c_indptr[3 * G] = row0; c_indptr[3 * G + 1] = row0 + 3 * nbG; c_indptr[3 * G + 2] = row0 + 6 * nbG;
        if (G == nc - 1) c_indptr[3 * nc] = row0 + 9 * nbG;
        for (;;) {
            int hmin = 0x7fffffff;
            for (int k = 0; k < nr; ++k)
                if (pos[k] < end[k]) { const int h = indices[lo[k] + 3 * pos[k]] / 12; hmin = h < hmin ? h : hmin; }
            if (hmin == 0x7fffffff) break;
            double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < nr; ++k) {                       // fixed order: rows ascending, blocks ascending
                const int e = e0 + k;
                const int len = indptr[3 * e + 1] - lo[k];       // scalar row length of the fine block row
                while (pos[k] < end[k] && indices[lo[k] + 3 * pos[k]] / 12 == hmin) {
                    const int f = indices[lo[k] + 3 * pos[k]] / 3;
                    double a[9], m[9];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) a[i * 3 + j] = data[lo[k] + i * len + 3 * pos[k] + j];
                    mgb_transform(tr + 8 * (size_t)e, tr + 8 * (size_t)f, a, m);
#pragma unroll
                    for (int v = 0; v < 9; ++v) acc[v] += m[v];
                    ++pos[k];
                    while (pos[k] < end[k] && indices[lo[k] + 3 * pos[k]] / 3 >= n) ++pos[k];
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    c_indices[row0 + i * 3 * nbG + 3 * count + j] = 3 * hmin + j;
                    c_data[row0 + i * 3 * nbG + 3 * count + j] = acc[i * 3 + j];
                }
            ++count;
        }
    }
}

__global__ void mgb_scale(int64_t n, double s, double* __restrict__ v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] *= s;
}

// bc[G] = sum over the members e of G of Pe^T t[e]   (fixed member order: deterministic)
__global__ void mgb_restrict(int nc, int n, const double* __restrict__ tr, const double* __restrict__ t, double* __restrict__ bc) {
    for (int G = blockIdx.x * blockDim.x + threadIdx.x; G < nc; G += gridDim.x * blockDim.x) {
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        for (int e = 4 * G; e < 4 * G + 4 && e < n; ++e) {
            const double* q = tr + 8 * (size_t)e;
            const double t0 = t[3 * (size_t)e], t1 = t[3 * (size_t)e + 1], t2 = t[3 * (size_t)e + 2];
            r0 = fma(q[0], t0, r0);
            r1 += fma(q[2], t0, q[1] * t1);
            r2 += fma(q[4], t0, q[3] * t2);
        }
        bc[3 * (size_t)G] = r0; bc[3 * (size_t)G + 1] = r1; bc[3 * (size_t)G + 2] = r2;
    }
}

// x[e] += alpha * Pe xc[G]
__global__ void mgb_prolong_add(int n, double alpha, const double* __restrict__ tr, const double* __restrict__ xc,
                                double* __restrict__ x) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const double* q = tr + 8 * (size_t)e;
        const double* c = xc + 3 * (size_t)(e >> 2);
        x[3 * (size_t)e] += alpha * (q[0] * c[0] + q[2] * c[1] + q[4] * c[2]);
        x[3 * (size_t)e + 1] += alpha * (q[1] * c[1]);
        x[3 * (size_t)e + 2] += alpha * (q[3] * c[2]);
    }
}

int launch_block_columns(int n_block_rows, int d, const int32_t* indptr, const int32_t* indices, int32_t* bcol,
                         cudaStream_t st);   // solver.cu
void block_jacobi_setup_launch(int n_blocks, int d, const int32_t* indptr, const int32_t* indices, const double* data,
                               double* minv, int* singular, cudaStream_t st);   // solver.cu

static inline int mg_grid(int64_t n) { return grid_for(n, kMgThreads, kNumSMs * 8); }

int dist_halo_d(rb_dist* h, double* x, int d, cudaStream_t st);                                           // dist.cu
int dist_allgather_rows(rb_dist* h, const double* local, double* global, const int32_t* bounds, cudaStream_t st);


// V-cycle on the P0 aggregation chain, level k: lev[k].b holds the right-hand side, the result is left in lev[k].x
static int mg_vcycle(rb_mg* mg, int k, cudaStream_t st) {
    MgLevel& L = mg->lev[k];
    if (k == mg->nlev - 1) {
        mg_dense_apply<<<(L.n + 31) / 32, 256, 0, st>>>(L.n, mg->dense_inv, L.b, L.x);
        RB_LAUNCH_CHECK("mg_dense_apply");
        return 0;
    }
    const int g = mg_grid(L.n);
    double *x = L.x, *xt = L.xt;
    mg_jacobi0<<<g, kMgThreads, 0, st>>>(L.n, mg->omega, L.dinv, L.b, x);
    for (int s = 1; s < mg->nu; ++s) {
        mg_jacobi<4><<<mg_grid((int64_t)L.n * 4), kMgThreads, 0, st>>>(L.n, L.n, mg->omega, L.indptr, L.indices, L.data, L.dinv, L.b, x, xt);
        double* t = x; x = xt; xt = t;
    }
    RB_LAUNCH_CHECK("mg_jacobi (pre)");
    MgLevel& C = mg->lev[k + 1];
    mg_restrict_residual<<<mg_grid((int64_t)C.n * 16), kMgThreads, 0, st>>>(C.n, L.n, L.indptr, L.indices, L.data, L.b, x, C.b);
    RB_LAUNCH_CHECK("mg_restrict_residual");
    int rc = mg_vcycle(mg, k + 1, st);
    if (rc) return rc;
    mg_prolong_add<<<g, kMgThreads, 0, st>>>(L.n, 4, mg->alpha, C.x, x);
    for (int s = 0; s < mg->nu; ++s) {
        mg_jacobi<4><<<mg_grid((int64_t)L.n * 4), kMgThreads, 0, st>>>(L.n, L.n, mg->omega, L.indptr, L.indices, L.data, L.dinv, L.b, x, xt);
        double* t = x; x = xt; xt = t;
    }
    RB_LAUNCH_CHECK("mg_jacobi (post)");
    if (x != L.x) {   // odd number of swaps: the result sits in the ping-pong buffer
        RB_CUDA(cudaMemcpyAsync(L.x, x, sizeof(double) * L.n, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}



// One sweep on a 3 x 3-block level, 8 lanes per BLOCK row (its three scalar rows are contiguous and share their columns):
// every x value is gathered once and used for the three rows.  SWEEP == false: out = b - A x;  SWEEP == true: the damped block
// Jacobi update out = x + Minv (b - A x) (omega folded into minv), written to another buffer than x.  Columns >= ncols are
// ghost couplings the level ignores.  Fixed shuffle tree: deterministic.
#ifndef RB_MGB_MINB
#define RB_MGB_MINB 1
#endif
template <bool SWEEP, typename T>
__global__ void __launch_bounds__(kMgThreads, RB_MGB_MINB) mgb_block_row(int n_blk, int ncols, const int32_t* __restrict__ indptr,
                                                            const int32_t* __restrict__ indices, const T* __restrict__ data,
                                                            const double* __restrict__ minv, const double* __restrict__ b,
                                                            const double* __restrict__ x, double* __restrict__ out,
                                                            const int32_t* __restrict__ bcol) {
    // bcol != null: one block column per 3 x 3 block (the P1 level shares the element pattern of the fine matrix) instead of
    // one column index per scalar entry -- a third of the index bytes
    const int lane = threadIdx.x & 7;
    const unsigned gm = 0xffu << ((threadIdx.x & 31) & ~7);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / 8;
    int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
    int lo_n = 0, len_n = 0;
    if (e < n_blk) {
        lo_n = indptr[3 * e];
        len_n = indptr[3 * e + 1] - lo_n;
    }
    for (; e < n_blk; e += stride) {
        // the dependent chain of a block row is row pointer -> column indices -> x -> tail operands: the row pointer of the NEXT
        // row and the tail operands (b, Minv, own x) are requested up front, so only indices -> x remains serial
        const int lo = lo_n, len = len_n;
        if (e + stride < n_blk) {
            lo_n = indptr[3 * (e + stride)];
            len_n = indptr[3 * (e + stride) + 1] - lo_n;
        }
        double b0 = 0.0, b1 = 0.0, b2 = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0, xo = 0.0;
#ifdef RB_MGB_LATE_TAIL
        if (false) {
#else
        if (lane < 3) {
#endif
            b0 = b[3 * e]; b1 = b[3 * e + 1]; b2 = b[3 * e + 2];
            if (SWEEP) {
                const double* m = minv + 9 * (size_t)e + 3 * lane;
                m0 = m[0]; m1 = m[1]; m2 = m[2];
                xo = x[3 * e + lane];
            }
        }
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#ifndef RB_MGB_NO_UNROLL
        // three column slots per lane and pass (24 scalar columns = 8 blocks cover a typical row in ONE pass): all index loads,
        // then all gathers and matrix values are in flight together; a skipped slot adds an exact fma(0, 0, s) = s, so the sums
        // are bitwise those of the one-slot loop
        for (int j0 = lane; j0 < len; j0 += 24) {
            int c[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int j = j0 + 8 * u;
                if (j >= len) c[u] = ncols;
                else if (bcol) c[u] = bcol[lo / 9 + j / 3] * 3 + j % 3;
                else c[u] = indices[lo + j];
            }
            double xv[3], d0[3], d1[3], d2[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const bool on = c[u] < ncols;
                const int j = j0 + 8 * u;
                xv[u] = on ? x[c[u]] : 0.0;
                d0[u] = on ? (double)data[lo + j] : 0.0;
                d1[u] = on ? (double)data[lo + len + j] : 0.0;
                d2[u] = on ? (double)data[lo + 2 * len + j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                s0 = fma(d0[u], xv[u], s0);
                s1 = fma(d1[u], xv[u], s1);
                s2 = fma(d2[u], xv[u], s2);
            }
        }
#else
        for (int j = lane; j < len; j += 8) {
            const int c = bcol ? bcol[lo / 9 + j / 3] * 3 + j % 3 : indices[lo + j];
            if (c < ncols) {
                const double xv = x[c];
                s0 = fma((double)data[lo + j], xv, s0);
                s1 = fma((double)data[lo + len + j], xv, s1);
                s2 = fma((double)data[lo + 2 * len + j], xv, s2);
            }
        }
#endif
#pragma unroll
        for (int o = 4; o; o >>= 1) {
            s0 += __shfl_xor_sync(gm, s0, o, 8);
            s1 += __shfl_xor_sync(gm, s1, o, 8);
            s2 += __shfl_xor_sync(gm, s2, o, 8);
        }
        if (lane < 3) {
#ifdef RB_MGB_LATE_TAIL
            b0 = b[3 * e]; b1 = b[3 * e + 1]; b2 = b[3 * e + 2];
            if (SWEEP) {
                const double* m = minv + 9 * (size_t)e + 3 * lane;
                m0 = m[0]; m1 = m[1]; m2 = m[2];
                xo = x[3 * e + lane];
            }
#endif
            const double r0 = b0 - s0, r1 = b1 - s1, r2 = b2 - s2;
            if (SWEEP) {
                out[3 * e + lane] = xo + fma(m0, r0, fma(m1, r1, m2 * r2));
            } else {
                out[3 * e + lane] = lane == 0 ? r0 : (lane == 1 ? r1 : r2);
            }
        }
    }
}

__global__ void mgb_to_float(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

// V-cycle on the agglomerated-P1 chain, level k: blev[k].b holds the right-hand side, the result is left in blev[k].x.
// Smoother: nu sweeps of 3 x 3 block Jacobi (the damping is folded into minv), coarse corrections scaled by alpha.
static int mgb_vcycle(rb_mg* mg, int k, cudaStream_t st) {
    MgLevel& L = mg->blev[k];
    if (k == mg->nblev - 1) {
        mg_dense_apply<<<(L.n + 31) / 32, 256, 0, st>>>(L.n, mg->dense_inv, L.b, L.x);
        RB_LAUNCH_CHECK("mg_dense_apply");
        return 0;
    }
    const int g = mg_grid(L.n);
    const int gb = mg_grid((int64_t)L.n_blk * 8);
    const bool dist0 = (k == 0 && mg->dist != nullptr);          // partitioned level: ghost coefficients behind the owned ones
    const int32_t* bc = (k == 0 && mg->has_p1 && !getenv("REYNA_B200_MGB_SCALAR_INDICES")) ? mg->bcol : nullptr;
    const int ncols = dist0 ? L.n + 3 * mg->n_ghost_elem : L.n;
    int rc = 0;
    // the iterate ping-pongs between L.xt and L.x; 2 nu - 1 sweeps in total, so starting in L.xt it ends in L.x
    double *cur = L.xt, *oth = L.x;
    mg_block_apply<false><<<g, kMgThreads, 0, st>>>(L.n, 3, L.minv, L.b, cur);
    auto sweep = [&]() {
        if (dist0 && !rc) rc = dist_halo_d(mg->dist, cur, 3, st);
        if (L.dataf) mgb_block_row<true, float><<<gb, kMgThreads, 0, st>>>(L.n_blk, ncols, L.indptr, L.indices, L.dataf, L.minv, L.b, cur, oth, bc);
        else mgb_block_row<true, double><<<gb, kMgThreads, 0, st>>>(L.n_blk, ncols, L.indptr, L.indices, L.data, L.minv, L.b, cur, oth, bc);
        double* t = cur; cur = oth; oth = t;
    };
    for (int s = 1; s < mg->nu; ++s) sweep();
    RB_LAUNCH_CHECK("mgb pre-smoothing");
    MgLevel& C = mg->blev[k + 1];
    if (dist0 && !rc) rc = dist_halo_d(mg->dist, cur, 3, st);
    if (rc) return rc;
    if (L.dataf) mgb_block_row<false, float><<<gb, kMgThreads, 0, st>>>(L.n_blk, ncols, L.indptr, L.indices, L.dataf, L.minv, L.b, cur, oth, bc);
    else mgb_block_row<false, double><<<gb, kMgThreads, 0, st>>>(L.n_blk, ncols, L.indptr, L.indices, L.data, L.minv, L.b, cur, oth, bc);
    if (dist0) {
        // this rank's aggregates -> slice of the global level-1 right-hand side, all-gathered; the chain below is replicated
        const int nagg_loc = (L.n_blk + 3) / 4;
        mgb_restrict<<<mg_grid(nagg_loc), kMgThreads, 0, st>>>(nagg_loc, L.n_blk, L.tr, oth, mg->p1_t0);
        RB_LAUNCH_CHECK("mgb_restrict");
        if ((rc = dist_allgather_rows(mg->dist, mg->p1_t0, C.b, mg->bounds_l1.data(), st))) return rc;
    } else {
        mgb_restrict<<<mg_grid(C.n_blk), kMgThreads, 0, st>>>(C.n_blk, L.n_blk, L.tr, oth, C.b);
        RB_LAUNCH_CHECK("mgb_restrict");
    }
    if ((rc = mgb_vcycle(mg, k + 1, st))) return rc;
    mgb_prolong_add<<<mg_grid(L.n_blk), kMgThreads, 0, st>>>(L.n_blk, mg->alpha, L.tr, C.x + (dist0 ? 3 * (size_t)(mg->lo_elem / 4) : 0),
                                                         cur);
    for (int s = 0; s < mg->nu; ++s) sweep();
    RB_LAUNCH_CHECK("mgb post-smoothing");
    if (cur != L.x) RB_CUDA(cudaMemcpyAsync(L.x, cur, sizeof(double) * (size_t)L.n, cudaMemcpyDeviceToDevice, st));   // nu == 0 only
    return rc;
}


// out[i] = in[offset + i]
__global__ void mg_copy_slice(int n, const double* __restrict__ in, int offset, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[offset + i];
}

// Distributed mode: the element-wise P0 right-hand side is all-gathered, every rank runs the (small, global) P0 chain
// redundantly and keeps its slice -- the coarse space stays GLOBAL, so iteration counts do not depend on the number of GPUs.
// z += (coarse-space correction of r); r, z fine-level vectors (owned part)
int mg_apply_add(rb_mg* mg, const double* r, double* z, cudaStream_t st) {
    const int E = mg->n_elem, d = mg->d;
    if (mg->agg_p1) {
        int rc;
        MgLevel& B0 = mg->blev[0];
        if (mg->has_p1) {
            // z += R1^T V(R1 r): V-cycle over the agglomerated-P1 chain starting on the P1 level
            const int s0 = mg->sel1[0], s1 = mg->sel1[1], s2 = mg->sel1[2];
            mg_select<<<mg_grid(3 * E), kMgThreads, 0, st>>>(E, d, 3, s0, s1, s2, r, B0.b);
            RB_LAUNCH_CHECK("mg_select");
            if ((rc = mgb_vcycle(mg, 0, st))) return rc;
            mg_pad_add<<<mg_grid(3 * E), kMgThreads, 0, st>>>(E, d, 3, s0, s1, s2, B0.x, z);
            RB_LAUNCH_CHECK("mg_pad_add");
        } else {
            // p = 1: the fine level is the P1 level and block Jacobi on it is the (additive) smoother of the caller;
            // z += alpha P V(P^T r) with the chain starting on the first agglomerated level
            if (mg->nblev == 1) {          // tiny mesh: the fine level itself is solved densely
                mg_dense_apply<<<(B0.n + 31) / 32, 256, 0, st>>>(B0.n, mg->dense_inv, r, B0.x);
                mg_pad_add<<<mg_grid(B0.n), kMgThreads, 0, st>>>(E, 3, 3, 0, 1, 2, B0.x, z);
                RB_LAUNCH_CHECK("mg dense fine level");
                return 0;
            }
            MgLevel& B1 = mg->blev[1];
            if (mg->dist) {
                const int nagg_loc = (B0.n_blk + 3) / 4;
                mgb_restrict<<<mg_grid(nagg_loc), kMgThreads, 0, st>>>(nagg_loc, B0.n_blk, B0.tr, r, mg->p1_t0);
                RB_LAUNCH_CHECK("mgb_restrict");
                if ((rc = dist_allgather_rows(mg->dist, mg->p1_t0, B1.b, mg->bounds_l1.data(), st))) return rc;
            } else {
                mgb_restrict<<<mg_grid(B1.n_blk), kMgThreads, 0, st>>>(B1.n_blk, B0.n_blk, B0.tr, r, B1.b);
                RB_LAUNCH_CHECK("mgb_restrict");
            }
            if ((rc = mgb_vcycle(mg, 1, st))) return rc;
            mgb_prolong_add<<<mg_grid(B0.n_blk), kMgThreads, 0, st>>>(B0.n_blk, mg->alpha, B0.tr,
                                                                  B1.x + (mg->dist ? 3 * (size_t)(mg->lo_elem / 4) : 0), z);
            RB_LAUNCH_CHECK("mgb_prolong_add");
        }
        return 0;
    }
    MgLevel& L0 = mg->lev[0];
    const bool dist = mg->dist != nullptr;
    // in distributed mode the local P0 right-hand side / correction live in the ping-pong vector of level 0's tail
    double* b0_local = dist ? mg->p1_t0 : L0.b;
    int rc;
    if (mg->has_p1) {
        const int n1 = 3 * E, g1 = mg_grid(n1);
        const int nc1 = 3 * (E + mg->n_ghost_elem);
        const int s0 = mg->sel1[0], s1 = mg->sel1[1], s2 = mg->sel1[2];
        mg_select<<<g1, kMgThreads, 0, st>>>(E, d, 3, s0, s1, s2, r, mg->p1_r);
        mg_block_apply<false><<<g1, kMgThreads, 0, st>>>(n1, 3, mg->p1_minv, mg->p1_r, mg->p1_x);          // pre-smoothing
        RB_LAUNCH_CHECK("mg p1 pre");
        if (dist && (rc = dist_halo_d(mg->dist, mg->p1_x, 3, st))) return rc;
        mg_residual_blocked3<8><<<mg_grid((int64_t)n1 * 8), kMgThreads, 0, st>>>(n1, nc1, mg->p1_indptr, mg->bcol, mg->p1_data,
                                                                                mg->p1_r, mg->p1_x, mg->p1_t);
        mg_select<<<mg_grid(E), kMgThreads, 0, st>>>(E, 3, 1, 0, 0, 0, mg->p1_t, b0_local);
        RB_LAUNCH_CHECK("mg p1 down");
    } else {
        mg_select<<<mg_grid(E), kMgThreads, 0, st>>>(E, d, 1, 0, 0, 0, r, b0_local);
        RB_LAUNCH_CHECK("mg_select");
    }
    if (dist && (rc = dist_allgather_rows(mg->dist, b0_local, L0.b, mg->bounds.data(), st))) return rc;
    if ((rc = mg_vcycle(mg, 0, st))) return rc;
    const double* x0_local = L0.x + (dist ? mg->lo_elem : 0);
    if (mg->has_p1) {
        const int n1 = 3 * E, g1 = mg_grid(n1);
        const int nc1 = 3 * (E + mg->n_ghost_elem);
        const int s0 = mg->sel1[0], s1 = mg->sel1[1], s2 = mg->sel1[2];
        mg_pad_add<<<mg_grid(E), kMgThreads, 0, st>>>(E, 3, 1, 0, 0, 0, x0_local, mg->p1_x);
        RB_LAUNCH_CHECK("mg_pad_add");
        if (dist && (rc = dist_halo_d(mg->dist, mg->p1_x, 3, st))) return rc;
        mg_residual_blocked3<8><<<mg_grid((int64_t)n1 * 8), kMgThreads, 0, st>>>(n1, nc1, mg->p1_indptr, mg->bcol, mg->p1_data,
                                                                                mg->p1_r, mg->p1_x, mg->p1_t);
        mg_block_apply<true><<<g1, kMgThreads, 0, st>>>(n1, 3, mg->p1_minv, mg->p1_t, mg->p1_x);           // post-smoothing
        mg_pad_add<<<g1, kMgThreads, 0, st>>>(E, d, 3, s0, s1, s2, mg->p1_x, z);
        RB_LAUNCH_CHECK("mg p1 up");
    } else {
        mg_pad_add<<<mg_grid(E), kMgThreads, 0, st>>>(E, d, 1, 0, 0, 0, x0_local, z);
        RB_LAUNCH_CHECK("mg_pad_add");
    }
    return 0;
}

}  // namespace rb

using namespace rb;

extern "C" void reyna_b200_mg_destroy(rb_mg* mg) {
    if (!mg) return;
    for (void* p : mg->owned) cudaFree(p);
    delete mg;
}

extern "C" int reyna_b200_mg_create(int32_t n, int32_t d, int32_t p, const int32_t* indptr, const int32_t* indices,
                                    const double* data, const rb_mg_opts* opts, const rb_mg_dist* dd, rb_mg** out,
                                    void* stream) {
    RB_REQUIRE(indptr && indices && data && out, RB_E_ARG, "reyna_b200_mg_create: null argument");
    RB_REQUIRE(d >= 1 && n % d == 0 && d == (p + 1) * (p + 2) / 2, RB_E_ARG, "reyna_b200_mg_create: bad block size");
    cudaStream_t st = (cudaStream_t)stream;
    rb_mg* mg = new rb_mg();
    mg->n_f = n; mg->d = d; mg->p = p; mg->n_elem = n / d;
    mg->f_indptr = indptr; mg->f_indices = indices; mg->f_data = data;
    if (opts) {
        if (opts->nu > 0) mg->nu = opts->nu;
        if (opts->omega > 0) mg->omega = opts->omega;
        if (opts->alpha > 0) mg->alpha = opts->alpha;
    }
    const int E = mg->n_elem;
    int rc = 0;
    auto fail = [&](int code) { reyna_b200_mg_destroy(mg); return code; };
    const bool dist_agg = dd && dd->l1_indptr && dd->l1_indices && dd->l1_data && dd->l1_bbox && dd->l0_tr;
    if (dd) {
        if (!dd->dist || !dd->bounds || (!dist_agg && (!dd->p0_indptr || !dd->p0_indices || !dd->p0_data))) {
            set_error("reyna_b200_mg_create: incomplete rb_mg_dist");
            return fail(RB_E_ARG);
        }
        mg->dist = dd->dist;
        mg->n_ghost_elem = dd->n_ghost_elem;
        mg->bounds.assign(dd->bounds, dd->bounds + dd->world + 1);
        mg->lo_elem = dd->bounds[dd->rank];
    }
    int32_t h_last = 0;
    RB_CUDA(cudaMemcpyAsync(&h_last, indptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz_f = h_last;
    const int64_t nblocks = nnz_f / ((int64_t)d * d);
    {   // arena chunks of about the size of the P1 level (the largest thing the hierarchy owns): 8 MB ... 2 GB
        size_t chunk = (size_t)nnz_f * 4;
        if (chunk < (8u << 20)) chunk = 8u << 20;
        if (chunk > ((size_t)2 << 30)) chunk = (size_t)2 << 30;
        mg->arena_chunk = chunk;
    }

    // ---- P1 level --------------------------------------------------------------------------------------------------
    MgLevel& L0 = mg->lev[0];
    if (dd) {
        // the GLOBAL piecewise-constant matrix (all ranks' rows, gathered by the caller), replicated on every rank
        L0.n = dd->n_global_elem;
        L0.nnz = dd->p0_nnz;
        L0.indptr = const_cast<int32_t*>(dd->p0_indptr);
        L0.indices = const_cast<int32_t*>(dd->p0_indices);
        L0.data = const_cast<double*>(dd->p0_data);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_t0, sizeof(double) * (size_t)E))) return fail(rc);
    } else {
        L0.n = E;
        L0.nnz = nblocks;
        if ((rc = dev_alloc(mg, (void**)&L0.indptr, sizeof(int32_t) * ((size_t)E + 1)))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&L0.indices, sizeof(int32_t) * (size_t)nblocks))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&L0.data, sizeof(double) * (size_t)nblocks))) return fail(rc);
    }
    if (d > 3) {
        mg->has_p1 = 1;
        mg->sel1[0] = 0; mg->sel1[1] = 1; mg->sel1[2] = p + 1;     // modes (0,0), (0,1), (1,0) in Basis_index2D order
        const int n1 = 3 * E;
        if ((rc = dev_alloc(mg, (void**)&mg->p1_indptr, sizeof(int32_t) * ((size_t)n1 + 1)))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_indices, sizeof(int32_t) * (size_t)nblocks * 9))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_data, sizeof(double) * (size_t)nblocks * 9))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_minv, sizeof(double) * (size_t)n1 * 3))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_r, sizeof(double) * (size_t)n1))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->p1_x, sizeof(double) * ((size_t)n1 + 3 * (size_t)mg->n_ghost_elem)))) return fail(rc);
        RB_CUDA(cudaMemsetAsync(mg->p1_x, 0, sizeof(double) * ((size_t)n1 + 3 * (size_t)mg->n_ghost_elem), st));
        if ((rc = dev_alloc(mg, (void**)&mg->p1_t, sizeof(double) * ((size_t)n1 + 3 * (size_t)mg->n_ghost_elem)))) return fail(rc);
        RB_CUDA(cudaMemsetAsync(mg->p1_t, 0, sizeof(double) * ((size_t)n1 + 3 * (size_t)mg->n_ghost_elem), st));
        mg_extract_modes<<<mg_grid(n1), kMgThreads, 0, st>>>(E, d, 3, mg->sel1[0], mg->sel1[1], mg->sel1[2], indptr, indices,
                                                            data, mg->p1_indptr, mg->p1_indices, mg->p1_data);
        if ((rc = dev_alloc(mg, (void**)&mg->bcol, sizeof(int32_t) * (size_t)(nblocks + 1)))) return fail(rc);
        if ((rc = launch_block_columns(E, d, indptr, indices, mg->bcol, st))) return fail(rc);
        int* sing = nullptr;
        if ((rc = dev_alloc(mg, (void**)&sing, sizeof(int)))) return fail(rc);
        cudaMemsetAsync(sing, 0, sizeof(int), st);
        block_jacobi_setup_launch(E, 3, mg->p1_indptr, mg->p1_indices, mg->p1_data, mg->p1_minv, sing, st);
        if (!dd)
            mg_extract_modes<<<mg_grid(E), kMgThreads, 0, st>>>(E, 3, 1, 0, 0, 0, mg->p1_indptr, mg->p1_indices, mg->p1_data,
                                                               L0.indptr, L0.indices, L0.data);
    } else if (!dd) {
        mg_extract_modes<<<mg_grid(E), kMgThreads, 0, st>>>(E, d, 1, 0, 0, 0, indptr, indices, data, L0.indptr, L0.indices,
                                                           L0.data);
    }
    if (cudaGetLastError() != cudaSuccess) { set_error("reyna_b200_mg_create: mode extraction failed"); return fail(RB_E_ARG); }

    // ---- agglomerated-P1 chain (single-GPU hierarchy with element bounding boxes) ----------------------------------------
    if (dist_agg) {
        for (int r = 0; r <= dd->world; ++r) {
            if (r < dd->world && dd->bounds[r] % 4 != 0) {
                set_error("reyna_b200_mg_create: element ranges of the agglomerated hierarchy must start at multiples of 4");
                return fail(RB_E_ARG);
            }
            mg->bounds_l1.push_back(3 * ((dd->bounds[r] + 3) / 4));
        }
    }
    if (d >= 3 && (dist_agg || (!dd && opts && opts->elem_bbox))) {
        mg->agg_p1 = 1;
        MgLevel& B0 = mg->blev[0];
        B0.n_blk = E; B0.n = 3 * E;
        if (mg->has_p1) {
            B0.indptr = mg->p1_indptr; B0.indices = mg->p1_indices; B0.data = mg->p1_data;
            B0.minv = mg->p1_minv; B0.b = mg->p1_r; B0.x = mg->p1_x; B0.xt = mg->p1_t;
            mgb_scale<<<mg_grid((int64_t)9 * E), kMgThreads, 0, st>>>((int64_t)9 * E, mg->omega, B0.minv);
        } else {
            B0.indptr = const_cast<int32_t*>(indptr); B0.indices = const_cast<int32_t*>(indices); B0.data = const_cast<double*>(data);
            if ((rc = dev_alloc(mg, (void**)&B0.x, sizeof(double) * (size_t)B0.n))) return fail(rc);
        }
        B0.nnz = nblocks * 9;
        B0.bbox = dist_agg ? nullptr : const_cast<double*>(opts->elem_bbox);
        int32_t* scratch = nullptr;
        const int64_t n_scan = dist_agg ? (int64_t)dd->n_global_elem + 2 : (int64_t)E + 2;
        if ((rc = dev_alloc(mg, (void**)&scratch, sizeof(int32_t) * (scan_scratch_ints(n_scan) + 4)))) return fail(rc);
        int* sing = nullptr;
        if ((rc = dev_alloc(mg, (void**)&sing, sizeof(int)))) return fail(rc);
        cudaMemsetAsync(sing, 0, sizeof(int), st);
        int k = 0;
        if (dist_agg) {
            // level 1 is the caller's replicated GLOBAL matrix; the transfer of the owned elements comes with it
            MgLevel& C = mg->blev[1];
            C.n_blk = (dd->n_global_elem + 3) / 4; C.n = 3 * C.n_blk;
            C.nnz = dd->l1_nnz;
            C.indptr = const_cast<int32_t*>(dd->l1_indptr); C.indices = const_cast<int32_t*>(dd->l1_indices);
            C.data = const_cast<double*>(dd->l1_data); C.bbox = const_cast<double*>(dd->l1_bbox);
            B0.tr = const_cast<double*>(dd->l0_tr);
            if ((rc = dev_alloc(mg, (void**)&C.minv, sizeof(double) * 9 * (size_t)C.n_blk))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.b, sizeof(double) * (size_t)C.n))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.x, sizeof(double) * (size_t)C.n))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.xt, sizeof(double) * (size_t)C.n))) return fail(rc);
            block_jacobi_setup_launch(C.n_blk, 3, C.indptr, C.indices, C.data, C.minv, sing, st);
            mgb_scale<<<mg_grid((int64_t)9 * C.n_blk), kMgThreads, 0, st>>>((int64_t)9 * C.n_blk, mg->omega, C.minv);
            if ((rc = dev_alloc(mg, (void**)&mg->p1_t0, sizeof(double) * 3 * ((size_t)(E + 3) / 4 + 1)))) return fail(rc);
            k = 1;
        }
        while (mg->blev[k].n > kDenseMax && k < kMaxLevels - 1) {
            MgLevel& L = mg->blev[k];
            MgLevel& C = mg->blev[k + 1];
            C.n_blk = (L.n_blk + 3) / 4; C.n = 3 * C.n_blk;
            if ((rc = dev_alloc(mg, (void**)&C.bbox, sizeof(double) * 4 * (size_t)C.n_blk))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&L.tr, sizeof(double) * 8 * (size_t)L.n_blk))) return fail(rc);
            mgb_coarsen_bbox<<<mg_grid(C.n_blk), kMgThreads, 0, st>>>(C.n_blk, L.n_blk, L.bbox, C.bbox);
            mgb_transfer<<<mg_grid(L.n_blk), kMgThreads, 0, st>>>(L.n_blk, L.bbox, C.bbox, L.tr);
            int32_t* boff = nullptr;
            if ((rc = dev_alloc(mg, (void**)&boff, sizeof(int32_t) * ((size_t)C.n_blk + 1)))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.indptr, sizeof(int32_t) * ((size_t)C.n + 1)))) return fail(rc);
            mgb_aggregate<false><<<mg_grid(C.n_blk), kMgThreads, 0, st>>>(C.n_blk, L.n_blk, L.indptr, L.indices, L.data, L.tr, boff,
                                                                      nullptr, nullptr, nullptr);
            RB_CUDA(cudaMemsetAsync(boff + C.n_blk, 0, sizeof(int32_t), st));
            if ((rc = exclusive_scan_i32(boff, boff, (int64_t)C.n_blk + 1, scratch, nullptr, st))) return fail(rc);
            int32_t tot = 0;
            RB_CUDA(cudaMemcpyAsync(&tot, boff + C.n_blk, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            RB_CUDA(cudaStreamSynchronize(st));
            C.nnz = (int64_t)tot * 9;
            if ((rc = dev_alloc(mg, (void**)&C.indices, sizeof(int32_t) * (size_t)C.nnz))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.data, sizeof(double) * (size_t)C.nnz))) return fail(rc);
            mgb_aggregate<true><<<mg_grid(C.n_blk), kMgThreads, 0, st>>>(C.n_blk, L.n_blk, L.indptr, L.indices, L.data, L.tr, boff,
                                                                     C.indptr, C.indices, C.data);
            if ((rc = dev_alloc(mg, (void**)&C.minv, sizeof(double) * 9 * (size_t)C.n_blk))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.b, sizeof(double) * (size_t)C.n))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.x, sizeof(double) * (size_t)C.n))) return fail(rc);
            if ((rc = dev_alloc(mg, (void**)&C.xt, sizeof(double) * (size_t)C.n))) return fail(rc);
            block_jacobi_setup_launch(C.n_blk, 3, C.indptr, C.indices, C.data, C.minv, sing, st);
            mgb_scale<<<mg_grid((int64_t)9 * C.n_blk), kMgThreads, 0, st>>>((int64_t)9 * C.n_blk, mg->omega, C.minv);
            if (cudaGetLastError() != cudaSuccess) { set_error("reyna_b200_mg_create: agglomeration failed"); return fail(RB_E_ARG); }
            ++k;
        }
        mg->nblev = k + 1;
        mg->nlev = mg->nblev;
        if (opts && opts->fp32_levels) {
            // the smoothing sweeps are bound by the matrix values: a single-precision copy halves their traffic (the hierarchy is
            // a preconditioner; the Krylov iteration and every residual stay in double precision)
            for (int q = 0; q + 1 < mg->nblev; ++q) {
                MgLevel& Lq = mg->blev[q];
                if ((rc = dev_alloc(mg, (void**)&Lq.dataf, sizeof(float) * (size_t)Lq.nnz))) return fail(rc);
                mgb_to_float<<<mg_grid(Lq.nnz), kMgThreads, 0, st>>>(Lq.nnz, Lq.data, Lq.dataf);
            }
            RB_LAUNCH_CHECK("mgb_to_float");
        }
        for (int q = 0; q < mg->nblev; ++q) { mg->lev[q].n = mg->blev[q].n; mg->lev[q].nnz = mg->blev[q].nnz; }   // reyna_b200_mg_info
        MgLevel& Lc = mg->blev[mg->nblev - 1];
        double* scratch_a = nullptr;
        if ((rc = dev_alloc(mg, (void**)&scratch_a, sizeof(double) * (size_t)Lc.n * Lc.n))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->dense_inv, sizeof(double) * (size_t)Lc.n * Lc.n))) return fail(rc);
        mg_dense_inverse<<<1, kInvThreads, 0, st>>>(Lc.n, Lc.indptr, Lc.indices, Lc.data, scratch_a, mg->dense_inv);
        if (cudaGetLastError() != cudaSuccess) { set_error("reyna_b200_mg_create: dense inverse failed"); return fail(RB_E_ARG); }
        RB_CUDA(cudaStreamSynchronize(st));
        *out = mg;
        return 0;
    }

    // ---- aggregation chain -------------------------------------------------------------------------------------------
    auto alloc_vectors = [&](MgLevel& L) -> int {
        int r;
        if ((r = dev_alloc(mg, (void**)&L.dinv, sizeof(double) * (size_t)L.n))) return r;
        if ((r = dev_alloc(mg, (void**)&L.b, sizeof(double) * (size_t)L.n))) return r;
        if ((r = dev_alloc(mg, (void**)&L.x, sizeof(double) * (size_t)L.n))) return r;
        if ((r = dev_alloc(mg, (void**)&L.xt, sizeof(double) * (size_t)L.n))) return r;
        return 0;
    };
    if ((rc = alloc_vectors(L0))) return fail(rc);
    mg_diag_inv<<<mg_grid(L0.n), kMgThreads, 0, st>>>(L0.n, L0.indptr, L0.indices, L0.data, L0.dinv);
    int32_t* scratch = nullptr;
    if ((rc = dev_alloc(mg, (void**)&scratch, sizeof(int32_t) * (scan_scratch_ints((int64_t)L0.n + 1) + 4)))) return fail(rc);
    int k = 0;
    while (mg->lev[k].n > kDenseMax && k < kMaxLevels - 1) {
        MgLevel& L = mg->lev[k];
        MgLevel& C = mg->lev[k + 1];
        C.n = (L.n + 3) / 4;
        if ((rc = dev_alloc(mg, (void**)&C.indptr, sizeof(int32_t) * ((size_t)C.n + 1)))) return fail(rc);
        if ((rc = alloc_vectors(C))) return fail(rc);
        // pass 1: entries per coarse row -> exclusive scan in place (entry C.n becomes the total)
        mg_aggregate<false><<<mg_grid(C.n), kMgThreads, 0, st>>>(C.n, L.n, 4, L.indptr, L.indices, L.data, C.indptr, nullptr,
                                                                 nullptr, nullptr);
        RB_CUDA(cudaMemsetAsync(C.indptr + C.n, 0, sizeof(int32_t), st));
        if ((rc = exclusive_scan_i32(C.indptr, C.indptr, (int64_t)C.n + 1, scratch, nullptr, st))) return fail(rc);
        int32_t tot = 0;
        RB_CUDA(cudaMemcpyAsync(&tot, C.indptr + C.n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        RB_CUDA(cudaStreamSynchronize(st));
        C.nnz = tot;
        if ((rc = dev_alloc(mg, (void**)&C.indices, sizeof(int32_t) * (size_t)tot))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&C.data, sizeof(double) * (size_t)tot))) return fail(rc);
        // pass 2: merged rows, diagonal inverse
        mg_aggregate<true><<<mg_grid(C.n), kMgThreads, 0, st>>>(C.n, L.n, 4, L.indptr, L.indices, L.data, C.indptr, C.indices,
                                                                C.data, C.dinv);
        if (cudaGetLastError() != cudaSuccess) { set_error("reyna_b200_mg_create: aggregation failed"); return fail(RB_E_ARG); }
        ++k;
    }
    mg->nlev = k + 1;
    {
        MgLevel& Lc = mg->lev[mg->nlev - 1];
        double* scratch_a = nullptr;
        if ((rc = dev_alloc(mg, (void**)&scratch_a, sizeof(double) * (size_t)Lc.n * Lc.n))) return fail(rc);
        if ((rc = dev_alloc(mg, (void**)&mg->dense_inv, sizeof(double) * (size_t)Lc.n * Lc.n))) return fail(rc);
        mg_dense_inverse<<<1, kInvThreads, 0, st>>>(Lc.n, Lc.indptr, Lc.indices, Lc.data, scratch_a, mg->dense_inv);
        if (cudaGetLastError() != cudaSuccess) { set_error("reyna_b200_mg_create: dense inverse failed"); return fail(RB_E_ARG); }
    }
    RB_CUDA(cudaStreamSynchronize(st));
    *out = mg;
    return 0;
}

extern "C" int reyna_b200_mg_info(const rb_mg* mg, int32_t* n_levels, int32_t* level_rows, long long* level_nnz) {
    RB_REQUIRE(mg && n_levels, RB_E_ARG, "reyna_b200_mg_info: null argument");
    *n_levels = mg->nlev;
    for (int k = 0; k < mg->nlev; ++k) {
        if (level_rows) level_rows[k] = mg->lev[k].n;
        if (level_nnz) level_nnz[k] = mg->lev[k].nnz;
    }
    return 0;
}

// z = (coarse correction of r) -- exposed for tests
extern "C" int reyna_b200_mg_apply(rb_mg* mg, const double* r, double* z, void* stream) {
    RB_REQUIRE(mg && r && z, RB_E_ARG, "reyna_b200_mg_apply: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    RB_CUDA(cudaMemsetAsync(z, 0, sizeof(double) * (size_t)mg->n_f, st));
    return mg_apply_add(mg, r, z, st);
}
