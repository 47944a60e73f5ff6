// Deterministic block structure of the DG stiffness matrix, built on the device with integer-only kernels:
// count -> scan -> fill -> per-element (segmented) sort of the face items by (neighbour, face).
// This replaces scipy's COO->CSR conversion (coo_tocsr + sum_duplicates) that the reference runs on 4 d^2 F + d^2 E
// triplets (reyna/DGFEM/two_dimensional/main.py:319-320, 371-372): the pattern of B is determined by the
// element-face adjacency alone, so only F face items are sorted, never the d^2-times larger triplet list.
#include "common.cuh"
#include <stdarg.h>

namespace rb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

// ----------------------------------------------------------------------------------------------------------------
// exclusive scan (three launches; integer, deterministic)
// ----------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

size_t scan_scratch_ints(int64_t n) { return (size_t)div_up(n, kScanTile) + 1; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, total in *total (all threads)
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int s_w[kScanThreads / 32];
    __shared__ int s_tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < kScanThreads / 32 ? s_w[lane] : 0;
        int xi = warp_incl_scan(x);
        if (lane < kScanThreads / 32) s_w[lane] = xi - x;
        if (lane == kScanThreads / 32 - 1) s_tot = xi;
    }
    __syncthreads();
    int r = inc - v + s_w[w];
    *total = s_tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                                                               int32_t* __restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    int tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets(int32_t* __restrict__ tile_sums, int64_t n_tiles,
                                                                  int32_t* __restrict__ total_dev) {
    int carry = 0;
    for (int64_t base = 0; base < n_tiles; base += kScanThreads) {
        int64_t i = base + threadIdx.x;
        int v = i < n_tiles ? tile_sums[i] : 0;
        int tot;
        int ex = block_excl_scan(v, &tot);
        if (i < n_tiles) tile_sums[i] = ex + carry;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        tile_sums[n_tiles] = carry;
        if (total_dev) *total_dev = carry;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                           int64_t n, const int32_t* __restrict__ tile_offsets) {
    // thread owns kScanItems consecutive values so that the result is a plain exclusive scan
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int tot;
    int ex = block_excl_scan(s, &tot) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, int32_t* total_dev,
                       cudaStream_t st) {
    if (n <= 0) {
        if (total_dev) RB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int32_t), st));
        return 0;
    }
    const int64_t tiles = div_up(n, kScanTile);
    scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, n, scratch);
    RB_LAUNCH_CHECK("scan_tile_sums");
    scan_tile_offsets<<<1, kScanThreads, 0, st>>>(scratch, tiles, total_dev);
    RB_LAUNCH_CHECK("scan_tile_offsets");
    scan_apply<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, scratch);
    RB_LAUNCH_CHECK("scan_apply");
    return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// adjacency
// ----------------------------------------------------------------------------------------------------------------
constexpr int32_t kDupBit = (int32_t)0x80000000;

// which sides of interior face f contribute rows: bit0 -> K1 rows, bit1 -> K2 rows
__device__ __forceinline__ int face_mask(int has_diff, const uint8_t* __restrict__ upw, int f) {
    if (has_diff) return 3;
    return upw[f] ? 2 : 1;    // K1 upwind -> only the rows of K2 are written (full_assembly.py:209-210)
}

__global__ void adj_count(int n_iface, int n_rows, const int32_t* __restrict__ ielem, int has_diff,
                          const uint8_t* __restrict__ upw, int32_t* __restrict__ cnt) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n_iface; f += gridDim.x * blockDim.x) {
        const int m = face_mask(has_diff, upw, f);
        const int e0 = ielem[2 * f], e1 = ielem[2 * f + 1];
        if ((m & 1) && e0 < n_rows) atomicAdd(&cnt[e0], 1);
        if ((m & 2) && e1 < n_rows) atomicAdd(&cnt[e1], 1);
    }
}

__global__ void adj_fill(int n_iface, int n_rows, const int32_t* __restrict__ ielem, int has_diff,
                         const uint8_t* __restrict__ upw, const int32_t* __restrict__ adj_off,
                         int32_t* __restrict__ cursor, int32_t* __restrict__ adj_item) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n_iface; f += gridDim.x * blockDim.x) {
        const int m = face_mask(has_diff, upw, f);
        const int e0 = ielem[2 * f], e1 = ielem[2 * f + 1];
        if ((m & 1) && e0 < n_rows) adj_item[adj_off[e0] + atomicAdd(&cursor[e0], 1)] = 2 * f;
        if ((m & 2) && e1 < n_rows) adj_item[adj_off[e1] + atomicAdd(&cursor[e1], 1)] = 2 * f + 1;
    }
}

// The segmented sort: one thread per element orders its (<= ~20) items by (neighbour, face); the atomic cursor order of
// adj_fill is erased, so the layout is reproducible run to run.  Items that repeat the previous neighbour are flagged.
__device__ __forceinline__ int gid_of(const int32_t* __restrict__ gid, int e) { return gid ? gid[e] : e; }

__global__ void adj_sort(int n_rows, const int32_t* __restrict__ ielem, const int32_t* __restrict__ gid,
                         const int32_t* __restrict__ adj_off, int32_t* __restrict__ adj_item,
                         int32_t* __restrict__ adj_nb, int32_t* __restrict__ grp_cnt, int32_t* __restrict__ counts) {
    int dups = 0, maxlen = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_rows; e += gridDim.x * blockDim.x) {
        const int lo = adj_off[e], hi = adj_off[e + 1];
        const int len = hi - lo;
        maxlen = max(maxlen, len);
        constexpr int kSortMax = 12;       // items sorted from a local key array (polygons have ~6 edges)
        if (len <= kSortMax) {
            // one gather per item for the neighbour id, then an insertion sort on the packed key (neighbour gid, item)
            unsigned long long key[kSortMax];
#pragma unroll
            for (int a = 0; a < kSortMax; ++a) {
                key[a] = ~0ull;
                if (a < len) {
                    const int it = adj_item[lo + a];
                    const int nb = gid_of(gid, ielem[2 * (it >> 1) + (1 - (it & 1))]);
                    key[a] = ((unsigned long long)(unsigned)nb << 32) | (unsigned)it;
                }
            }
            // odd-even transposition network: fixed compare/exchange pattern, every index a compile-time constant
#pragma unroll
            for (int round = 0; round < kSortMax; ++round) {
#pragma unroll
                for (int a = round & 1; a + 1 < kSortMax; a += 2) {
                    const unsigned long long x = key[a], y = key[a + 1];
                    key[a] = x < y ? x : y;
                    key[a + 1] = x < y ? y : x;
                }
            }
#pragma unroll
            for (int a = 0; a < kSortMax; ++a)
                if (a < len) adj_item[lo + a] = (int)(unsigned)key[a];
        } else {
        // insertion sort in place (global memory, L1-resident)
        for (int a = lo + 1; a < hi; ++a) {
            const int it = adj_item[a];
            const int f = it >> 1, side = it & 1;
            const int nb = gid_of(gid, ielem[2 * f + (1 - side)]);
            int b = a - 1;
            while (b >= lo) {
                const int jt = adj_item[b];
                const int jf = jt >> 1, js = jt & 1;
                const int jn = gid_of(gid, ielem[2 * jf + (1 - js)]);
                if (jn > nb || (jn == nb && jt > it)) {
                    adj_item[b + 1] = jt;
                    --b;
                } else {
                    break;
                }
            }
            adj_item[b + 1] = it;
        }
        }
        int groups = 0, prev = -1;
        for (int a = lo; a < hi; ++a) {
            const int it = adj_item[a];
            const int onb = ielem[2 * (it >> 1) + (1 - (it & 1))];
            if (adj_nb) adj_nb[a] = onb;            // neighbour (local id) of every item, for the streaming assembly kernel
            const int nb = gid_of(gid, onb);
            if (a > lo && nb == prev) {
                adj_item[a] = it | kDupBit;
                ++dups;
            } else {
                ++groups;
            }
            prev = nb;
        }
        grp_cnt[e] = groups;
    }
    if (dups) atomicAdd(&counts[2], dups);
    if (maxlen) atomicMax(&counts[3], maxlen);
}

}  // namespace rb

using namespace rb;

extern "C" size_t reyna_b200_structure_workspace(int32_t n_rows, int32_t n_iface) {
    (void)n_iface;
    // cnt/cursor [n_rows+1] + grp_cnt [n_rows+1] + scan scratch
    return sizeof(int32_t) * (2 * ((size_t)n_rows + 1) + scan_scratch_ints((int64_t)n_rows + 1) + 16);
}

static int structure_impl(const rb_geom* g, int has_diffusion, const uint8_t* if_upwind, rb_structure* s, void* workspace,
                          size_t workspace_bytes, bool read_counters, void* stream) {
    RB_REQUIRE(g && s && workspace, RB_E_ARG, "reyna_b200_structure: null argument");
    RB_REQUIRE(has_diffusion || if_upwind, RB_E_NOCOEF,
               "reyna_b200_structure: no diffusion and no upwind flags (advection missing)");
    RB_REQUIRE(workspace_bytes >= reyna_b200_structure_workspace(g->n_rows, g->n_iface), RB_E_WORKSPACE,
               "reyna_b200_structure: workspace too small");
    RB_REQUIRE(s->adj_off && s->adj_item && s->grp_off && s->counts, RB_E_ARG, "reyna_b200_structure: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = g->n_rows;
    int32_t* cnt = (int32_t*)workspace;
    int32_t* grp_cnt = cnt + (n + 1);
    int32_t* scratch = grp_cnt + (n + 1);
    RB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * 2 * ((size_t)n + 1), st));
    RB_CUDA(cudaMemsetAsync(s->counts, 0, sizeof(int32_t) * 4, st));
    const int threads = 256;
    const int fblocks = grid_for(g->n_iface, threads, kNumSMs * 16);
    const int eblocks = grid_for(n, threads, kNumSMs * 16);
    if (g->n_iface > 0) {
        adj_count<<<fblocks, threads, 0, st>>>(g->n_iface, n, g->ielem, has_diffusion, if_upwind, cnt);
        RB_LAUNCH_CHECK("adj_count");
    }
    int rc = exclusive_scan_i32(cnt, s->adj_off, (int64_t)n + 1, scratch, s->counts + 0, st);
    if (rc) return rc;
    RB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)n + 1), st));
    if (g->n_iface > 0) {
        adj_fill<<<fblocks, threads, 0, st>>>(g->n_iface, n, g->ielem, has_diffusion, if_upwind, s->adj_off, cnt,
                                              s->adj_item);
        RB_LAUNCH_CHECK("adj_fill");
    }
    if (n > 0) {
        adj_sort<<<eblocks, threads, 0, st>>>(n, g->ielem, g->elem_gid, s->adj_off, s->adj_item, s->adj_nb, grp_cnt, s->counts);
        RB_LAUNCH_CHECK("adj_sort");
    }
    rc = exclusive_scan_i32(grp_cnt, s->grp_off, (int64_t)n + 1, scratch, s->counts + 1, st);
    if (rc) return rc;
    if (!read_counters) return 0;
    int32_t h[4];
    RB_CUDA(cudaMemcpyAsync(h, s->counts, sizeof(h), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    s->n_items = h[0];
    s->n_groups = h[1];
    s->n_dup_items = h[2];
    s->max_items = h[3];
    return 0;
}

extern "C" int reyna_b200_structure(const rb_geom* g, int has_diffusion, const uint8_t* if_upwind, rb_structure* s,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    return structure_impl(g, has_diffusion, if_upwind, s, workspace, workspace_bytes, true, stream);
}

extern "C" int reyna_b200_structure_known(const rb_geom* g, int has_diffusion, const uint8_t* if_upwind, rb_structure* s,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(s && s->n_items >= 0 && s->n_groups >= 0, RB_E_ARG, "reyna_b200_structure_known: the host counters must be filled in");
    return structure_impl(g, has_diffusion, if_upwind, s, workspace, workspace_bytes, false, stream);
}

// ----------------------------------------------------------------------------------------------------------------
// CSR index arrays
// ----------------------------------------------------------------------------------------------------------------
namespace rb {

constexpr int kPatWarps = 8;
constexpr int kMaxNb = 64;   // neighbours per element held in shared memory (polygons have <= ~20 edges)
constexpr int kPatBlocksPerSM = 4;
constexpr int kFastNb = 16;  // block columns per element of the lane-parallel path (element itself included)

// One warp per 32 CONSECUTIVE block rows.  Phase 1 is lane-parallel: lane l reads the offsets and the (sorted) adjacency
// of element e0 + l and writes its ascending block-column list, diagonal inserted, to shared memory -- the scalar
// bookkeeping is done once per lane instead of once per warp (ncu of the warp-per-element version: 250 warp
// instructions per element, issue-bound at 0.34 ms for 1 M elements).  Phase 2 is warp-cooperative: every scalar row (e,j)
// of a block row has the same column list, which is streamed d times to the contiguous index range of the block row
// with 16-byte stores.  Elements with more than kFastNb block columns rebuild their list with the warp-wide ballot path.
template <bool LOCAL, int D>
__global__ void __launch_bounds__(kPatWarps * 32, kPatBlocksPerSM) csr_pattern_kernel(int n_rows, int d_runtime, const int32_t* __restrict__ ielem,
                                                                     const int32_t* __restrict__ adj_nb,
                                                                     const int32_t* __restrict__ gid,
                                                                     const int32_t* __restrict__ adj_off,
                                                                     const int32_t* __restrict__ adj_item,
                                                                     const int32_t* __restrict__ grp_off,
                                                                     int32_t* __restrict__ indptr,
                                                                     int32_t* __restrict__ indices) {
    __shared__ int s_nb[kPatWarps][32][kFastNb];
    __shared__ int s_big[kPatWarps][kMaxNb + 1];
    constexpr int d = D;               // compile-time: the index arithmetic below divides by it
    (void)d_runtime;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int dd = d * d;
    const int n_groups32 = (n_rows + 31) >> 5;
    for (int grp = blockIdx.x * kPatWarps + w; grp < n_groups32; grp += gridDim.x * kPatWarps) {
        // ---- phase 1: one element per lane ----------------------------------------------------------------------------
        const int e = (grp << 5) + lane;
        int nb = 0, rowlen = 0;
        long long base = 0;
        if (e < n_rows) {
            const int lo = adj_off[e], hi = adj_off[e + 1];
            const int g0 = grp_off[e];
            nb = grp_off[e + 1] - g0 + 1;
            rowlen = nb * d;
            base = (long long)dd * (g0 + e);
            const int eg = gid_of(gid, e);                   // block columns are ordered and numbered by GLOBAL element id
            if (nb <= kFastNb) {
                int cnt = 0;
                bool placed = false;
                // the first kPre items are loaded together (independent loads in flight), the rest one by one
                constexpr int kPre = 8;
                int pit[kPre], pol[kPre];
#pragma unroll
                for (int a = 0; a < kPre; ++a) {
                    pit[a] = -1;
                    pol[a] = -1;
                    if (adj_nb && lo + a < hi) {
                        pit[a] = adj_item[lo + a];
                        pol[a] = adj_nb[lo + a];
                    }
                }
#pragma unroll
                for (int a = 0; a < kPre; ++a) {
                    if (pit[a] < 0) continue;                // absent, or same neighbour as the previous item
                    const int o = gid_of(gid, pol[a]);
                    if (!placed && o > eg) {
                        s_nb[w][lane][cnt++] = LOCAL ? e : eg;
                        placed = true;
                    }
                    s_nb[w][lane][cnt++] = LOCAL ? pol[a] : o;
                }
                for (int a = adj_nb ? lo + kPre : lo; a < hi; ++a) {
                    const int it = adj_item[a];
                    if (it < 0) continue;                    // same neighbour as the previous item
                    const int ol = adj_nb ? adj_nb[a] : ielem[2 * (it >> 1) + (1 - (it & 1))];
                    const int o = gid_of(gid, ol);
                    if (!placed && o > eg) {
                        s_nb[w][lane][cnt++] = LOCAL ? e : eg;
                        placed = true;
                    }
                    s_nb[w][lane][cnt++] = LOCAL ? ol : o;
                }
                if (!placed) s_nb[w][lane][cnt] = LOCAL ? e : eg;
            }
#pragma unroll
            for (int j = 0; j < d; ++j) indptr[(int64_t)e * d + j] = (int32_t)(base + (long long)j * rowlen);
        }
        __syncwarp();
        // ---- phase 2: the warp writes the index block of each of its elements -------------------------------------------
        const int n_here = min(32, n_rows - (grp << 5));
        for (int i = 0; i < n_here; ++i) {
            const int nb_i = __shfl_sync(0xffffffffu, nb, i);
            const int rl = nb_i * d;
            const long long base_i = __shfl_sync(0xffffffffu, base, i);
            const int* list = s_nb[w][i];
            if (nb_i > kFastNb) {
                // rare: rebuild the list of element ei with one item per lane, compacted with a ballot
                const int ei = (grp << 5) + i;
                const int lo = adj_off[ei], hi = adj_off[ei + 1];
                const int eg = gid_of(gid, ei);
                int before = 0, below = 0;
                for (int a0 = lo; a0 < hi; a0 += 32) {
                    const int a = a0 + lane;
                    int it = -1, o = -1, ol = -1;
                    if (a < hi) {
                        it = adj_item[a];
                        if (it >= 0) {
                            ol = adj_nb ? adj_nb[a] : ielem[2 * (it >> 1) + (1 - (it & 1))];
                            o = gid_of(gid, ol);
                        }
                    }
                    const bool leader = it >= 0;
                    const unsigned m = __ballot_sync(0xffffffffu, leader);
                    const unsigned mlow = __ballot_sync(0xffffffffu, leader && o < eg);
                    if (leader) {
                        const int slot = before + __popc(m & ((1u << lane) - 1u)) + (o > eg ? 1 : 0);
                        if (slot <= kMaxNb) s_big[w][slot] = LOCAL ? ol : o;
                    }
                    before += __popc(m);
                    below += __popc(mlow);
                }
                if (lane == 0 && below <= kMaxNb) s_big[w][below] = LOCAL ? ei : eg;
                __syncwarp();
                list = s_big[w];
            }
            // every scalar row of the block row repeats the same rl columns: a lane computes its piece of the row once
            // and stores it d times, rl ints apart (fewer instructions than re-deriving the position of every store)
            if ((d & 1) == 0) {
                // 8-byte pieces: rows start at even offsets (d even) and a pair of columns never straddles two blocks
                const int half = rl >> 1;
                int2* out2 = reinterpret_cast<int2*>(indices + base_i);
                for (int c = lane; c < half; c += 32) {
                    const int col = c << 1;
                    const int slot = col / d;
                    const int v = list[slot] * d + (col - slot * d);
                    const int2 vv = make_int2(v, v + 1);
#pragma unroll
                    for (int j = 0; j < d; ++j) out2[j * half + c] = vv;
                }
            } else {
                int32_t* out = indices + base_i;
                for (int c = lane; c < rl; c += 32) {
                    const int slot = c / d;
                    const int v = list[slot] * d + (c - slot * d);
#pragma unroll
                    for (int j = 0; j < d; ++j) out[j * rl + c] = v;
                }
            }
            if (nb_i > kFastNb) __syncwarp();                 // s_big is reused by the next large element
        }
        __syncwarp();                                         // s_nb is rewritten by the next group
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        indptr[(int64_t)n_rows * d] = (int32_t)((int64_t)dd * (grp_off[n_rows] + n_rows));
}

// ----------------------------------------------------------------------------------------------------------------
// exact-zero pruning (scipy csr_plus_csr / csr_minus_csr semantics)
// ----------------------------------------------------------------------------------------------------------------
__global__ void compact_count_kernel(int n, const int32_t* __restrict__ indptr, const double* __restrict__ data,
                                     int32_t* __restrict__ row_nnz) {
    // one warp per row
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r <= n; r += gridDim.x * wpb) {
        int c = 0;
        if (r < n)
            for (int k = indptr[r] + lane; k < indptr[r + 1]; k += 32) c += (data[k] != 0.0);
#pragma unroll
        for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) row_nnz[r] = c;
    }
}

__global__ void compact_fill_kernel(int n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const double* __restrict__ data, const int32_t* __restrict__ out_indptr,
                                    int32_t* __restrict__ out_indices, double* __restrict__ out_data) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
        int o = out_indptr[r];
        for (int k0 = indptr[r]; k0 < indptr[r + 1]; k0 += 32) {
            const int k = k0 + lane;
            const bool live = k < indptr[r + 1];
            const double v = live ? data[k] : 0.0;
            const bool keep = live && v != 0.0;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = o + __popc(m & ((1u << lane) - 1u));
                out_data[pos] = v;
                out_indices[pos] = indices[k];
            }
            o += __popc(m);
        }
    }
}

}  // namespace rb

template <bool LOCAL>
static int launch_pattern(const rb_geom* g, const rb_structure* s, int p, int d, int blocks, int32_t* indptr, int32_t* indices,
                          cudaStream_t st) {
#define RB_PATTERN(DD)                                                                                                   \
    rb::csr_pattern_kernel<LOCAL, DD><<<blocks, rb::kPatWarps * 32, 0, st>>>(g->n_rows, d, g->ielem, s->adj_nb, g->elem_gid, \
                                                                             s->adj_off, s->adj_item, s->grp_off, indptr, indices)
    switch (p) {
        case 1: RB_PATTERN(3); break;
        case 2: RB_PATTERN(6); break;
        default: RB_PATTERN(10); break;
    }
#undef RB_PATTERN
    RB_LAUNCH_CHECK("csr_pattern_kernel");
    return 0;
}

extern "C" int reyna_b200_csr_pattern(const rb_geom* g, const rb_structure* s, int p, int32_t* indptr,
                                      int32_t* indices, void* stream) {
    RB_REQUIRE(g && s && indptr && indices, RB_E_ARG, "reyna_b200_csr_pattern: null argument");
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_csr_pattern: polynomial degree must be 1..3");
    RB_REQUIRE(s->max_items <= rb::kMaxNb, RB_E_ARG, "reyna_b200_csr_pattern: an element has more than 64 contributing interior faces");
    const int d = (p + 1) * (p + 2) / 2;
    const int blocks = grid_for((g->n_rows + 31) / 32, kPatWarps, kNumSMs * kPatBlocksPerSM);
    return launch_pattern<false>(g, s, p, d, blocks, indptr, indices, (cudaStream_t)stream);
}

extern "C" int reyna_b200_csr_pattern_local(const rb_geom* g, const rb_structure* s, int p, int32_t* indptr,
                                            int32_t* indices, void* stream) {
    RB_REQUIRE(g && s && indptr && indices, RB_E_ARG, "reyna_b200_csr_pattern_local: null argument");
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_csr_pattern_local: polynomial degree must be 1..3");
    RB_REQUIRE(s->max_items <= rb::kMaxNb, RB_E_ARG, "reyna_b200_csr_pattern_local: an element has more than 64 contributing interior faces");
    const int d = (p + 1) * (p + 2) / 2;
    const int blocks = grid_for((g->n_rows + 31) / 32, kPatWarps, kNumSMs * kPatBlocksPerSM);
    return launch_pattern<true>(g, s, p, d, blocks, indptr, indices, (cudaStream_t)stream);
}

extern "C" size_t reyna_b200_compact_workspace(int32_t n_scalar_rows) {
    return sizeof(int32_t) * ((size_t)n_scalar_rows + 1 + scan_scratch_ints((int64_t)n_scalar_rows + 1) + 16);
}

extern "C" int reyna_b200_compact_count(int32_t n, const int32_t* indptr, const double* data, int32_t* out_indptr,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    RB_REQUIRE(indptr && data && out_indptr && workspace, RB_E_ARG, "reyna_b200_compact_count: null argument");
    RB_REQUIRE(workspace_bytes >= reyna_b200_compact_workspace(n), RB_E_WORKSPACE,
               "reyna_b200_compact_count: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* row_nnz = (int32_t*)workspace;
    int32_t* scratch = row_nnz + (n + 1);
    const int blocks = grid_for((int64_t)n + 1, 8, kNumSMs * 16);
    compact_count_kernel<<<blocks, 256, 0, st>>>(n, indptr, data, row_nnz);
    RB_LAUNCH_CHECK("compact_count_kernel");
    return exclusive_scan_i32(row_nnz, out_indptr, (int64_t)n + 1, scratch, nullptr, st);
}

extern "C" int reyna_b200_compact_fill(int32_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                                       const int32_t* out_indptr, int32_t* out_indices, double* out_data,
                                       void* stream) {
    RB_REQUIRE(indptr && indices && data && out_indptr && out_indices && out_data, RB_E_ARG,
               "reyna_b200_compact_fill: null argument");
    const int blocks = grid_for(n, 8, kNumSMs * 16);
    compact_fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, indptr, indices, data, out_indptr, out_indices,
                                                                  out_data);
    RB_LAUNCH_CHECK("compact_fill_kernel");
    return 0;
}

extern "C" const char* reyna_b200_last_error(void) { return rb::g_err; }
extern "C" unsigned long long reyna_b200_launch_count(void) { return rb::g_launches; }
extern "C" const char* reyna_b200_version(void) { return "reyna_b200 0.1 (sm_100a)"; }
