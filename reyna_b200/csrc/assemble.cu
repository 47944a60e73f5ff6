// C entry point of the DG assembly (volume + interior-face terms -> CSR values) and its geometry prepass.
//
// Replaces, for all elements at once,
//   DGFEM._stiffness_matrix + localstiff            reyna/DGFEM/two_dimensional/main.py:275-327,
//                                                   _auxilliaries/assembly/full_assembly.py:9-77
//   DGFEM._interior_stiffness_matrix + int_localstiff   main.py:329-374, full_assembly.py:80-214
//   the scipy COO->CSR conversions and `B += ...`    main.py:319-320, 371-372, 211
// The kernels themselves are in assemble_items.cu (one lane per sub-triangle / per (element, face) item).
#include "assemble_common.cuh"
#include <stdlib.h>

namespace rb {

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

// per-element basis scalings (needed by both assembly kernels)
__global__ void __launch_bounds__(256) dg_prepare_boxes_kernel(rb_geom g, ElemBox* __restrict__ ebox) {
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.n_elem; e += stride) ebox[e] = make_box(g.bbox, e);
}

// per-face record: flattened geometry + SIPG penalty (needed by the face kernel only)
__global__ void __launch_bounds__(256) dg_prepare_faces_kernel(rb_geom g, const double* __restrict__ if_a_mid, double sigma_pp,
                                                               double p2, FaceRec* __restrict__ frec) {
    const int stride = gridDim.x * blockDim.x;
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
        }
        FaceRec r;
        r.midx = (P0.x + P1.x) * 0.5; r.midy = (P0.y + P1.y) * 0.5;     // same expressions as face_setup
        r.tanx = tx; r.tany = ty;
        r.nx = n.x; r.ny = n.y;
        r.De = De; r.sigma = sigma;
        frec[f] = r;
    }
}

// geometry tables (once per geometry): the face record without the penalty, and |k_b|_K of both elements of every face
__global__ void __launch_bounds__(256) dg_geometry_faces_kernel(rb_geom g, FaceRec* __restrict__ frec, double2* __restrict__ fkb) {
    const int stride = gridDim.x * blockDim.x;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.n_iface; f += stride) {
        const double2 P0 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f]];
        const double2 P1 = reinterpret_cast<const double2*>(g.nodes)[g.iedge[2 * (size_t)f + 1]];
        const double2 n = reinterpret_cast<const double2*>(g.inorm)[f];
        const double ex = P1.x - P0.x, ey = P1.y - P0.y;
        const double tx = 0.5 * ex, ty = 0.5 * ey;
        const int e1 = g.ielem[2 * (size_t)f], e2 = g.ielem[2 * (size_t)f + 1];
        fkb[f] = make_double2(max_sub_area_g(g, e1, P0, ex, ey), max_sub_area_g(g, e2, P0, ex, ey));
        FaceRec r;
        r.midx = (P0.x + P1.x) * 0.5; r.midy = (P0.y + P1.y) * 0.5;     // same expressions as dg_prepare_faces_kernel
        r.tanx = tx; r.tany = ty;
        r.nx = n.x; r.ny = n.y;
        r.De = sqrt(tx * tx + ty * ty); r.sigma = 0.0;
        frec[f] = r;
    }
}

// per-assembly prepass when the geometry tables exist: only the SIPG penalty depends on the coefficients
__global__ void __launch_bounds__(256) dg_face_sigma_kernel(rb_geom g, const double* __restrict__ if_a_mid, double sigma_pp, double p2,
                                                            const FaceRec* __restrict__ frec, const double2* __restrict__ fkb,
                                                            double* __restrict__ fsigma) {
    const int stride = gridDim.x * blockDim.x;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.n_iface; f += stride) {
        const double2 n = make_double2(frec[f].nx, frec[f].ny);
        const double De = frec[f].De;
        const int e1 = g.ielem[2 * (size_t)f], e2 = g.ielem[2 * (size_t)f + 1];
        const double2 m0 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f];
        const double2 m1 = reinterpret_cast<const double2*>(if_a_mid)[2 * (size_t)f + 1];
        const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;   // full_assembly.py:140
        const double a1 = g.area[e1], a2 = g.area[e2];
        const double2 kb = fkb[f];
        const double c1 = fmin(a1 / kb.x, p2);
        const double c2 = fmin(a2 / kb.y, p2);
        fsigma[f] = sigma_pp * lam * (2.0 * De) * fmax(c1 / a1, c2 / a2);        // full_assembly.py:146-148
    }
}

}  // namespace rb

using namespace rb;

static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

extern "C" size_t reyna_b200_geometry_tables_bytes(int32_t n_elem, int32_t n_iface) {
    const size_t ne = (size_t)(n_elem > 0 ? n_elem : 1), nf = (size_t)(n_iface > 0 ? n_iface : 1);
    return align16(sizeof(ElemBox) * ne) + align16(sizeof(FaceRec) * nf) + align16(sizeof(double2) * nf) + 16;
}

extern "C" int reyna_b200_geometry_tables(const rb_geom* g, void* tables, size_t tables_bytes, void* stream) {
    RB_REQUIRE(g && tables, RB_E_ARG, "reyna_b200_geometry_tables: null argument");
    RB_REQUIRE(((uintptr_t)tables & 15) == 0 && tables_bytes >= reyna_b200_geometry_tables_bytes(g->n_elem, g->n_iface), RB_E_WORKSPACE,
               "reyna_b200_geometry_tables: buffer too small or not 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ne = (size_t)(g->n_elem > 0 ? g->n_elem : 1), nf = (size_t)(g->n_iface > 0 ? g->n_iface : 1);
    uintptr_t w = (uintptr_t)tables;
    ElemBox* ebox = reinterpret_cast<ElemBox*>(w);
    w += align16(sizeof(ElemBox) * ne);
    FaceRec* frec = reinterpret_cast<FaceRec*>(w);
    w += align16(sizeof(FaceRec) * nf);
    double2* fkb = reinterpret_cast<double2*>(w);
    if (g->n_elem > 0) {
        dg_prepare_boxes_kernel<<<grid_for(g->n_elem, 256, kNumSMs * 8), 256, 0, st>>>(*g, ebox);
        RB_LAUNCH_CHECK("dg_prepare_boxes_kernel");
    }
    if (g->n_iface > 0) {
        dg_geometry_faces_kernel<<<grid_for(g->n_iface, 256, kNumSMs * 8), 256, 0, st>>>(*g, frec, fkb);
        RB_LAUNCH_CHECK("dg_geometry_faces_kernel");
    }
    return 0;
}

extern "C" size_t reyna_b200_assemble_workspace(int32_t n_elem, int32_t n_iface, int32_t n_vsegs, int32_t p) {
    const size_t ne = (size_t)(n_elem > 0 ? n_elem : 1), nf = (size_t)(n_iface > 0 ? n_iface : 1);
    const size_t ns = (size_t)(n_vsegs > 0 ? n_vsegs : 1);
    const size_t d = (size_t)(p + 1) * (p + 2) / 2;
    return align16(sizeof(ElemBox) * ne) + align16(sizeof(FaceRec) * nf) + align16(sizeof(double) * ns * (d * d + d)) +
           align16(sizeof(double) * nf) + 256;
}

// Shared argument check + kernel argument block of the three entry points below.
static int make_args(const char* who, const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q, double sigma_D,
                     double* data, double* load, long long* zero_count, void* workspace, size_t workspace_bytes, AsmArgs& A,
                     bool& diff, bool& adv) {
    (void)who;
    RB_REQUIRE(g && s && c && q && workspace, RB_E_ARG, "reyna_b200_assemble: null argument");
    const int p = q->p;
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_assemble: polynomial degree must be 1..3");
    const int n_vsegs = g->elem_seg0 ? g->n_vsegs : g->n_rows;
    const int d = (p + 1) * (p + 2) / 2;
    RB_REQUIRE(workspace_bytes >= reyna_b200_assemble_workspace(g->n_elem, g->n_iface, n_vsegs, p), RB_E_WORKSPACE,
               "reyna_b200_assemble: workspace too small");
    // which terms are present is decided by the FACE samples when there are faces (the volume samples may still be on their
    // way to the device when the prepass runs)
    diff = c->vol_a != nullptr || c->if_a != nullptr;
    adv = c->vol_b != nullptr || c->if_b != nullptr;
    RB_REQUIRE(diff || adv, RB_E_NOCOEF, "reyna_b200_assemble: no advection or diffusion specified");
    RB_REQUIRE(!diff || (c->if_a && c->if_a_mid) || g->n_iface == 0, RB_E_ARG,
               "reyna_b200_assemble: diffusion given without interior-face samples");
    RB_REQUIRE(!adv || (c->if_b && c->if_upwind) || g->n_iface == 0, RB_E_ARG,
               "reyna_b200_assemble: advection given without interior-face samples / upwind flags");
    memset(&A, 0, sizeof(A));
    A.g = *g;
    A.adj_off = s->adj_off;
    A.adj_item = s->adj_item;
    A.grp_off = s->grp_off;
    A.adj_nb = s->adj_nb;
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
    A.has_giants = s->max_items > items_max_face_items();
    // prepass tables and the volume sums (16-byte aligned slices of the workspace)
    const size_t ne = (size_t)(g->n_elem > 0 ? g->n_elem : 1), nf = (size_t)(g->n_iface > 0 ? g->n_iface : 1);
    uintptr_t w = ((uintptr_t)workspace + 15) & ~(uintptr_t)15;
    A.ebox = reinterpret_cast<ElemBox*>(w);
    w += align16(sizeof(ElemBox) * ne);
    A.frec = reinterpret_cast<FaceRec*>(w);
    w += align16(sizeof(FaceRec) * nf);
    A.vdiag = reinterpret_cast<double*>(w);
    w += align16(sizeof(double) * (size_t)(n_vsegs > 0 ? n_vsegs : 1) * ((size_t)d * d + d));
    double* fsigma = reinterpret_cast<double*>(w);
    w += align16(sizeof(double) * nf);
    A.vol_next = reinterpret_cast<int*>(w);      // inside the 256 spare bytes of the workspace
    A.fsigma = nullptr;
    A.fkb = nullptr;
    if (g->tables) {
        // geometry tables of reyna_b200_geometry_tables(): boxes and face records are read from there, the prepass only
        // computes the penalties (separate array: the cached records carry sigma = 0)
        uintptr_t t = (uintptr_t)g->tables;
        A.ebox = reinterpret_cast<const ElemBox*>(t);
        t += align16(sizeof(ElemBox) * ne);
        A.frec = reinterpret_cast<const FaceRec*>(t);
        t += align16(sizeof(FaceRec) * nf);
        A.fkb = reinterpret_cast<const double*>(t);
        if (diff && g->n_iface > 0) A.fsigma = fsigma;
    }
    return 0;
}

static int run_prepare(const AsmArgs& A, bool diff, cudaStream_t st) {
    const rb_geom& g = A.g;
    if (g.tables) {
        if (A.fsigma) {
            dg_face_sigma_kernel<<<grid_for(g.n_iface, 256, kNumSMs * 8), 256, 0, st>>>(
                g, A.c.if_a_mid, A.sigma_pp, A.p2, A.frec, reinterpret_cast<const double2*>(A.fkb), const_cast<double*>(A.fsigma));
            RB_LAUNCH_CHECK("dg_face_sigma_kernel");
        }
        return 0;
    }
    dg_prepare_boxes_kernel<<<grid_for(g.n_elem, 256, kNumSMs * 8), 256, 0, st>>>(g, const_cast<ElemBox*>(A.ebox));
    RB_LAUNCH_CHECK("dg_prepare_boxes_kernel");
    if (g.n_iface > 0) {
        dg_prepare_faces_kernel<<<grid_for(g.n_iface, 256, kNumSMs * 8), 256, 0, st>>>(g, diff ? A.c.if_a_mid : nullptr, A.sigma_pp,
                                                                                       A.p2, const_cast<FaceRec*>(A.frec));
        RB_LAUNCH_CHECK("dg_prepare_faces_kernel");
    }
    return 0;
}

static int run_range(AsmArgs& A, const rb_structure* s, int p, bool diff, bool adv, int elem_lo, int elem_hi, int vround_lo,
                     int vround_hi, cudaStream_t st, int phases = 3) {
    const rb_geom& g = A.g;
    RB_REQUIRE(A.data && A.load && A.zero_count, RB_E_ARG, "reyna_b200_assemble: null output");
    RB_REQUIRE(0 <= elem_lo && elem_lo <= elem_hi && elem_hi <= g.n_rows && 0 <= vround_lo && vround_lo <= vround_hi &&
                   vround_hi <= g.n_vrounds,
               RB_E_ARG, "reyna_b200_assemble: invalid element / volume-round range");
    if (elem_lo == elem_hi) return 0;
    RB_REQUIRE(!diff || A.c.vol_a, RB_E_ARG, "reyna_b200_assemble: diffusion without volume samples");
    RB_REQUIRE(!adv || A.c.vol_b, RB_E_ARG, "reyna_b200_assemble: advection without volume samples");
    RB_REQUIRE(g.n_tri == 0 || (g.tri_xy && g.vround_tri0 && g.n_vrounds > 0), RB_E_ARG,
               "reyna_b200_assemble: the geometry carries no volume stream (rb_geom.tri_xy / vround_tri0)");
    RB_REQUIRE((g.tri_seg == nullptr) == (g.elem_seg0 == nullptr), RB_E_ARG,
               "reyna_b200_assemble: tri_seg and elem_seg0 must be given together");
    RB_REQUIRE(!(phases & 2) || s->max_items <= 64, RB_E_ARG,
               "reyna_b200_assemble: an element has more than 64 contributing interior faces");
    RB_REQUIRE(((uintptr_t)A.data & 15) == 0 && ((uintptr_t)A.load & 7) == 0, RB_E_ARG,
               "reyna_b200_assemble: data must be 16-byte aligned");
    const rb_coefs& c = A.c;
    RB_REQUIRE((!c.vol_a || ((uintptr_t)c.vol_a & 15) == 0) && (!c.vol_b || ((uintptr_t)c.vol_b & 15) == 0) &&
                   (!c.vol_c || ((uintptr_t)c.vol_c & 15) == 0) && (!c.vol_f || ((uintptr_t)c.vol_f & 15) == 0) &&
                   (!g.tri_xy || ((uintptr_t)g.tri_xy & 15) == 0),
               RB_E_ARG, "reyna_b200_assemble: volume sample arrays must be 16-byte aligned");
    A.elem_lo = elem_lo;
    A.elem_hi = elem_hi;
    A.vround_lo = vround_lo;
    A.vround_hi = vround_hi;
    return launch_assemble_items(A, p, diff, adv, st, phases);
}

extern "C" int reyna_b200_assemble(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                   double sigma_D, double* data, double* load, long long* zero_count, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    RB_REQUIRE(data && load && zero_count, RB_E_ARG, "reyna_b200_assemble: null argument");
    AsmArgs A;
    bool diff, adv;
    int rc = make_args("reyna_b200_assemble", g, s, c, q, sigma_D, data, load, zero_count, workspace, workspace_bytes, A, diff, adv);
    if (rc) return rc;
    if (g->n_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    rc = run_prepare(A, diff, st);
    if (rc) return rc;
    return run_range(A, s, q->p, diff, adv, 0, g->n_rows, 0, g->n_vrounds, st);
}

extern "C" int reyna_b200_assemble_prepare(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                           double sigma_D, void* workspace, size_t workspace_bytes, void* stream) {
    AsmArgs A;
    bool diff, adv;
    int rc = make_args("reyna_b200_assemble_prepare", g, s, c, q, sigma_D, nullptr, nullptr, nullptr, workspace, workspace_bytes, A,
                       diff, adv);
    if (rc) return rc;
    if (g->n_rows == 0) return 0;
    return run_prepare(A, diff, (cudaStream_t)stream);
}

extern "C" int reyna_b200_assemble_range(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                         double sigma_D, double* data, double* load, long long* zero_count, void* workspace,
                                         size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                                         int32_t vround_hi, void* stream) {
    RB_REQUIRE(data && load && zero_count, RB_E_ARG, "reyna_b200_assemble_range: null argument");
    AsmArgs A;
    bool diff, adv;
    int rc = make_args("reyna_b200_assemble_range", g, s, c, q, sigma_D, data, load, zero_count, workspace, workspace_bytes, A, diff,
                       adv);
    if (rc) return rc;
    return run_range(A, s, q->p, diff, adv, elem_lo, elem_hi, vround_lo, vround_hi, (cudaStream_t)stream);
}

// The two halves of reyna_b200_assemble_range: the volume kernel needs neither the adjacency nor the CSR layout, so a caller
// may run it while reyna_b200_structure_known / reyna_b200_csr_pattern are still working on another stream, and launch the
// face kernel (which needs both, and the volume sums) afterwards.
extern "C" int reyna_b200_assemble_volume_range(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                                double sigma_D, double* data, double* load, long long* zero_count, void* workspace,
                                                size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                                                int32_t vround_hi, void* stream) {
    RB_REQUIRE(data && load && zero_count, RB_E_ARG, "reyna_b200_assemble_volume_range: null argument");
    AsmArgs A;
    bool diff, adv;
    int rc = make_args("reyna_b200_assemble_volume_range", g, s, c, q, sigma_D, data, load, zero_count, workspace, workspace_bytes, A,
                       diff, adv);
    if (rc) return rc;
    return run_range(A, s, q->p, diff, adv, elem_lo, elem_hi, vround_lo, vround_hi, (cudaStream_t)stream, 1);
}

extern "C" int reyna_b200_assemble_faces_range(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                               double sigma_D, double* data, double* load, long long* zero_count, void* workspace,
                                               size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                                               int32_t vround_hi, void* stream) {
    RB_REQUIRE(data && load && zero_count, RB_E_ARG, "reyna_b200_assemble_faces_range: null argument");
    AsmArgs A;
    bool diff, adv;
    int rc = make_args("reyna_b200_assemble_faces_range", g, s, c, q, sigma_D, data, load, zero_count, workspace, workspace_bytes, A,
                       diff, adv);
    if (rc) return rc;
    return run_range(A, s, q->p, diff, adv, elem_lo, elem_hi, vround_lo, vround_hi, (cudaStream_t)stream, 2);
}
