// Boundary-face terms: Dirichlet SIPG on the elliptic boundary and upwind inflow on the hyperbolic inflow boundary,
// subtracted in place from the diagonal blocks and the load AFTER the fused volume/interior kernel, so that every entry
// is formed as ((volume + interior) - dirichlet) - inflow, the order of DGFEM.dgfem (main.py:210-225).
//
// Replaces DGFEM._diffusion_bcs_contribution + local_diffusion_dirichlet (main.py:376-427, boundary_assembly.py:59-137)
// and DGFEM._advection_bcs_contribution + local_advection_inflow (main.py:429-474, boundary_assembly.py:9-56).
// There are only O(sqrt(E)) boundary faces: one thread per boundary element, faces visited in ascending index order
// (the reference's accumulation order), accumulators in local memory -- this kernel is latency-, not throughput-bound.
#include "basis.cuh"

namespace rb {

struct BndArgs {
    rb_geom g;
    const int32_t* adj_off;
    const int32_t* adj_item;
    const int32_t* grp_off;
    rb_coefs c;
    double we[RB_MAX_QE], xe[RB_MAX_QE];
    LegConst lc;
    double sigma_pp, p2;
    int n_bnd;
    const int32_t* bnd_elem;
    const int32_t* bnd_off;
    const int32_t* bnd_face;
    const uint8_t* flags;
    double* data;
    double* load;
    double* spill;
    unsigned long long* zero_count;
};

template <int P>
__global__ void __launch_bounds__(64) dg_boundary_kernel(const __grid_constant__ BndArgs A) {
    constexpr int D = BasisDim<P>::D, DD = D * D, QE = P + 1;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.n_bnd) return;
    const int e = A.bnd_elem[k];
    const ElemBox bx = make_box(A.g.bbox, e);
    double Zd[DD], Za[DD], fd[D], fa[D];
    for (int v = 0; v < DD; ++v) Zd[v] = Za[v] = 0.0;
    for (int v = 0; v < D; ++v) fd[v] = fa[v] = 0.0;
    double spill = 0.0;
    bool any_d = false, any_a = false, spill_d = false;
    const size_t FB = (size_t)A.g.n_bface;
    for (int a = A.bnd_off[k]; a < A.bnd_off[k + 1]; ++a) {
        const int f = A.bnd_face[a];
        const int fl = A.flags[f];
        if (!(fl & 3)) continue;
        const int v0 = A.g.bedge[2 * (size_t)f], v1 = A.g.bedge[2 * (size_t)f + 1];
        const double2 P0 = reinterpret_cast<const double2*>(A.g.nodes)[v0];
        const double2 P1 = reinterpret_cast<const double2*>(A.g.nodes)[v1];
        const double2 mid = make_double2((P0.x + P1.x) * 0.5, (P0.y + P1.y) * 0.5);
        const double2 tan = make_double2(0.5 * (P1.x - P0.x), 0.5 * (P1.y - P0.y));
        const double De = sqrt(tan.x * tan.x + tan.y * tan.y);
        const double2 n = reinterpret_cast<const double2*>(A.g.bnorm)[f];
        double sigma = 0.0;
        if (fl & 1) {
            // boundary_assembly.py:103-108
            const double2 m0 = reinterpret_cast<const double2*>(A.c.bf_a_mid)[2 * (size_t)f];
            const double2 m1 = reinterpret_cast<const double2*>(A.c.bf_a_mid)[2 * (size_t)f + 1];
            const double lam = (n.x * m0.x + n.y * m1.x) * n.x + (n.x * m0.y + n.y * m1.y) * n.y;
            const double ex = P1.x - P0.x, ey = P1.y - P0.y;
            double kb = 0.0;
            for (int q = A.g.poly_off[e]; q < A.g.poly_off[e + 1]; ++q) {
                const double2 v = reinterpret_cast<const double2*>(A.g.nodes)[A.g.poly_vtx[q]];
                kb = fmax(kb, 0.5 * fabs(ex * (v.y - P0.y) - ey * (v.x - P0.x)));
            }
            const double ar = A.g.area[e];
            sigma = A.sigma_pp * lam * (2.0 * De) * fmin(ar / kb, A.p2) / ar;
            any_d = true;
            if (fl & 4) spill_d = true;
        }
        if (fl & 2) any_a = true;
        for (int q = 0; q < QE; ++q) {
            const double x = mid.x + A.xe[q] * tan.x, y = mid.y + A.xe[q] * tan.y;
            const double W = De * A.we[q];
            const size_t m = (size_t)q * FB + f;
            double phi[D], gx[D], gy[D];
            eval_basis<P, true>(x, y, bx, A.lc, phi, gx, gy);
            const double gD = A.c.bf_g[m];
            if (fl & 1) {
                const double2 ar0 = reinterpret_cast<const double2*>(A.c.bf_a)[2 * m];
                const double2 ar1 = reinterpret_cast<const double2*>(A.c.bf_a)[2 * m + 1];
                // coe = a . n  (boundary_assembly.py:114);  G_i = grad(phi_i) . coe
                const double c0 = ar0.x * n.x + ar0.y * n.y, c1 = ar1.x * n.x + ar1.y * n.y;
                double G[D];
                for (int i = 0; i < D; ++i) G[i] = gx[i] * c0 + gy[i] * c1;
                for (int r = 0; r < D; ++r) {
                    for (int cidx = 0; cidx < D; ++cidx)
                        Zd[r * D + cidx] += W * (G[r] * phi[cidx] + phi[r] * G[cidx] - sigma * phi[r] * phi[cidx]);
                    fd[r] += W * gD * (G[r] - sigma * phi[r]);
                }
            }
            if (fl & 2) {
                const double2 bb = reinterpret_cast<const double2*>(A.c.bf_b)[m];
                const double wb = W * (bb.x * n.x + bb.y * n.y);
                for (int r = 0; r < D; ++r) {
                    for (int cidx = 0; cidx < D; ++cidx) Za[r * D + cidx] += wb * phi[r] * phi[cidx];
                    fa[r] += wb * gD * phi[r];
                }
            }
        }
    }
    if (!any_d && !any_a) return;
    // position of the diagonal block of e: groups are sorted by neighbour, the diagonal sits after those < e
    int dslot = 0;
    for (int a = A.adj_off[e]; a < A.adj_off[e + 1]; ++a) {
        const int it = A.adj_item[a];
        if (it < 0) continue;
        const int o = A.g.ielem[2 * (size_t)(it >> 1) + (1 - (it & 1))];
        dslot += A.g.elem_gid ? (A.g.elem_gid[o] < A.g.elem_gid[e]) : (o < e);
    }
    const int nb = A.grp_off[e + 1] - A.grp_off[e] + 1;
    const int64_t base = (int64_t)DD * (A.grp_off[e] + e) + (int64_t)dslot * D;
    unsigned zeros = 0;
    for (int r = 0; r < D; ++r) {
        for (int cidx = 0; cidx < D; ++cidx) {
            // Z[r][c] is subtracted from B[(e,c),(e,r)] (main.py:394-400, F-order flatten)
            double* p = A.data + base + (int64_t)cidx * nb * D + r;
            double v = *p;
            if (any_d) {
                if (spill_d) spill += Zd[r * D + cidx];
                else v -= Zd[r * D + cidx];
            }
            if (any_a) v -= Za[r * D + cidx];
            *p = v;
            zeros += (v == 0.0);
        }
        double l = A.load[(size_t)e * D + r];
        if (any_d) l -= fd[r];
        if (any_a) l -= fa[r];
        A.load[(size_t)e * D + r] = l;
    }
    if (A.spill) A.spill[k] = spill;
    if (zeros) atomicAdd(A.zero_count, (unsigned long long)zeros);
}

}  // namespace rb

using namespace rb;

extern "C" int reyna_b200_boundary(const rb_geom* g, const rb_structure* s, const rb_coefs* c, const rb_quad* q,
                                   double sigma_D, int32_t n_bnd, const int32_t* bnd_elem, const int32_t* bnd_off,
                                   const int32_t* bnd_face, const uint8_t* flags, double* data, double* load,
                                   double* spill, long long* zero_count, void* stream) {
    RB_REQUIRE(g && s && c && q && data && load && zero_count, RB_E_ARG, "reyna_b200_boundary: null argument");
    if (n_bnd <= 0) return 0;
    RB_REQUIRE(bnd_elem && bnd_off && bnd_face && flags && c->bf_g, RB_E_ARG, "reyna_b200_boundary: null boundary lists");
    const int p = q->p;
    RB_REQUIRE(p >= 1 && p <= RB_MAX_P, RB_E_DEGREE, "reyna_b200_boundary: polynomial degree must be 1..3");
    BndArgs A;
    memset(&A, 0, sizeof(A));
    A.g = *g;
    A.adj_off = s->adj_off;
    A.adj_item = s->adj_item;
    A.grp_off = s->grp_off;
    A.c = *c;
    for (int i = 0; i <= p; ++i) {
        A.we[i] = q->we[i];
        A.xe[i] = q->xe[i];
    }
    A.lc = make_leg_const();
    A.sigma_pp = sigma_D * (double)(p * p);
    A.p2 = (double)(p * p);
    A.n_bnd = n_bnd;
    A.bnd_elem = bnd_elem;
    A.bnd_off = bnd_off;
    A.bnd_face = bnd_face;
    A.flags = flags;
    A.data = data;
    A.load = load;
    A.spill = spill;
    A.zero_count = reinterpret_cast<unsigned long long*>(zero_count);
    const int blocks = (int)div_up(n_bnd, 64);
    cudaStream_t st = (cudaStream_t)stream;
    switch (p) {
        case 1: dg_boundary_kernel<1><<<blocks, 64, 0, st>>>(A); break;
        case 2: dg_boundary_kernel<2><<<blocks, 64, 0, st>>>(A); break;
        default: dg_boundary_kernel<3><<<blocks, 64, 0, st>>>(A); break;
    }
    RB_LAUNCH_CHECK("dg_boundary_kernel");
    return 0;
}
