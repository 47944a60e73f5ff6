// Shared declarations of the DG assembly kernels (assemble.cu: prepass + C entry point, assemble_items.cu: volume and face
// kernels).
#pragma once
#include "basis.cuh"

namespace rb {

// per-face record written by dg_prepare_kernel: everything an (element, face) item needs besides the coefficient samples
struct __align__(16) FaceRec {
    double midx, midy;   // edge mid point
    double tanx, tany;   // (P1 - P0) / 2
    double nx, ny;       // unit normal K1 -> K2
    double De;           // |tan| = half edge length
    double sigma;        // SIPG penalty (0 without diffusion)
};

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
    int has_dups;      // adjacency contains merged (duplicate-neighbour) items
    int has_giants;    // an element has more than 32 (element, face) items: dg_face_giant_kernel takes those
    const ElemBox* ebox;   // [n_elem]  bounding-box scalings, written by dg_prepare_kernel
    const FaceRec* frec;   // [n_iface] flattened face geometry + SIPG penalty, written by dg_prepare_kernel
    const int32_t* adj_nb; // [n_items] neighbour element of every adjacency item (rb_structure.adj_nb)
    double* vdiag;         // [n_vsegs][d*d+d] per-segment volume sums (block + load): dg_volume_kernel -> dg_face_kernel
    const double* fsigma;      // [n_iface] SIPG penalties when the face records come from the geometry tables (else in frec)
    const double* fkb;         // [n_iface][2] |k_b|_K of both elements of a face (geometry tables)
    int* vol_next;             // round counter of the persistent volume kernel (workspace)
    int window;                // elements per CTA of the face kernel
    int elem_lo, elem_hi;      // elements assembled by this launch (a range of whole volume rounds)
    int vround_lo, vround_hi;  // their volume rounds
};

// item-per-lane kernels (assemble_items.cu)
// phases: bit 0 = volume kernel, bit 1 = face kernel(s)
int launch_assemble_items(const AsmArgs& A, int p, bool diff, bool adv, cudaStream_t st, int phases);
int items_max_face_items();
int items_max_segments();

}  // namespace rb
