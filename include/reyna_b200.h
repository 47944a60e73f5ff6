/*
 * reyna_b200 -- C-ABI of the B200-native DGFEM assemble-and-solve path.
 *
 * The reference (mattevs24/reyna, pure Python) has no FFI/plugin layer: its hot path is the Python method
 * DGFEM.dgfem() (reyna/DGFEM/two_dimensional/main.py:181-231).  This header is the seam a maintainer would bind
 * with ctypes from that method; every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer inside rb_* structs is a DEVICE pointer unless the field comment
 *     says "host"; all buffers are caller-owned (the library allocates nothing persistent);
 *   - every call is stream-ordered on the caller-supplied cudaStream_t (passed as void*); none synchronises
 *     unless documented;
 *   - return value: 0 = ok, <0 = invalid argument (RB_E_*), >0 = cudaError_t; text via reyna_b200_last_error()
 *     (thread-local); no exception crosses the boundary;
 *   - all floating point is IEEE binary64; indices are int32 (as in the reference's final scipy CSR).
 *
 * Data layout
 *   DoF (e,k) -> e*d + k, d = (p+1)(p+2)/2, k enumerates (i,j), i+j<=p row-major
 *   (polygonal_basis_utils.py:7-26).  The matrix is scalar CSR whose rows (e,j) hold, per neighbouring element in
 *   ascending order (the element itself included), d consecutive columns: exactly scipy's canonical CSR of B.
 *   Face and boundary coefficient samples are "q-major": value of quadrature point q of item t at [q*n_items + t];
 *   volume samples follow the round order described at rb_geom.
 */
#ifndef REYNA_B200_H
#define REYNA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_MAX_P 3            /* polynomial degrees 1..3 have register-resident kernels                      */
#define RB_MAX_QV 16          /* (p+1)^2 volume points                                                       */
#define RB_MAX_QE 4           /* p+1 face points                                                             */

enum {
    RB_OK = 0,
    RB_E_ARG = -1,            /* null / inconsistent argument                                                */
    RB_E_DEGREE = -2,         /* polynomial degree outside 1..RB_MAX_P                                       */
    RB_E_WORKSPACE = -3,      /* workspace too small                                                         */
    RB_E_NOCOEF = -4,         /* neither diffusion nor advection given (main.py:172-173)                     */
    RB_E_OVERFLOW = -5        /* nnz does not fit int32                                                      */
};

/* Flat DGFEMGeometry (reyna/geometry/two_dimensional/DGFEM.py:64-86).  Elements [0,n_rows) are assembled; elements
 * [n_rows,n_elem) are read-only ghosts (multi-GPU partitions).  Triangles belong to row elements only. */
typedef struct {
    int32_t n_nodes, n_elem, n_rows, n_tri, n_iface, n_bface;
    const double  *nodes;      /* [n_nodes][2]                      geometry.nodes                           */
    const int32_t *tri;        /* [n_tri][3]                        geometry.subtriangulation                */
    const int32_t *tri_elem;   /* [n_tri] non-decreasing            geometry.triangle_to_polygon             */
    const int32_t *tri_off;    /* [n_rows+1] CSR offsets of tri_elem                                         */
    const double  *bbox;       /* [n_elem][4] xmin,xmax,ymin,ymax   geometry.elem_bounding_boxes             */
    const double  *area;       /* [n_elem]                          geometry.areas                           */
    const int32_t *poly_off;   /* [n_elem+1]                        ragged mesh.filtered_regions             */
    const int32_t *poly_vtx;   /* [poly_off[n_elem]]                                                          */
    const int32_t *iedge;      /* [n_iface][2]                      geometry.interior_edges                  */
    const int32_t *ielem;      /* [n_iface][2] (K1,K2)              geometry.interior_edges_to_element       */
    const double  *inorm;      /* [n_iface][2] unit, K1 -> K2       geometry.interior_normals                */
    const int32_t *bedge;      /* [n_bface][2]                      geometry.boundary_edges                  */
    const int32_t *belem;      /* [n_bface]                         geometry.boundary_edges_to_element       */
    const double  *bnorm;      /* [n_bface][2] outward              geometry.boundary_normals                */
    const int32_t *elem_gid;   /* [n_elem] global element id of every local element (multi-GPU partitions: block columns are
                                  ordered and numbered by it); NULL = identity                              */
    /* Volume stream (host-built once per geometry, reyna_b200/stream.py): sub-triangles are packed, in order, into
     * "volume rounds" of whole elements with at most 32 triangles (one warp, one lane per triangle).  An element with
     * more than 32 triangles fills rounds of its own; its pieces are "segments".  The volume samples of rb_coefs are
     * ordered round by round, q-major inside a round: sample q of the l-th triangle of round r (n_r triangles, first
     * triangle t_r = vround_tri0[r]) sits at index Qv*t_r + q*n_r + l -- every round is one contiguous range.        */
    const double  *tri_xy;     /* [n_tri][3][2] vertex coordinates of every sub-triangle (nodes[tri])         */
    const int32_t *vround_tri0;/* [n_vrounds+1] first triangle of every volume round                          */
    const int32_t *tri_seg;    /* [n_tri] segment of every triangle; NULL = tri_elem (no element has > 32 triangles) */
    const int32_t *elem_seg0;  /* [n_rows+1] first segment of every element; NULL = identity                  */
    int32_t n_vrounds;
    int32_t n_vsegs;           /* number of segments (== n_rows when elem_seg0 is NULL)                       */
    const void *tables;        /* geometry tables written once per geometry by reyna_b200_geometry_tables() (basis scalings
                                  of every element, flattened face geometry, |k_b|_K of the penalty): with them the
                                  per-assembly prepass only computes the SIPG penalties; NULL = it computes everything */
} rb_geom;

/* Geometry-only tables of the assembly (they depend on nodes / faces / polygons, not on coefficients or the degree). */
size_t reyna_b200_geometry_tables_bytes(int32_t n_elem, int32_t n_iface);
int reyna_b200_geometry_tables(const rb_geom *g, void *tables, size_t tables_bytes, void *stream);

/* Reference quadrature (HOST arrays; DGFEM._intialise_quadrature, main.py:233-257). */
typedef struct {
    int32_t p;                 /* polynomial degree                                                          */
    double wv[RB_MAX_QV];      /* triangle weights (sum 4)                                                   */
    double rv[RB_MAX_QV][2];   /* unit-triangle points                                                       */
    double we[RB_MAX_QE];      /* Gauss-Legendre weights on [-1,1]                                           */
    double xe[RB_MAX_QE];      /* Gauss-Legendre nodes                                                       */
} rb_quad;

/* Batched coefficient samples, evaluated ONCE by the caller on the quadrature points (replaces the per-item calls
 * in full_assembly.py:57,65,69,73,140,159,194-195 and boundary_assembly.py:44-45,103,111-112).  NULL <=> the
 * coefficient is None. */
typedef struct {
    /* volume samples: [Qv*n_tri] items in the round order of rb_geom (NOT q-major over the whole mesh); vol_c / vol_f
     * must be readable up to the next 16-byte boundary behind their last sample (bulk copies move 16-byte units) */
    const double *vol_a;       /* [Qv*n_tri][2][2]   diffusion at volume points                               */
    const double *vol_b;       /* [Qv*n_tri][2]      advection                                                */
    const double *vol_c;       /* [Qv*n_tri]         reaction                                                 */
    const double *vol_f;       /* [Qv*n_tri]         forcing                                                  */
    const double *if_a;        /* [Qe][n_iface][2][2]                                                         */
    const double *if_b;        /* [Qe][n_iface][2]                                                            */
    const double *if_a_mid;    /* [n_iface][2][2]    diffusion at the edge midpoint                           */
    const uint8_t *if_upwind;  /* [n_iface] 1 <=> b(mid).n >= 1e-12, K1 is upwind (full_assembly.py:194)      */
    const double *bf_a;        /* [Qe][n_bface][2][2]                                                         */
    const double *bf_b;        /* [Qe][n_bface][2]                                                            */
    const double *bf_g;        /* [Qe][n_bface]      Dirichlet data                                           */
    const double *bf_a_mid;    /* [n_bface][2][2]                                                             */
} rb_coefs;

/* Block structure of B (replaces scipy's COO->CSR sort, main.py:319-320,371-372).  Built by
 * reyna_b200_structure(); all arrays caller-allocated. */
typedef struct {
    int32_t *adj_off;          /* [n_rows+1]   contributing (face,side) items per row element                 */
    int32_t *adj_item;         /* [<=2*n_iface] item = face*2+side, sorted by (neighbour, face); bit31 = same
                                  neighbour as the previous item (merged block)                              */
    int32_t *adj_nb;           /* [<=2*n_iface] neighbour element (local id) of every item                    */
    int32_t *grp_off;          /* [n_rows+1]   distinct neighbours per row element (== adj_off without dups)  */
    int32_t *counts;           /* [4] device scratch: n_items, n_groups, n_dup_items, max_items_per_element   */
    int32_t n_items, n_groups, n_dup_items, max_items;   /* host copies, filled by reyna_b200_structure()     */
} rb_structure;

size_t reyna_b200_structure_workspace(int32_t n_rows, int32_t n_iface);

/* Element -> face adjacency and block-row layout.  has_diffusion != 0: every interior face contributes rows to both
 * of its elements; otherwise only to the downwind element chosen by coefs->if_upwind (the reference leaves the
 * other half of the 2d x 2d face matrix exactly zero and scipy prunes it, main.py:211 + full_assembly.py:209-212). */
/* Synchronises the stream once, to return the four counters in *s (the caller sizes the CSR arrays from them). */
int reyna_b200_structure(const rb_geom *g, int has_diffusion, const uint8_t *if_upwind, rb_structure *s,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Same device work WITHOUT the synchronisation: for repeated assemblies on one geometry.  The four host counters of *s
 * must hold the values an earlier reyna_b200_structure() call returned for the same (geometry, has_diffusion, if_upwind)
 * -- with diffusion the structure does not depend on the coefficients at all. */
int reyna_b200_structure_known(const rb_geom *g, int has_diffusion, const uint8_t *if_upwind, rb_structure *s,
                               void *workspace, size_t workspace_bytes, void *stream);

/* CSR index arrays of the structural pattern: indptr [n_rows*d+1], indices [nnz] (nnz = d*d*(n_rows+n_groups)). */
int reyna_b200_csr_pattern(const rb_geom *g, const rb_structure *s, int p, int32_t *indptr, int32_t *indices,
                           void *stream);

/* Same pattern with LOCAL column numbering (owned elements [0,n_rows), ghosts behind them): the matrix the distributed
 * Krylov solver multiplies with (vector = owned coefficients followed by ghost coefficients). */
int reyna_b200_csr_pattern_local(const rb_geom *g, const rb_structure *s, int p, int32_t *indptr, int32_t *indices,
                                 void *stream);

/* Volume + interior-face terms, summed per element in a fixed order and written straight into the CSR value array
 * (replaces DGFEM._stiffness_matrix + localstiff, main.py:275-327 / full_assembly.py:9-77, and
 * DGFEM._interior_stiffness_matrix + int_localstiff, main.py:329-374 / full_assembly.py:80-214).
 * data [nnz], load [n_rows*d]; *zero_count (device int64, caller-zeroed) receives the number of exact zeros written.
 * workspace: reyna_b200_assemble_workspace() bytes of device scratch (per-element basis scalings, per-face records,
 * per-segment volume sums).  Limits (RB_E_ARG otherwise): at most 64 contributing interior faces per element
 * (rb_structure.max_items; elements with more than 32 take a slower one-warp-per-element kernel) and at most 8 volume
 * segments (256 sub-triangles) per element. */
size_t reyna_b200_assemble_workspace(int32_t n_elem, int32_t n_iface, int32_t n_vsegs, int32_t p);
int reyna_b200_assemble(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                        double sigma_D, double *data, double *load, long long *zero_count, void *workspace,
                        size_t workspace_bytes, void *stream);
/* The same assembly in two steps, for callers that pipeline the host evaluation of the volume coefficients with the
 * device work (DGFEM.dgfem evaluates the user's callables chunk by chunk): reyna_b200_assemble_prepare() runs the
 * geometry / penalty prepass (it needs the FACE samples only), reyna_b200_assemble_range() then assembles the block rows
 * of elements [elem_lo, elem_hi) = volume rounds [vround_lo, vround_hi) (a range of WHOLE rounds of rb_geom.vround_tri0),
 * reading only the volume samples of those rounds.  reyna_b200_assemble() == prepare + range over everything. */
int reyna_b200_assemble_prepare(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                                double sigma_D, void *workspace, size_t workspace_bytes, void *stream);
int reyna_b200_assemble_range(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                              double sigma_D, double *data, double *load, long long *zero_count, void *workspace,
                              size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                              int32_t vround_hi, void *stream);
/* The two halves of reyna_b200_assemble_range.  The volume kernel reads neither the adjacency nor the CSR layout: it may
 * run while reyna_b200_structure_known / reyna_b200_csr_pattern still work on another stream; the face half needs both
 * (and the volume sums) and is launched after them. */
int reyna_b200_assemble_volume_range(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                                     double sigma_D, double *data, double *load, long long *zero_count, void *workspace,
                                     size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                                     int32_t vround_hi, void *stream);
int reyna_b200_assemble_faces_range(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                                    double sigma_D, double *data, double *load, long long *zero_count, void *workspace,
                                    size_t workspace_bytes, int32_t elem_lo, int32_t elem_hi, int32_t vround_lo,
                                    int32_t vround_hi, void *stream);

/* Boundary faces: elliptic (Dirichlet SIPG) and inflow terms subtracted in place from the diagonal blocks and the
 * load (replaces _diffusion_bcs_contribution / _advection_bcs_contribution, main.py:376-474, and
 * boundary_assembly.py:9-137).  bnd_elem [n_bnd] = elements owning boundary faces, bnd_off [n_bnd+1] and bnd_face
 * list their boundary faces (ascending); flags[f] bit0 = elliptic, bit1 = inflow (boundary_information.py:76-91);
 * bit2 = "matrix part goes to B[0,0]" (the reference's index quirk, main.py:397-400); spill [n_bnd] receives those sums. */
int reyna_b200_boundary(const rb_geom *g, const rb_structure *s, const rb_coefs *c, const rb_quad *q,
                        double sigma_D, int32_t n_bnd, const int32_t *bnd_elem, const int32_t *bnd_off,
                        const int32_t *bnd_face, const uint8_t *flags, double *data, double *load, double *spill,
                        long long *zero_count, void *stream);

/* scipy's value-dependent pruning (csr_plus_csr / csr_minus_csr drop exact zeros, main.py:211-225): two passes.
 * pass 1 writes row_nnz [n+1]-shaped exclusive scan into out_indptr and the total into *nnz_out (device);
 * pass 2 copies the survivors.  Only needed when zero_count != 0. */
size_t reyna_b200_compact_workspace(int32_t n_scalar_rows);
int reyna_b200_compact_count(int32_t n_scalar_rows, const int32_t *indptr, const double *data, int32_t *out_indptr,
                             void *workspace, size_t workspace_bytes, void *stream);
int reyna_b200_compact_fill(int32_t n_scalar_rows, const int32_t *indptr, const int32_t *indices, const double *data,
                            const int32_t *out_indptr, int32_t *out_indices, double *out_data, void *stream);

/* Krylov solve replacing scipy.sparse.linalg.spsolve (main.py:231).  Block-Jacobi (d x d diagonal blocks)
 * preconditioned CG (method 0, symmetric) or BiCGStab (method 1).  The matrix is the block-structured CSR produced
 * above (blocked != 0) or any scalar CSR (blocked == 0, e.g. after compaction).  n_ghost extra vector entries
 * follow the n owned ones in x; when halo != NULL it is called back (host function) before every SpMV to refresh
 * them and to all-reduce dot products (multi-GPU); NULL for a single GPU. */
typedef struct {
    int32_t method;            /* 0 = CG, 1 = BiCGStab                                                       */
    int32_t max_iter;
    double  rtol;              /* on ||r||_2 / ||b||_2 (preconditioned recurrence, verified on the true one) */
    int32_t check_every;       /* iterations between host convergence checks                                 */
    double  accept_rtol;       /* a TRUE residual that stops improving at or below this is the rounding floor: the solve
                                  ends there as converged (the floor spsolve has too); above it the iteration restarts
                                  twice from the true residual before giving up.  <= 0: 1e-7                       */
    /* Flow-ordered block Gauss-Seidel sweep (single GPU, BiCGStab, no multigrid): for advection-reaction WITHOUT diffusion
     * the block rows only couple to their upwind neighbours, so in a topological order of the upwind graph B is block lower
     * triangular and one sweep solves the system; with cycles in the flow the sweep is the preconditioner.  sweep_elems
     * lists all elements level by level (level = 1 + max level of the upwind neighbours; elements on cycles last),
     * sweep_level_off the first entry of every level.  sweep_levels == 0: block Jacobi / multigrid as before.        */
    const int32_t *sweep_elems;      /* DEVICE [n / d]                                                              */
    const int32_t *sweep_level_off;  /* HOST   [sweep_levels + 1]                                                   */
    int32_t sweep_levels;
} rb_solve_opts;

typedef struct {
    int32_t iterations;
    double  rel_residual;      /* true residual ||b - A x|| / ||b|| at exit                                   */
    int32_t converged;         /* true residual <= rtol, or rounding floor reached at <= accept_rtol          */
    int32_t restarts;          /* restarts from the true residual after the recurrence residual converged     */
    int32_t stagnated;         /* the true residual no longer improves (floor, or a plateau when !converged)  */
    int32_t singular_blocks;   /* singular d x d diagonal blocks the block-Jacobi set-up replaced by identity */
} rb_solve_info;

typedef int (*rb_halo_fn)(void *user, double *x_dev, void *stream);          /* refresh ghosts of x            */
typedef int (*rb_allreduce_fn)(void *user, double *vals_dev, int n, void *stream);   /* in-place sum           */

/* Multigrid preconditioner (p-multigrid over the hierarchical Legendre modes, then aggregation multigrid on the
 * piecewise-constant level): makes the Krylov iteration count mesh independent.  The handle OWNS its device memory
 * (cudaMalloc; the only exception to "caller-owned buffers") and keeps pointers to the fine CSR arrays, which must stay
 * alive and must carry the STRUCTURAL block pattern produced by reyna_b200_csr_pattern (not a compacted matrix).
 * Elements are assumed to be numbered with spatial locality (aggregates are groups of 4 consecutive elements). */
typedef struct rb_mg rb_mg;
typedef struct {
    int32_t nu;                /* damped-Jacobi sweeps before and after every coarse correction (default 3)   */
    double  omega;             /* Jacobi damping (default 0.7)                                                */
    double  alpha;             /* scaling of the aggregation coarse corrections (default 1.8)                 */
    const double *elem_bbox;   /* DEVICE [n_elem][4] (xmin,xmax,ymin,ymax) or NULL.  When given (single-GPU hierarchy),
                                * the levels below P1 are AGGLOMERATED P1 spaces: groups of 4 consecutive elements,
                                * linear polynomials on the bounding box of the group, 3 x 3 block Jacobi smoothing --
                                * instead of the piecewise-constant aggregation chain                          */
    int32_t fp32_levels;       /* != 0: the smoothing sweeps of the agglomerated-P1 levels read a single-precision copy of the
                                * level matrices (half the traffic of the sweeps; vectors, residuals and the Krylov iteration
                                * stay in double precision)                                                            */
} rb_mg_opts;
/* Distributed hierarchy (dd != NULL): the fine and P1 levels are the rank's rows in LOCAL column numbering (halo exchange
 * through dd->dist), the piecewise-constant level is GLOBAL: the caller gathers every rank's P0 rows (global element
 * numbering) into one CSR matrix, which every rank coarsens and cycles redundantly -- the coarse space keeps its global
 * reach, so iteration counts do not grow with the number of GPUs. */
typedef struct rb_dist rb_dist;
typedef struct {
    rb_dist *dist;             /* NCCL context + halo plan (reyna_b200_dist_create)                          */
    int32_t rank, world;
    int32_t n_ghost_elem;      /* ghost elements behind the owned ones                                        */
    int32_t n_global_elem;
    const int32_t *bounds;     /* HOST [world+1]: rank r owns elements [bounds[r], bounds[r+1])               */
    const int32_t *p0_indptr;  /* DEVICE [n_global_elem+1]   global P0 matrix                                 */
    const int32_t *p0_indices; /* DEVICE [p0_nnz]            (global element ids, ascending per row)          */
    const double  *p0_data;    /* DEVICE [p0_nnz]                                                             */
    long long p0_nnz;
    /* Agglomerated-P1 hierarchy (optional; l1_indptr == NULL selects the piecewise-constant chain above).  Element ranges must
     * start at multiples of 4 so that no aggregate straddles two ranks.  The first agglomerated level (3 x 3 blocks, scalar
     * CSR, GLOBAL aggregate numbering, gathered by the caller) is replicated like the P0 matrix; rb_mg_opts.elem_bbox is
     * ignored in distributed mode. */
    const int32_t *l1_indptr;  /* DEVICE [3*n_agg+1], n_agg = ceil(n_global_elem / 4)                         */
    const int32_t *l1_indices; /* DEVICE [l1_nnz]                                                             */
    const double  *l1_data;    /* DEVICE [l1_nnz]                                                             */
    long long      l1_nnz;
    const double  *l1_bbox;    /* DEVICE [n_agg][4]  bounding boxes of the aggregates                         */
    const double  *l0_tr;      /* DEVICE [n_owned_elem][8]  {s00, syy, sy0, sxx, sx0, 0, 0, 0}: P1 basis of the
                                * aggregate expressed in the element's basis (see csrc/multigrid.cu)          */
} rb_mg_dist;
int reyna_b200_mg_create(int32_t n_rows_scalar, int32_t d, int32_t p, const int32_t *indptr, const int32_t *indices,
                         const double *data, const rb_mg_opts *opts, const rb_mg_dist *dd, rb_mg **out, void *stream);
void reyna_b200_mg_destroy(rb_mg *mg);
int reyna_b200_mg_info(const rb_mg *mg, int32_t *n_levels, int32_t *level_rows /*[16]*/, long long *level_nnz /*[16]*/);
/* z = coarse-space correction of r (the multigrid part of the preconditioner alone; exposed for tests) */
int reyna_b200_mg_apply(rb_mg *mg, const double *r, double *z, void *stream);

/* nnz > 0 adds room for the block-column list that lets the SpMV of a block-structured matrix (blocked != 0) skip the
 * per-entry column indices; with nnz == 0 the solver falls back to the generic CSR SpMV. */
size_t reyna_b200_solve_workspace(int32_t n_rows_scalar, int32_t n_ghost_scalar, int32_t d, long long nnz);
/* mg == NULL: block-Jacobi preconditioning only. */
int reyna_b200_solve(int32_t n_rows_scalar, int32_t n_ghost_scalar, int32_t d, int blocked,
                     const int32_t *indptr, const int32_t *indices, const double *data,
                     const double *rhs, double *x, const rb_solve_opts *opts, rb_solve_info *info,
                     rb_halo_fn halo, rb_allreduce_fn allreduce, void *user, rb_mg *mg,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Multi-GPU solve (one process per GPU): NCCL halo exchange + all-reduce behind the rb_halo_fn / rb_allreduce_fn hooks.
 * NCCL is resolved at run time (nccl_path, else the libnccl.so.2 already loaded by the process).  Rank 0 makes the
 * 128-byte unique id and distributes it (e.g. torch.distributed.broadcast); every rank creates the communicator ONCE
 * (reyna_b200_comm_create; seconds) and then, per partitioned system, a handle with the
 * halo plan: peers, per-peer offsets (in elements) into the concatenated send list (device int32 of owned element ids)
 * and into the ghost block (ghosts sorted by global id are contiguous per peer).  Pass reyna_b200_nccl_halo /
 * reyna_b200_nccl_allreduce and the handle (as `user`) to reyna_b200_solve; x then has n_ghost extra entries. */
int reyna_b200_nccl_unique_id(const char *nccl_path, char out[128]);
int reyna_b200_comm_create(const char *nccl_path, const char id[128], int rank, int world, void **comm_out);
void reyna_b200_comm_destroy(void *comm);
int reyna_b200_dist_create(void *comm, int rank, int world, int d, int n_rows, int n_peers,
                           const int *peer, const int *send_off, const int *recv_off, const int32_t *send_idx_dev,
                           rb_dist **out);
void reyna_b200_dist_destroy(rb_dist *h);
int reyna_b200_dist_counters(const rb_dist *h, unsigned long long *exchanges, unsigned long long *allreduces);
int reyna_b200_nccl_halo(void *user, double *x_dev, void *stream);
int reyna_b200_nccl_allreduce(void *user, double *vals_dev, int n, void *stream);

/* y = A x on the same matrix formats (exposed for tests and for the roofline measurement of the SpMV). */
int reyna_b200_spmv(int32_t n_rows_scalar, int32_t d, int blocked, const int32_t *indptr, const int32_t *indices,
                    const double *data, const double *x, double *y, void *stream);

/* Error norms of a computed solution (replaces DGFEM.errors, main.py:476-615, and error_assembly/full_assembly.py:9-187,
 * error_assembly/boundary_assembly.py:8-129).  Samples are q-major like rb_coefs; NULL <=> term absent. */
typedef struct {
    const double *vol_u;       /* [Qv][n_tri]        exact solution at the volume points                      */
    const double *vol_gu;      /* [Qv][n_tri][2]     its gradient (required iff vol_a given)                   */
    const double *vol_a;       /* [Qv][n_tri][2][2]  diffusion                                                 */
    const double *vol_c0;      /* [Qv][n_tri]        reaction - div(advection)/2  (main.py:517)                */
    const double *if_a_mid;    /* [n_iface][2][2]    diffusion at interior edge midpoints                      */
    const double *if_b;        /* [Qe][n_iface][2]   advection at interior face points                         */
    const double *bf_u;        /* [Qe][n_bface]      exact solution at boundary face points                    */
    const double *bf_a_mid;    /* [n_bface][2][2]                                                              */
    const double *bf_b;        /* [Qe][n_bface][2]                                                             */
    const uint8_t *bf_flags;   /* [n_bface] bit0 = elliptic (boundary_information.py:76-84)                    */
} rb_err_samples;

size_t reyna_b200_errors_workspace(void);
/* out3 (device) = { ||u-uh||_L2^2, |||u-uh|||_DG^2, |u-uh|_H1^2 }; the caller takes the square roots (main.py:608-610). */
int reyna_b200_errors(const rb_geom *g, const rb_quad *q, const rb_err_samples *s, double sigma_D,
                      const double *solution, double *out3, void *workspace, size_t workspace_bytes, void *stream);

/* Point evaluation of the DG solution (replaces tools.evaluate, reyna/DGFEM/two_dimensional/tools.py:117-153), batched:
 * out[i] = u_h(xy[i]) on element elem[i] (the points must lie in / on their elements, as upstream requires).
 * bbox [n_elem][4], xy [n_points][2], elem [n_points], solution [n_elem*d], out [n_points]: device pointers. */
int reyna_b200_evaluate(int p, int32_t n_points, const double *bbox, const double *xy, const int32_t *elem,
                        const double *solution, double *out, void *stream);

/* FP64 FMA micro-benchmark used as the compute roofline denominator: runs `iters` dependent-chain FMAs on every
 * lane of a full grid and returns the flop count through *flops (host). */
int reyna_b200_fma_probe(int iters, double *scratch_dev, double *flops, void *stream);
/* The same for the FP64 tensor-core path (mma.sync m8n8k4 f64, 8 independent accumulator tiles per warp): the measured DMMA
 * rate that DESIGN.md section 4 weighs against the 6x6 blocks of the assembly (an 8x8 tile is 56 % full). */
int reyna_b200_dmma_probe(int iters, double *scratch_dev, double *flops, void *stream);

/* Number of CUDA kernels this library has launched in the calling process (monotone; bench.py reports the delta). */
unsigned long long reyna_b200_launch_count(void);

const char *reyna_b200_last_error(void);
const char *reyna_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* REYNA_B200_H */
