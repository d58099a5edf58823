/* cashocs_b200 -- C-ABI of the B200-native shape-gradient hot path.
 *
 * Drop-in boundary for the per-iteration shape-gradient pipeline of sblauth/cashocs
 * (SURVEY.md section 8b).  Plain pointers and sizes only; no torch / dolfin / PETSc types.
 * Every function returns 0 on success and a negative CB_ERR_* code on failure
 * (cb_last_error() gives the message).  Krylov outcomes are reported separately as
 * PETSc-compatible `reason` values so the Python shim can raise PETScKSPError(reason)
 * exactly like cashocs/_utils/linalg.py:543-544.
 *
 * Ownership: the CALLER allocates and frees every device buffer (torch CUDA tensors in
 * the Python host layer); the library allocates only stream-ordered scratch that it
 * frees before returning, plus the opaque cb_dist communicator object.
 * Threading: not re-entrant per stream; all work is stream-ordered on the cudaStream_t
 * passed as `stream` (void*), with a host synchronisation only where a scalar result is
 * returned to the host (documented per function).
 *
 * Index types: "ptr" arrays (rowptr, v2c_ptr, colour offsets) are int64; column / dof /
 * cell-to-nnz indices are int32.  All floating point data is IEEE fp64.
 * Matrices are block-CSR with square blocks of size bs in {1,2,3}, row-major inside a
 * block (bs == 1 is plain CSR; layout == PETSc AIJ / BAIJ), columns sorted ascending.
 */
#ifndef CASHOCS_B200_H
#define CASHOCS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK 0
#define CB_ERR_CUDA -1
#define CB_ERR_ARG -2
#define CB_ERR_CAPACITY -3
#define CB_ERR_NCCL -4
#define CB_ERR_UNSUPPORTED -5

/* PETSc KSPConvergedReason values (cashocs/_exceptions.py:93-129). */
#define CB_KSP_CONVERGED_RTOL 2
#define CB_KSP_CONVERGED_ATOL 3
#define CB_KSP_CONVERGED_ITS 4
#define CB_KSP_DIVERGED_ITS -3
#define CB_KSP_DIVERGED_DTOL -4
#define CB_KSP_DIVERGED_BREAKDOWN -5
#define CB_KSP_DIVERGED_INDEFINITE_PC -8
#define CB_KSP_DIVERGED_NANORINF -9
#define CB_KSP_DIVERGED_INDEFINITE_MAT -10

const char* cb_last_error(void);
int cb_version(void);
/* sizeof() of a header struct inside the compiled library: 0 cb_poly, 1 cb_quad, 2 cb_mat, 3 cb_mesh_view,
 * 4 cb_ksp_opts, 5 cb_amg_level, 6 cb_amg, 7 cb_ksp_result, 8 cb_p2_view; -1 otherwise.  Bindings compare their
 * mirrors with it at load time. */
int64_t cb_struct_size(int which);
/* Number of SMs of the current device (grid sizing); <0 on error. */
int cb_num_sms(void);
/* Total number of CUDA kernels this library has launched in this process (monotonic). */
long long cb_launch_count(void);

/* ------------------------------------------------------------------ polynomial data
 * A source term f(x) is passed as a polynomial in the spatial coordinates (the reference
 * JIT-compiles the UFL expression through FFC; cashocs/_forms/general_form_handler.py:101-140).
 */
#define CB_POLY_MAXT 24
typedef struct {
  int32_t nterms;
  int32_t pad;
  double coef[CB_POLY_MAXT];
  int8_t ex[CB_POLY_MAXT];
  int8_t ey[CB_POLY_MAXT];
  int8_t ez[CB_POLY_MAXT];
} cb_poly;

/* Quadrature rule on the reference simplex in barycentric coordinates; weights sum to 1. */
#define CB_QUAD_MAXQ 64
typedef struct {
  int32_t nq;
  int32_t pad;
  double lam[CB_QUAD_MAXQ][4];
  double w[CB_QUAD_MAXQ];
} cb_quad;

/* ------------------------------------------------------------------ sparsity pattern
 * Replaces dolfin's dofmap-graph sparsity (fenics.assemble_system with
 * bilinear_boundary_form_modification, cashocs/_utils/forms.py:274-288,
 * cashocs/_utils/linalg.py:149-156): pattern = all (test dof, trial dof) pairs per cell.
 */
/* vertex(dof)-to-cell adjacency; v2c_ptr [nd+1], v2c_idx [nc*k] (cells ascending per dof). */
int cb_v2c_build(void* stream, int64_t nd, int64_t nc, int k, const int32_t* cell_dofs,
                 int64_t* v2c_ptr, int32_t* v2c_idx);
/* rowptr [nrows+1] of the dofmap graph restricted to rows < nrows (owned rows). */
int cb_pattern_rowptr(void* stream, int64_t nrows, int k, const int32_t* cell_dofs,
                      const int64_t* v2c_ptr, const int32_t* v2c_idx, int64_t* rowptr);
/* col [nnz] sorted ascending per row, cell2nnz [nc*k*k] (-1 where the row is not owned). */
int cb_pattern_fill(void* stream, int64_t nrows, int64_t nc, int k, const int32_t* cell_dofs,
                    const int64_t* v2c_ptr, const int32_t* v2c_idx, const int64_t* rowptr,
                    int32_t* col, int32_t* cell2nnz);
/* Deterministic greedy colouring on the HOST (cells sharing a dof get different colours).
 * cell_dofs_host [nc*k], colour_host [nc] out; returns the number of colours (<0 on error). */
int cb_colour_cells_host(int64_t nd, int64_t nc, int k, const int32_t* cell_dofs_host,
                         int32_t* colour_host);

/* ------------------------------------------------------------------ assembly
 * Cells are passed colour-sorted: colour c owns cells [colour_off[c], colour_off[c+1]) of
 * `cells` / `cell2nnz` (colour_off is a HOST array of ncolours+1 int64).  bc_flag [ndof]
 * (uint8, nullable) marks Dirichlet dofs, bc_val [ndof] their values; elimination is
 * symmetric per element tensor like dolfin's SystemAssembler (assemble_system,
 * cashocs/_utils/linalg.py:149-156).  val / rhs may be NULL to skip that tensor; they
 * are zeroed by the call.  nrows_owned: rows >= nrows_owned (ghost dofs) are skipped.
 */
typedef struct {
  int32_t dim;            /* 2 or 3 */
  int32_t ncolours;
  const int64_t* colour_off; /* host */
  const int32_t* cells;      /* [nc, dim+1] colour-sorted, vertex ids (== scalar P1 dofs) */
  const int32_t* cell2nnz;   /* [nc, (dim+1)^2] colour-sorted */
  const double* coords;      /* [nv, dim] */
  int64_t nv;                /* vertices incl. ghosts */
  int64_t nrows_owned;       /* owned vertices (rows assembled) */
  int64_t nnz;               /* block nnz */
  /* gather (nnz-centric) assembly data, nullable: for every block-nnz entry q the list
   * n2c_idx[n2c_ptr[q] .. n2c_ptr[q+1]) of contributions t = cell * (dim+1)^2 + a*(dim+1) + b in
   * ascending t (cell ids of the ORIGINAL cell order `cells_orig`).  When present the matrix
   * kernels run one thread per nnz entry: no colours, no read-modify-write, fixed summation order. */
  const int32_t* cells_orig; /* [nc, dim+1] original order */
  const int64_t* n2c_ptr;    /* [nnz+1] */
  const int32_t* n2c_idx;    /* [sum of valid cell2nnz entries] */
  /* gather assembly of vectors, nullable: the cell kernel stores each cell's local vector in
   * cell_scratch ([nc * (dim+1) * dim] doubles, original cell order, one launch over all cells),
   * then one thread per owned vertex sums its incident cells (v2c lists, ascending cell id). */
  const int64_t* v2c_ptr;    /* [nv+1] */
  const int32_t* v2c_idx;    /* [nc*(dim+1)] */
  double* cell_scratch;
  double* cell_geo;          /* [nc * 16] per-cell geometry cache of the gather matrix kernels (scratch) */
  /* row-centric matrix assembly, nullable: the CSR pattern itself (cb_pattern_rowptr / cb_pattern_fill).  When present
   * (together with v2c_* and cells_orig) a group of lanes owns one row: the cells around its vertex are staged once in
   * shared memory and every entry of the row sums the cells that contain its column vertex, ascending cell id. */
  const int64_t* rowptr;     /* [nrows_owned+1] */
  const int32_t* col;        /* [nnz] */
} cb_mesh_view;

/* Scalar P1:  A = c_stiff*K + c_mass*M ;  rhs = s_poly*int(f phi) + s_nodal*M fnodal.
 * Replaces assemble_petsc_system (cashocs/_utils/linalg.py:107-188) for the state,
 * adjoint, Lame-mu and mass-matrix systems. */
int cb_assemble_scalar(void* stream, const cb_mesh_view* m, double c_stiff, double c_mass,
                       double s_poly, const cb_poly* f, const cb_quad* q, double s_nodal,
                       const double* fnodal, const uint8_t* bc_flag, const double* bc_val,
                       double* val, double* rhs);

/* Vector P1 elasticity Riesz scalar product (cashocs/_forms/shape_form_handler.py:655-685):
 * 2 mu eps(U):eps(W) + lambda div U div W + delta U.W, mu a P1 nodal field, optional per-cell
 * scaling 1/vol^gamma (cell_scale [nc], nullable; colour-sorted for the scatter path, original
 * cell order when the view carries the gather data).  Block-CSR bs = dim.
 * bc_flag / bc_val are per vector dof [dim*nv] (vertex-blocked dof = dim*v + c). */
int cb_assemble_elasticity(void* stream, const cb_mesh_view* m, const double* mu,
                           double lambda_lame, double damping, const double* cell_scale,
                           const uint8_t* bc_flag, double* val);

/* Shape derivative vector of the Poisson Lagrangian with J = int u dx
 * (cashocs/_forms/shape_form_handler.py:533-549; analytic form
 * tests/test_shape_optimization.py:252-259):
 * dJ[V] = int u divV + divV gu.gp - (gradV + gradV^T) gu.gp - divV f p - (gradf.V) p.
 * gradf: dim polynomials.  BC rows are zeroed (SystemAssembler lifting with g = 0). */
int cb_assemble_shape_derivative_poisson(void* stream, const cb_mesh_view* m, const double* u,
                                         const double* p, const cb_poly* f, const cb_poly* gradf,
                                         const cb_quad* q, double c_int_u,
                                         const uint8_t* bc_flag, double* rhs);
/* The same for the whole supported Lagrangian family
 *   L = int c_int_u u + c_track/2 (u - u_d)^2 + kappa grad u . grad p + reaction u p - f p dx
 * with P1 u, p and a P1 (vertex-valued) desired state u_d (nullable when c_track = 0):
 *   dL[V] = int (density of L) divV - kappa (gradV + gradV^T) gu.gp - (gradf.V) p
 *           - [pull_back] c_track (u - u_d) (grad u_d . V)
 * -- the last term is the pull-back the reference adds for coefficients that are Functions
 * (cashocs/_forms/shape_form_handler.py:507-531, `[ShapeGradient] use_pull_back`); analytic forms:
 * tests/test_shape_optimization.py:252-259, :284-291. */
int cb_assemble_shape_derivative(void* stream, const cb_mesh_view* m, const double* u, const double* p,
                                 const cb_poly* f, const cb_poly* gradf, const cb_quad* q, double c_int_u,
                                 double kappa, double reaction, double c_track, const double* u_d,
                                 int pull_back, const uint8_t* bc_flag, double* rhs);

/* Cubic term c u^3 p of a semilinear state equation, P1 u: what cashocs.newton_solve
 * (cashocs/nonlinear_solvers/newton_solver.py:285-362) re-assembles per Newton step for the nonlinearity of
 * tests/test_nonlinear_solvers.py:28-61 and tests/test_optimal_control.py:364-376.
 *   val (nullable, scalar CSR of the mesh pattern): val += 3 c int u^2 phi_a phi_b  (the Jacobian of the term; entries
 *       in Dirichlet rows / columns -- bc_flag, nullable -- are left alone: the linear part made them identity rows)
 *   vec (nullable) [nrows_owned]: vec = c int u^3 phi_a  (overwritten; the caller zeroes Dirichlet rows)
 * q: a rule exact for degree 4.  Needs the gather data of the view (cells_orig, n2c, v2c, cell_scratch). */
int cb_assemble_cubic(void* stream, const cb_mesh_view* m, const double* u, double c, const cb_quad* q,
                      const uint8_t* bc_flag, double* val, double* vec);

/* PETScMatrix.ident_zeros (cashocs/_utils/linalg.py:167): rows with |row|_1 < 3e-16 get a
 * unit diagonal.  n_fixed (device int64, nullable) receives the number of rows fixed. */
int cb_ident_zeros(void* stream, int64_t nrows, int bs, const int64_t* rowptr,
                   const int32_t* col, double* val, int64_t* n_fixed);

/* ------------------------------------------------------------------ P2 (quadratic Lagrange) spaces
 * State / adjoint spaces FunctionSpace(mesh, "CG", 2) of BASELINE configs[4]; the deformation space
 * stays vector CG1 (cashocs/_optimization/shape_optimization/shape_optimization_problem.py:251-256).
 * dofs: vertex v -> v, edge e -> nv + e; local order: the dim+1 vertices, then the edges
 * (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) (2-D: (0,1),(0,2),(1,2)).  The pattern (rowptr/col) and the
 * dof->cell adjacency come from cb_pattern_* / cb_v2c_build with k = ndof_e.  Reference tensors of the
 * affine element (relative to |K|), computed once by the host layer:
 *   ref_R[a][b][k][l] = int d(phi_a)/d(lam_k) d(phi_b)/d(lam_l),  K_ab = |K| sum_kl ref_R G_k.G_l
 *   ref_M[a][b] = int phi_a phi_b,  ref_I[a] = int phi_a. */
typedef struct {
  int32_t dim;               /* 2 or 3 */
  int32_t ndof_e;            /* 6 or 10 */
  int64_t nc, nv, ndofs, nnz;
  int64_t nrows_owned;       /* owned dofs = matrix rows (== ndofs on one GPU; owned dofs come first) */
  int64_t nv_owned;          /* owned vertices (rows of the vector-P1 shape derivative) */
  int64_t nc_owned;          /* owned cells (cost functional) */
  const int32_t* cells;      /* [nc, dim+1] vertex ids */
  const int32_t* cell_dofs;  /* [nc, ndof_e] */
  const double* coords;      /* [nv, dim] */
  const int64_t* rowptr;     /* [ndofs+1] */
  const int32_t* col;        /* [nnz] */
  const int64_t* d2c_ptr;    /* [ndofs+1] dof -> cells (ascending) */
  const int32_t* d2c_idx;    /* [nc*ndof_e] */
  const double* ref_R;       /* [ndof_e, ndof_e, dim+1, dim+1] */
  const double* ref_M;       /* [ndof_e, ndof_e] */
  const double* ref_I;       /* [ndof_e] */
  double* cell_scratch;      /* [nc * max(ndof_e, (dim+1)*dim)] local vectors of the cell kernels */
} cb_p2_view;

/* A = c_stiff*K + c_mass*M on the P2 space, symmetric Dirichlet elimination per element tensor
 * (bc_flag [ndofs], nullable) -- assemble_petsc_system (cashocs/_utils/linalg.py:107-188). */
int cb_p2_assemble_matrix(void* stream, const cb_p2_view* m, double c_stiff, double c_mass,
                          const uint8_t* bc_flag, double* val);
/* rhs = s_poly*int(f phi) + s_const*int(phi) + s_nodal*M fnodal, with the BC lifting of the matrix
 * c_stiff*K + c_mass*M.  qtab [nq][dim+1 + 1 + ndof_e]: barycentric point, weight (sum 1), basis
 * values; the rule must be exact for degree(f) + 2. */
int cb_p2_assemble_rhs(void* stream, const cb_p2_view* m, double c_stiff, double c_mass, double s_poly,
                       const cb_poly* f, int nq, const double* qtab, double s_const, double s_nodal,
                       const double* fnodal, const uint8_t* bc_flag, const double* bc_val, double* rhs);
/* Shape derivative of the Poisson Lagrangian (cb_assemble_shape_derivative_poisson) for P2 u, p and
 * vector-P1 test functions; v2c_* = vertex->cell adjacency of the mesh; bc_flag per P1 vector dof;
 * rhs [nv*dim]; the rule must be exact for degree(f) + 2. */
int cb_p2_shape_derivative_poisson(void* stream, const cb_p2_view* m, const int64_t* v2c_ptr,
                                   const int32_t* v2c_idx, const double* u, const double* p,
                                   const cb_poly* f, const cb_poly* gradf, int nq, const double* qtab,
                                   double c_int_u, const uint8_t* bc_flag, double* rhs);
/* int (c_u*u + c_track/2 (u-ud)^2) dx for P2 u, ud (IntegralFunctional.evaluate). */
int cb_p2_functional(void* stream, const cb_p2_view* m, double c_u, const double* u, double c_track,
                     const double* ud, double* result_dev, double* scratch, int64_t scratch_len);

/* ------------------------------------------------------------------ SpMV / BLAS-1 */
typedef struct {
  int64_t nrows;          /* owned block rows */
  int64_t ncols;          /* block columns incl. ghosts */
  int32_t bs;             /* 1, 2 or 3 */
  int32_t row_split;      /* SpMV hint: lanes groups per row, 0/1, 2, 4 or 8 (set ~ avg row length / 16) */
  const int64_t* rowptr;  /* [nrows+1] */
  const int32_t* col;     /* [nnzb] */
  const double* val;      /* [nnzb*bs*bs] */
  const float* val32;     /* nullable: fp32 copy of val (cb_to_f32).  Read ONLY by preconditioner passes (AMG smoothers,
                           * residuals and the A*P post-smoothing inside cb_ksp_solve): half the bytes per pass, products
                           * and sums stay fp64.  The Krylov method's own operator is always `val`. */
  int32_t cta_warps;      /* launch hint, 0 = default: warps per CTA of the A*P post-smoothing pass.  A CTA streams the
                           * values of (rows per warp) x (warps) consecutive rows as ONE bulk copy; operators with short
                           * rows (A*P: ~9 blocks per row against ~15 of A) need more rows per CTA to keep the copies
                           * -- and with them the bytes in flight per SM -- as large as for A. */
  int32_t reserved;
} cb_mat;

/* y = A x  (PETSc Mat.mult; cashocs/_forms/shape_form_handler.py:776-777). */
int cb_spmv(void* stream, const cb_mat* A, const double* x, double* y);
/* y = A32 x with the fp32 copy of the values (A->val32 must be set): the pass the AMG V-cycle runs.  x, y, products
 * and sums are fp64, so the result equals cb_spmv on the rounded matrix bit for bit. */
int cb_spmv_f32(void* stream, const cb_mat* A, const double* x, double* y);
/* One step of the power method on D^-1 A (AMG numeric phase: lambda_max for the Chebyshev smoothers and the prolongator
 * damping): y = alpha * dinv .* (A x) and *norm2_dev = |y|^2 over this rank's rows, one fused pass (use32 != 0: over
 * A->val32 when present). */
int cb_power_step(void* stream, const cb_mat* A, const double* dinv, double alpha, const double* x, double* y,
                  double* norm2_dev, double* scratch, int64_t scratch_len, int use32);
/* dst = (float) src, element-wise (fills cb_mat.val32). */
int cb_to_f32(void* stream, int64_t n, const double* src, float* dst);
/* result (device double) = y^T A x without materialising A x
 * (ShapeFormHandler.scalar_product, cashocs/_forms/shape_form_handler.py:748-780). */
int cb_scalar_product(void* stream, const cb_mat* A, const double* x, const double* y,
                      double* result_dev, double* scratch, int64_t scratch_len);
/* Deterministic dot product, result on device. scratch >= cb_reduce_scratch_len() doubles. */
int cb_dot(void* stream, int64_t n, const double* x, const double* y, double* result_dev,
           double* scratch, int64_t scratch_len);
int64_t cb_reduce_scratch_len(void);
/* dinv = 1 / diag(A) per scalar row [nrows*bs]; also row-wise Gershgorin sum_j|a_ij|/a_ii max
 * into lmax_dev (device double, nullable). */
int cb_extract_diag_inv(void* stream, const cb_mat* A, double* dinv, double* lmax_dev,
                        double* scratch, int64_t scratch_len);

/* ------------------------------------------------------------------ Krylov solvers
 * Replaces LinearSolver.solve (cashocs/_utils/linalg.py:480-546): zero initial guess,
 * left preconditioning, preconditioned residual norm, KSPConvergedDefault. */
#define CB_KSP_CG 0
#define CB_KSP_MINRES 1
#define CB_KSP_PREONLY 2
#define CB_PC_NONE 0
#define CB_PC_JACOBI 1
#define CB_PC_CHEBYSHEV 2
#define CB_PC_AMG 3

typedef struct cb_dist cb_dist; /* opaque multi-GPU context (communicator + halo plan), NULL for one GPU */

/* Smoothed-aggregation AMG hierarchy (device counterpart of `pc_type hypre|gamg`, the reference's
 * iterative default cashocs/_utils/linalg.py:44-52).  All arrays are caller-owned device memory;
 * the symbolic data (patterns) is built once per sparsity pattern by the host layer, the values by
 * cb_amg_smooth_prolongator / cb_amg_galerkin for every new matrix.  P = P_s (x) I_bs: a scalar
 * CSR prolongator applied to every component, so each level keeps the block size of the fine
 * matrix.  Level l < nlevels-1 owns the prolongator from level l+1. */
typedef struct {
  cb_mat A;                 /* level operator */
  const double* dinv;       /* [A.nrows*bs] inverse diagonal (Jacobi / Chebyshev smoother) */
  double lmax;              /* upper bound of the spectrum of D^-1 A (Gershgorin) */
  int64_t nc;               /* number of coarse nodes (0 on the last level) */
  const int64_t* p_rowptr;  /* P  : [A.nrows] rows, nc columns */
  const int32_t* p_col;
  const double* p_val;
  const int64_t* pt_rowptr; /* P^T: [nc] rows */
  const int32_t* pt_col;
  const double* pt_val;
  double* x;                /* [A.ncols*bs] level iterate   (unused on level 0: caller's z) */
  double* b;                /* [A.nrows*bs] level rhs       (unused on level 0: caller's r) */
  double* t;                /* [A.nrows*bs] A*x scratch */
  double* d;                /* [A.nrows*bs] Chebyshev direction */
  cb_dist* dist;            /* halo plan of this level (NULL on one GPU) */
  double* sendbuf;          /* [send count * bs] pack buffer of this level's halo exchange */
  cb_dist* rep_dist;        /* non-NULL: first *replicated* level of a distributed hierarchy -- A is the
                             * global operator on every rank, b is all-reduced (sum) over this
                             * communicator after the restriction wrote this rank's rows */
  int64_t rep_off;          /* first node of this rank's rows inside the replicated level */
  cb_mat AP;                /* A*P of this level (rows: fine nodes, columns: owned + ghost coarse nodes), a
                             * by-product of the Galerkin product; val == NULL disables its use.  The first
                             * post-smoothing step computes its residual as b - A x0 - (AP) e instead of
                             * b - A (x0 + P e): one pass over AP (~1/3 of the bytes of A) */
  const int32_t* coarse_gid;/* [AP.ncols] position of every local coarse node in the replicated next level
                             * (only when the next level has rep_dist) */
  double* xc_local;         /* [AP.ncols*bs] gather buffer for that case */
  /* One-sided all-gather of the restricted residual of a replicated level (replaces memset + NCCL all-reduce: aggregates
   * never cross ranks, so every coarse row is produced by exactly one rank).  rep_gather: halo plan whose "ghosts" are
   * all other ranks' rows, in rank order; rep_tmp [A.nrows*bs]: this rank's rows first, then the ghosts; rep_lo: number
   * of rows owned by lower ranks (== rep_off).  NULL: the all-reduce path. */
  cb_dist* rep_gather;
  double* rep_tmp;
} cb_amg_level;
typedef struct {
  int32_t nlevels;
  int32_t smooth_degree;    /* Chebyshev degree of the pre / post smoother */
  int32_t coarse_degree;    /* Chebyshev degree of the coarsest-level solve (coarse_inv == NULL) */
  int32_t p_block;          /* 1: prolongator entries are 3 x 3 blocks (rigid-body hierarchy, cb_amg_rbm_*): p_val holds
                             * the blocks of P, pt_val the TRANSPOSED blocks in P^T's entry order; 0: scalar P (x) I */
  double smooth_ratio;      /* smoother interval [lmax/ratio, lmax] */
  double coarse_ratio;
  int32_t level_smooth_degree; /* > 0: Chebyshev degree of the smoothers of the levels >= 1 (0: smooth_degree) */
  int32_t pad3;
  const cb_amg_level* levels; /* host array */
  const double* coarse_inv;   /* dense inverse of the last level's operator (nullable) */
} cb_amg;

/* ---- Rigid-body modes inside the hierarchy (3-D, 3 x 3 blocks): WORK IN PROGRESS, not yet used by cb_ksp_solve.
 * Every aggregate J is two coarse nodes (2J translations, 2J+1 rotations about its centroid scaled by its radius);
 * layout and symbolic phase: cashocs_b200/_utils/amg_rbm.py; arithmetic checked on the CPU by
 * tests/test_amg_rbm.py.  All value arrays hold row-major 3 x 3 blocks. */
/* C [nc,3] centroid of the translation nodes, S [nc] radius max(|x_i - C| + s_i) of every aggregate. */
int cb_amg_rbm_aggregate_geometry(void* stream, int64_t nc, const int64_t* member_ptr, const int32_t* member_idx,
                                  const uint8_t* kind, const double* centre, const double* scale, double* C, double* S);
/* Tentative prolongator blocks (translation node: I and cross(x - C)/S; rotation node: s/S I). */
int cb_amg_rbm_tentative(void* stream, int64_t n, const int64_t* t_rowptr, const int32_t* agg, const uint8_t* kind,
                         const double* centre, const double* scale, const double* C, const double* S, double* t_val);
/* out[q] = sum over the term segment of q of op(lhs[src_l]) * rhs[src_r]; trans_l: op = transpose (P^T (A P)). */
int cb_amg_rbm_spgemm_terms(void* stream, int trans_l, int64_t nnz_out, const int64_t* seg_ptr, const int32_t* src_l,
                            const int32_t* src_r, const double* lhs, const double* rhs, double* out);
/* p_val holds A*T on entry and P = T - omega D^-1 (A T) on exit (D = diag(A), dinv [3n]). */
int cb_amg_rbm_prolongator_finish(void* stream, int64_t n, const int64_t* p_rowptr, int64_t nt, const int64_t* t2p,
                                  const double* t_val, const double* dinv, double omega, double* p_val);
/* pt_val[k] = transpose(p_val[perm[k]]). */
int cb_amg_rbm_transpose_blocks(void* stream, int64_t nnz, const int64_t* perm, const double* p_val, double* pt_val);
/* bc = P^T (b - t)  and  x += P xc. */
int cb_amg_rbm_restrict(void* stream, int64_t ncn, const int64_t* pt_rowptr, const int32_t* pt_col, const double* pt_val,
                        const double* b, const double* t, double* bc);
int cb_amg_rbm_prolong(void* stream, int64_t n, const int64_t* p_rowptr, const int32_t* p_col, const double* p_val,
                       const double* xc, double* x);

typedef struct {
  int32_t ksp_type;
  int32_t pc_type;
  double rtol, atol, dtol;
  int32_t max_it;
  int32_t cheb_degree;
  double cheb_ratio;      /* lmax / lmin of the Chebyshev interval (default 30) */
  int32_t check_every;    /* host polls the device status every this many iterations */
  int32_t monitor;        /* 1: record the residual history (hist, up to hist_len) */
  int32_t profile_every;  /* >0: bracket every profile_every-th SpMV launch with CUDA events */
  int32_t use_graph;      /* PCG iteration as a CUDA graph (captured after the first eager chunk, replayed once
                           * per iteration; bitwise the same result): 0 = on (default; eager when a partitioned run
                           * had to fall back from the one-sided transport to NCCL send/recv), 1 = always, -1 = never */
  const cb_amg* amg;      /* hierarchy for pc_type CB_PC_AMG (NULL otherwise) */
  /* Rigid-body coarse correction for vector problems (bs = 2: 3 modes, bs = 3: 6 modes): the AMG preconditioner
   * M becomes M + Z (Z^T A Z)^-1 Z^T with Z the global translations and rotations -- the rotations are
   * low-energy modes of the elasticity operator that the scalar prolongator does not interpolate.
   * rbm_coords [nrows, bs]: coordinates of the owned nodes relative to a centre common to all ranks (NULL: off);
   * rbm_einv (device, k x k, row-major): (Z^T A Z)^-1; rbm_work (device, >= 8 doubles): Z^T r. */
  const double* rbm_coords;
  const double* rbm_einv;
  double* rbm_work;
  int32_t flexible;       /* 1: Polak-Ribiere beta = z_new.(r_new - r_old) / z_old.r_old (PETSc KSPFCG with mmax = 1): equal to
                           * the Fletcher-Reeves beta for a fixed symmetric preconditioner, keeps PCG converging at the unperturbed
                           * rate when the preconditioner is only symmetric to fp32 round-off (mixed-precision AMG) */
  int32_t pad2;
} cb_ksp_opts;
/* Set-up of that correction for one matrix version, all on the device.  cb_rbm_galerkin: c_out [ncols, bs] = coords -
 * centre (centre: HOST array of bs doubles, the same on all ranks; any fixed point spans the same modes), then row j of
 * G_dev [k, k] = Z^T (A z_j) over this rank's rows (k launches of the SpMV; sum G over the ranks before inverting; G is
 * symmetric, so rows = columns).  v_work [ncols*bs] and w_work [nrows*bs] are scratch vectors.  cb_rbm_invert: einv = G^-1
 * (k <= 6, one small kernel).  rbm_coords of cb_ksp_opts is c_out (its first nrows nodes). */
int cb_rbm_galerkin(void* stream, const cb_mat* A, const double* coords, const double* centre, double* c_out,
                    double* v_work, double* w_work, double* G_dev, double* scratch, int64_t scratch_len);
int cb_rbm_invert(void* stream, int k, const double* G_dev, double* einv_dev);
typedef struct {
  int32_t reason;
  int32_t its;
  double rnorm;
  double rnorm0;
  double spmv_ms;         /* summed duration of the profiled SpMV launches (profile_every > 0) */
  int32_t spmv_profiled;  /* number of SpMV launches profiled */
  int32_t spmv_launches;  /* number of SpMV launches enqueued by this solve */
} cb_ksp_result;

typedef struct cb_dist cb_dist; /* opaque multi-GPU context, NULL for one GPU */

/* Work-space size in doubles for cb_ksp_solve on a matrix with n = nrows*bs scalar rows
 * and nc = ncols*bs scalar columns. */
int64_t cb_ksp_workspace_len(int64_t n, int64_t nc, const cb_ksp_opts* opts);
/* Solve A x = b. x [>= ncols*bs] is overwritten (zero initial guess). Host-synchronises on
 * `stream` (the result is returned to the host).  hist (host, nullable) gets rnorm per
 * iteration when opts->monitor. */
int cb_ksp_solve(void* stream, const cb_mat* A, const double* b, double* x,
                 const cb_ksp_opts* opts, cb_dist* dist, double* work, int64_t work_len,
                 cb_ksp_result* result, double* hist, int64_t hist_len);

/* AMG numeric phase.  s_val [nnzb] = trace(A_blk)/bs (the scalar strength matrix S). */
int cb_amg_block_trace(void* stream, const cb_mat* A, double* s_val);
/* P = T - omega D_F^-1 S^F T for the piecewise-constant tentative prolongator T of `agg` [nrows].
 * S^F is S restricted to its local columns (col < nrows) with the dropped ghost couplings lumped into
 * the diagonal (identity on one GPU), so prolongator rows never reference another rank's aggregates
 * and still sum to one.  Rows flagged in `tentative` (uint8 [nrows], nullable) keep the row e_agg. */
int cb_amg_smooth_prolongator(void* stream, const cb_mat* S, double omega, const int32_t* agg,
                              const uint8_t* tentative, const int64_t* p_rowptr, const int32_t* p_col,
                              double* p_val);
/* Numeric sparse product with a precomputed term list (symbolic phase of the host layer):
 * out[q] = sum_{t in [seg_ptr[q], seg_ptr[q+1])} lhs[src_l[t]] * rhs[src_r[t]], where exactly one
 * operand holds bs x bs blocks (lhs_block != 0: lhs, else rhs) and the other is scalar.  Used for
 * A*P and P^T*(A*P) of the Galerkin coarse operators. */
int cb_amg_spgemm_terms(void* stream, int bs, int lhs_block, int64_t nnz_out, const int64_t* seg_ptr,
                        const int32_t* src_l, const int32_t* src_r, const double* lhs,
                        const double* rhs, double* out);
/* Dense inverse of the coarsest operator (<= 256 scalar rows); work [2 n^2], inv [n^2] row-major. */
/* The same product without term lists, for operators whose lists would not fit (8 bytes per scalar
 * product term): out (known pattern o_rowptr / o_col, sorted columns) = L * R with exactly one block-valued
 * operand (lhs_block != 0: L), rows of L = rows of out.  Bitwise the result of cb_amg_spgemm_terms. */
int cb_amg_spgemm_rowwise(void* stream, int bs, int lhs_block, int64_t nrows, const int64_t* l_rowptr,
                          const int32_t* l_col, const double* l_val, const int64_t* r_rowptr,
                          const int32_t* r_col, const double* r_val, const int64_t* o_rowptr,
                          const int32_t* o_col, double* out);
int cb_amg_dense_inverse(void* stream, const cb_mat* A, double* work, double* inv);

/* ------------------------------------------------------------------ mesh kernels */
#define CB_Q_SKEWNESS 0
#define CB_Q_MAXIMUM_ANGLE 1
#define CB_Q_RADIUS_RATIOS 2
#define CB_Q_CONDITION_NUMBER 3
/* Per-cell quality (cashocs/geometry/quality.py:101-271, :371-461) over cells [0, nc_owned);
 * q (nullable) gets the per-cell values; red_dev[0] = min, red_dev[1] = sum (device). */
int cb_mesh_quality(void* stream, int dim, int64_t nc, const double* coords,
                    const int32_t* cells, int measure, double* q, double* red_dev,
                    double* scratch, int64_t scratch_len);
/* det(I + grad V) per cell (APrioriMeshTester.test, cashocs/geometry/mesh_testing.py:84-132):
 * red_dev[0] = min, red_dev[1] = max;  V [nv, dim] vertex field scaled by `step`. */
int cb_det_check(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                 const double* V, double step, double* det, double* red_dev, double* scratch,
                 int64_t scratch_len);
/* max_K sqrt(grad d : grad d) (compute_decreases, cashocs/geometry/mesh_handler.py:329-377). */
int cb_gradient_frobenius_max(void* stream, int dim, int64_t nc, const double* coords,
                              const int32_t* cells, const double* D, double* red_dev,
                              double* scratch, int64_t scratch_len);
/* vol [nc] = |K| per cell and *max_dev = max |K| over this rank's cells: `CellVolume` of the inhomogeneous elasticity
 * scaling (cashocs/_forms/shape_form_handler.py:629-644; its l2_projection to DG0, cashocs/_utils/linalg.py:768-809,
 * is the identity for cellwise constants). */
int cb_cell_volumes(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells, double* vol,
                    double* max_dev, double* scratch, int64_t scratch_len);
/* coords_old = coords; coords += step * V  (DeformationHandler.move_mesh,
 * cashocs/geometry/deformations.py:132-133); n = nv*dim. */
int cb_move_mesh(void* stream, int64_t n, double* coords, double* coords_old, const double* V,
                 double step);
/* int_Omega (c_u * u + c_track/2 (u-ud)^2) dx  (IntegralFunctional.evaluate,
 * cashocs/_optimization/cost_functional.py:160-168); ud nullable; result_dev device double. */
int cb_functional(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                  double c_u, const double* u, double c_track, const double* ud,
                  double* result_dev, double* scratch, int64_t scratch_len);
/* int_Gamma 1 ds over the boundary facets [nf, dim] (SurfaceRegularization._compute_surface,
 * cashocs/_forms/shape_regularization.py:307-316); result_dev device double. */
int cb_boundary_measure(void* stream, int dim, int64_t nf, const double* coords, const int32_t* facets,
                        double* result_dev, double* scratch, int64_t scratch_len);
/* out[v, :] = sum over the boundary facets F of vertex v of d|F|/dx_v, v < nrows -- the vector of
 * int_Gamma t_div(V, n) ds for V = phi_v e_i (shape_regularization.py:63-74,248-269).  v2f_ptr [nrows+1] /
 * v2f_idx: vertex -> facet adjacency (cb_v2c_build with k = dim); facet_vec [nf*dim*dim] scratch. */
int cb_boundary_measure_gradient(void* stream, int dim, int64_t nf, int64_t nrows, const double* coords,
                                 const int32_t* facets, const int64_t* v2f_ptr, const int32_t* v2f_idx,
                                 double* facet_vec, double* out);
/* Mass matrix of the boundary, int_Gamma phi_a phi_b ds (CurvatureRegularization.a_curvature per component,
 * cashocs/_forms/shape_regularization.py:536-542,596-601), on the compressed numbering of the nb boundary
 * vertices: bvert [nb] compressed -> mesh vertex, facets_c [nf, dim] compressed ids, v2f_* their vertex -> facet
 * adjacency, rowptr / col the CSR pattern of the facet graph (cb_pattern_* with the facets as cells); val [nnz]. */
int cb_boundary_mass(void* stream, int dim, int64_t nb, const double* coords, const int32_t* bvert,
                     const int32_t* facets_c, const int64_t* v2f_ptr, const int32_t* v2f_idx,
                     const int64_t* rowptr, const int32_t* col, double* val);
/* out[v, :] = vector of int_Gamma inner((I - (P + P^T)) t_grad(V, n), t_grad(kappa, n)) + 1/2 t_div(V, n)
 * t_div(kappa, n) ds for V = phi_v e_i, P = t_grad(x, n)  (CurvatureRegularization.compute_shape_derivative,
 * shape_regularization.py:554-575; the factor mu is applied by the caller); kappa [nv*dim] the mean-curvature
 * vector; other arguments as cb_boundary_measure_gradient. */
int cb_boundary_curvature_derivative(void* stream, int dim, int64_t nf, int64_t nrows, const double* coords,
                                     const int32_t* facets, const double* kappa, const int64_t* v2f_ptr,
                                     const int32_t* v2f_idx, double* facet_vec, double* out);
/* Right-hand side of one sweep of the eikonal iteration, out[v] = int (grad u / |grad u|) . grad phi_v dx for the
 * owned vertices v < nrows, 0 where bc_flag (uint8, nullable) is set  (compute_boundary_distance_eikonal,
 * cashocs/geometry/boundary_distance.py:262); v2c_* the vertex -> cell adjacency. */
int cb_eikonal_rhs(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                   const int32_t* cells, const double* coords, const double* u, const uint8_t* bc_flag,
                   double* out);
/* int (|grad u| - 1)^2 dx over the nc (owned) cells, the stopping criterion of the same iteration
 * (boundary_distance.py:264-279); result_dev device double. */
int cb_eikonal_residual(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                        const double* u, double* result_dev, double* scratch, int64_t scratch_len);
/* mu[i] = the reference's mu_expression at dist[i]: mu_min below dist_min, mu_max above dist_max, the linear
 * (smooth = 0) or cubic (smooth = 1) blend in between (Stiffness._setup_mu_computation,
 * cashocs/_forms/shape_form_handler.py:141-186). */
int cb_distance_mu(void* stream, int64_t n, const double* dist, double dist_min, double dist_max,
                   double mu_min, double mu_max, int smooth, double* mu);
/* p-Laplace projection of the shape derivative (_PLaplaceProjector,
 * cashocs/_pde_problems/shape_gradient_problem.py:329-427).  Residual of
 * int mu (eps + |grad U|^(p-2)) grad U : grad W + delta U . W dx - shift[W] for the vector-P1 field U [nv*dim] and
 * the owned vertices (rows) < nrows; mu [nv] the P1 Lame parameter, shift [nrows*dim] the assembled shape derivative
 * (nullable), bc_flag (uint8 per dof, nullable): Dirichlet rows are 0.  v2c_*: vertex -> cell adjacency. */
int cb_plaplace_residual(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                         const int32_t* cells, const double* coords, const double* U, const double* mu, int p,
                         double eps, double delta, const double* shift, const uint8_t* bc_flag, double* out);
/* Jacobian of the same form at U (fenics.derivative of the nonlinear form, nonlinear_solvers/snes.py:121-131) as
 * dim x dim blocks on the P1 pattern rowptr / col, symmetric Dirichlet elimination like dolfin's SystemAssembler;
 * val [nnz*dim*dim]. */
int cb_plaplace_jacobian(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                         const int32_t* cells, const double* coords, const double* U, const double* mu, int p,
                         double eps, double delta, const uint8_t* bc_flag, const int64_t* rowptr, const int32_t* col,
                         double* val);
/* int mu (eps + |grad a|^(p-2)) grad a : grad b + delta a . b dx over the nc (owned) cells: the scalar product of
 * deformations while the p-Laplace projection is on (ShapeFormHandler.scalar_product,
 * cashocs/_forms/shape_form_handler.py:761-770,785-815); a, b [nv*dim]; result_dev device double. */
int cb_plaplace_product(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                        const double* a, const double* b, const double* mu, int p, double eps, double delta,
                        double* result_dev, double* scratch, int64_t scratch_len);
/* opp[f] = vertex of the cell behind boundary facet f that does not lie on the facet (orients the outward facet
 * normal, fenics.FacetNormal); v2c_*: vertex -> cell adjacency of the mesh. */
int cb_boundary_facet_opposite(void* stream, int dim, int64_t nf, const int32_t* facets, const int64_t* v2c_ptr,
                               const int32_t* v2c_idx, const int32_t* cells, int32_t* opp);
/* Right-hand sides of the two L2 projections of the re-extension mode `normal`
 * (_compute_normal_deformation, cashocs/_pde_problems/shape_gradient_problem.py:262-297) over the facets [nf, dim]
 * with the outward unit normal n: vector_in != 0: in [nv*dim] = g, out[v] = int (g . n) phi_v ds; vector_in == 0:
 * in [nv] = phi, out[v, i] = int phi n_i phi_v ds; v < nrows.  v2f_*: vertex -> facet adjacency of these facets,
 * facet_vec [nf*dim*dim] scratch. */
int cb_boundary_normal_rhs(void* stream, int dim, int vector_in, int64_t nf, int64_t nrows, const double* coords,
                           const int32_t* facets, const int32_t* opp, const double* in, const int64_t* v2f_ptr,
                           const int32_t* v2f_idx, double* facet_vec, double* out);
/* Per-vertex collision counts (compute_collisions, cashocs/geometry/mesh_testing.py:201-241)
 * with a uniform-grid broad phase built on the device; mismatch_dev (device int64) receives the
 * number of vertices whose count differs from `valence`. */
int cb_intersection_test(void* stream, int dim, int64_t nv, int64_t nc, const double* coords,
                         const int32_t* cells, const int32_t* valence, int32_t* counts,
                         int64_t* mismatch_dev);

/* ------------------------------------------------------------------ multi-GPU (NCCL) */
/* 128-byte NCCL unique id (rank 0 creates it, the host layer broadcasts it). */
int cb_dist_unique_id(char* id128);
/* Create the communicator + halo plan.  nneigh neighbours; for neighbour j: rank
 * neigh_rank[j], send_idx (device int32, scalar-vertex indices into the owned range,
 * send_off[j]..send_off[j+1]) and a contiguous ghost range recv_off[j]..recv_off[j+1]
 * relative to the first ghost vertex. */
int cb_dist_create(cb_dist** out, const char* id128, int rank, int nranks, int nneigh,
                   const int32_t* neigh_rank, const int64_t* send_off, const int32_t* send_idx_dev,
                   const int64_t* recv_off, int64_t n_owned, int64_t n_ghost);
/* A second halo plan (e.g. an AMG coarse level) on the communicator of `parent`. */
int cb_dist_clone_plan(cb_dist* parent, cb_dist** out, int nneigh, const int32_t* neigh_rank,
                       const int64_t* send_off, const int32_t* send_idx_dev, const int64_t* recv_off,
                       int64_t n_owned, int64_t n_ghost);
int cb_dist_destroy(cb_dist* d);
/* One-sided transport over NVLink peer memory (optional; NCCL send/recv + all-reduce otherwise).
 * cb_dist_p2p_export allocates this plan's window (landing zone for ghost values, reduction slots,
 * flags) and returns its 64-byte CUDA IPC handle; the host layer all-gathers the handles and the
 * ghost offsets, cb_dist_p2p_open maps the peers' windows.  Afterwards cb_dist_halo_exchange is one
 * put kernel (pack + store into the neighbours' windows + release flag) and one get kernel (acquire
 * + copy behind the owned part), and all-reduces of <= 8 doubles are put + fixed-order local sum
 * (bit-identical on every rank).  cb_dist_p2p_error: 1 after a peer timed out in a wait. */
int cb_dist_p2p_export(cb_dist* d, char* handle64);
int64_t cb_dist_p2p_ghost_cap(cb_dist* d); /* parity stride (doubles) of this rank's landing zone */
int cb_dist_p2p_open(cb_dist* d, const char* handles, const int64_t* remote_off, const int64_t* remote_cap);
int cb_dist_p2p_disable(cb_dist* d);
int cb_dist_p2p_error(cb_dist* d);
/* Ghost update of a vertex-blocked vector x [(n_owned+n_ghost)*bs] (Vec.ghostUpdate). */
int cb_dist_halo_exchange(void* stream, cb_dist* d, double* x, int bs, double* sendbuf);
/* In-place allreduce of n doubles (op: 0 sum, 1 min, 2 max). */
int cb_dist_allreduce(void* stream, cb_dist* d, double* buf_dev, int n, int op);

#ifdef __cplusplus
}
#endif
#endif /* CASHOCS_B200_H */
