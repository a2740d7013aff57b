/*
 * sulcusfem.h -- C ABI of libsulcusfem.so, the B200 (sm_100a) finite-element solve path behind
 * the reference's solvers.py / analysis.py entry points.
 *
 * The reference (jesstunn/fenics-eff-uptake) has no FFI of its own: every entry below replaces a
 * piece of work that the reference hands to dolfin 2019.1 (C++/FFC/PETSc) through a Python call.
 * The "replaces" notes cite the reference call site (file:line under /root/reference) whose native
 * work the entry point takes over.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors are used as buffers by
 *     the Python host) unless the parameter name starts with h_ (host pointer);
 *   - all floating point data is IEEE double, all index data int32;
 *   - every call takes a cudaStream_t as `void* stream` and only enqueues work on it (the Krylov
 *     drivers synchronise that stream when they poll convergence);
 *   - return value 0 = success, negative = error; sfem_last_error() returns a message for the
 *     calling thread;
 *   - handles own their internal work space, allocated at *_create and released at *_destroy.
 */
#ifndef SULCUSFEM_H
#define SULCUSFEM_H

#ifdef __cplusplus
extern "C" {
#endif

#define SFEM_OK 0
#define SFEM_ERR_CUDA (-1)
#define SFEM_ERR_ARG (-2)
#define SFEM_ERR_NOCONV (-3)

const char* sfem_last_error(void);
int sfem_version(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int sfem_device_sms(void);
/* kernels launched by this library since the last reset (bench.py `gpu_launches`) */
long long sfem_launch_count(void);
void sfem_launch_count_reset(void);

/* Per-launch timing for bench.py's roofline: between start and stop every instrumented launch is
 * bracketed by CUDA events on its stream.  stop() synchronises the device and returns the number of
 * records copied: category (0 spmv, 1 spmv+dot, 2 chebyshev step, 3 residual+d0, 4 element kernels,
 * 5 gather, 6 vector ops, 7 other, 8 staged spmv, 9 peer-memory halo exchange / vector all-reduce), algorithmic bytes,
 * milliseconds. */
int sfem_profile_start(int max_records);
int sfem_profile_stop(int cap, int* h_cat, double* h_bytes, float* h_ms);

/* ------------------------------------------------------------------ sparse mat-vec ---------- */
/* y = A x (mode 0), y = b - A x (mode 1), y += A x (mode 2).  FP64 values, int32 indices.
 * replaces: PETSc MatMult inside solve(a == L, ...)  solvers.py:55,84,151,213,298
 * (sub-warp-per-row vector kernel, grid-stride, grid = multiple of the SM count) */
int sfem_spmv_csr_f64(int nrows, int nnz, const int* rowptr, const int* cols, const double* vals,
                      const double* x, const double* b, double* y, int mode, void* stream);
/* Same for nb = 1 or 2 right-hand sides stored interleaved (x: [ncols][nb], y, b: [nrows][nb]); the
 * matrix stream is read once for both.  Rectangular matrices allowed (multigrid transfers, B, B^T). */
int sfem_spmv_csr_f64_nb(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                         const double* x, const double* b, double* y, int mode, int nb, void* stream);
/* TMA-staged engine.  A tile plan is attached to a matrix by the DEVICE address of its rowptr array;
 * once registered, every SpMV-family launch on that matrix (this call, the multigrid smoother, the
 * Krylov drivers) streams the matrix HBM -> shared memory with bulk async copies and reduces rows
 * from shared memory.  rowptr / cols / vals must be 16-byte aligned and readable >= 8 elements past
 * their logical end.
 *   sfem_staged_plan      HOST helper: greedy tiles of consecutive rows with <= cap_nnz entries and
 *                         <= max_rows rows (64: 128 consumer threads per CTA, 128: 256); h_tile_row has room for nrows+1 ints; returns ntiles or -1
 *   sfem_staged_register  tile_row: DEVICE copy of the plan, kept alive by the caller until unregister
 *   sfem_spmv_csr_f64_staged  like sfem_spmv_csr_f64_nb but fails if the matrix has no usable plan */
int sfem_staged_plan(int nrows, const int* h_rowptr, int cap_nnz, int max_rows, int* h_tile_row);
int sfem_staged_register(const int* rowptr, int nrows, const int* tile_row, int ntiles, int cap_nnz, int max_rows);
void sfem_staged_unregister(const int* rowptr);
/* matrices with fewer tiles than min_tiles use the vector engine (<= 0: default = SM count); returns the old value */
int sfem_staged_set_min_tiles(int min_tiles);
int sfem_spmv_csr_f64_staged(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                             const double* x, const double* b, double* y, int mode, int nb, void* stream);

/* Sliced-ELL (SELL-32-sigma) mirror.  Finite-element rows are short and nearly uniform, so the solver keeps a
 * second copy of every operator it iterates on in a sliced ELLPACK layout: rows sorted by length inside windows
 * of sigma rows, slices of 32 rows stored column-major and padded to their longest row --
 *     entry j of the row held by lane l of slice s  lives at  slice_ptr[s] + 32*j + l.
 * One lane owns one row, so the matrix stream is read in contiguous 256-byte (values) / 128-byte (indices)
 * spans.  The mirror is attached to the CSR by the DEVICE address of its rowptr; it follows the CSR value array
 * `csr_vals`: every entry of this library that writes CSR values marks the mirror dirty, and a dirty mirror is
 * re-packed (svals[i] = csr_vals[src[i]]) at the next solver entry / SpMV -- it is never read stale.
 *   slice_ptr [nslices+1]  start of every slice in entries (multiples of 32)
 *   perm      [nslices*32] CSR row of every lane (-1: no row)
 *   scols, src [padded]    column index / CSR slot of every position (-1 on padding)
 *   svals     [padded]     mirror values (written by the library)
 *   parts     [nparts+1]   first slice of each of nparts = sfem_sell_parts() contiguous parts holding equal numbers
 *                          of entries; a warp of the SpMV kernel walks a run of consecutive parts, i.e. one
 *                          contiguous span of the mirror, in fixed-size chunks that ignore slice boundaries
 * All arrays are DEVICE arrays owned by the caller and kept alive until unregister (sulcusfem/device.py builds
 * them once per pattern on the host).
 *   sfem_sell_mark_dirty   for callers that write the CSR values themselves
 *   sfem_sell_sync         re-pack all dirty mirrors now
 *   sfem_sell_set_min_rows matrices with fewer rows keep the lane-group engine (<= 0: default); returns old value
 *   sfem_spmv_csr_f64_sell like sfem_spmv_csr_f64_nb but fails if the matrix has no usable mirror
 * replaces: PETSc MatMult (as above); the layout is new -- nothing like it exists in the reference. */
int sfem_sell_parts(void);
int sfem_sell_register(const int* rowptr, const double* csr_vals, int nrows, int nslices, const int* slice_ptr,
                       const int* perm, const int* scols, const int* src, double* svals, long long padded,
                       const int* parts, int nparts);
void sfem_sell_unregister(const int* rowptr);
int sfem_sell_mark_dirty(const int* rowptr);
int sfem_sell_sync(void* stream);
int sfem_sell_set_min_rows(int min_rows);
int sfem_spmv_csr_f64_sell(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                           const double* x, const double* b, double* y, int mode, int nb, void* stream);

/* ------------------------------------------------------------------ assembly ---------------- */
/* Element matrices of  D grad(c).grad(phi) + (u.grad c) phi  on P2 triangles.
 *   geo      [6][nc] SoA vertex coordinates x0,y0,x1,y1,x2,y2 of every cell
 *   celldofs [6][nc] SoA global P2 dofs (only read when ux != NULL)
 *   ux, uy   P2 nodal velocity components (NULL = no advection)
 *   E        [nc][36] row-major element matrices (out)
 * replaces: FFC tabulate_tensor of the cell integrals in solvers.py:43-44,78,140,207 */
int sfem_elem_p2_advdiff(int nc, const double* geo, const int* celldofs, double D,
                         const double* ux, const double* uy, double* E, void* stream);
/* Same on P1 triangles for the multigrid coarse levels (velocity given at vertices; `upwind`
 * adds cell-wise artificial diffusion D*max(1,Pe_h) -- preconditioner only, never the solution
 * operator).  cellverts [3][nc] SoA, E [nc][9]. */
int sfem_elem_p1_advdiff(int nc, const double* geo, const int* cellverts, double D,
                         const double* ux, const double* uy, int upwind, double* E, void* stream);
/* Taylor-Hood element matrices of grad u:grad v - div(v) p - q div(u), cell layout
 * [ux x6, uy x6, p x3]; E [nc][225].   replaces: tabulate_tensor for solvers.py:291-293 */
int sfem_elem_th_stokes(int nc, const double* geo, double* E, void* stream);
/* Divergence blocks of the same form only: EB [nc][3][12], EB[c][k][comp*6 + j] = -int l_k d(phi_j)/dx_comp -- the
 * rows of the cell's three pressure dofs against [u_x x6 | u_y x6], bit-identical to the corresponding entries of
 * sfem_elem_th_stokes.  Together with sfem_elem_p2_advdiff(D = 1, no velocity) for K this feeds the block-form solver
 * directly (72 doubles per cell instead of 225; the full matrix is then only assembled for export / parity).
 * replaces: tabulate_tensor of the  - div(v) p - q div(u)  terms of solvers.py:291-293 */
int sfem_elem_th_div(int nc, const double* geo, double* EB, void* stream);
/* vals[k] = 0 where row_flag[row] or col_flag[col] (either may be NULL): Dirichlet elimination of rectangular blocks.
 * replaces: DirichletBC.apply on the velocity-pressure coupling blocks inside solve(), solvers.py:262-264,298 */
int sfem_csr_zero_flagged(int nrows, const int* rowptr, const int* cols, double* vals, const unsigned char* row_flag,
                          const unsigned char* col_flag, void* stream);
/* P1 mass element matrices (pressure Schur-complement preconditioner); E [nc][9]. */
int sfem_elem_p1_mass(int nc, const double* geo, double* E, void* stream);
/* Robin boundary mass  mu phi_j phi_i ds  on exterior facets.
 *   fgeo   [4][nf] SoA xa,ya,xb,yb;  fdofs [ndof][nf] SoA global dofs (P2: va,vb,edge; P1: va,vb)
 *   mu_nodal  nodal values of mu at the dofs of the space (NULL -> mu_const); P2: the P2
 *             interpolant of the expression, as dolfin builds it for a degree-2 UserExpression
 *   clamp  1 = conditional(ge(mu,0),mu,0) at the quadrature points (solvers.py:204)
 *   F      [nf][9] (P2) or [nf][4] (P1)
 * replaces: tabulate_tensor of  mu*c*phi*ds(4)  solvers.py:48,79,144,208 */
int sfem_facet_p2_robin(int nf, const double* fgeo, const int* fdofs, double mu_const,
                        const double* mu_nodal, int clamp, double* F, void* stream);
int sfem_facet_p1_robin(int nf, const double* fgeo, const int* fdofs, double mu_const,
                        const double* mu_nodal, int clamp, double* F, void* stream);
/* vals[s] = sum_k E[contrib_code[k]], k in [contrib_ptr[s], contrib_ptr[s+1])  -- fixed order, no
 * atomics.   replaces: dolfin Assembler global scatter-add (MatSetValues ADD_VALUES) */
int sfem_gather_csr(int nnz, const int* contrib_ptr, const int* contrib_code, const double* E,
                    double* vals, void* stream);
/* Dirichlet conditions.  bc_flag[n] (uint8) marks constrained dofs, bc_val[n] their values.
 *   mode 0: DirichletBC.apply(A,b) as dolfin does it -- identity rows, b_i=g_i, columns untouched
 *   mode 1: symmetric elimination -- additionally b -= A[:,bc] g and the bc columns are zeroed
 * replaces: DirichletBC::apply inside solve()  solvers.py:30-31,68-69,127-128,188-189,262-264 */
int sfem_apply_dirichlet(int n, int nnz, const int* rowptr, const int* cols, double* vals, double* rhs,
                         const unsigned char* bc_flag, const double* bc_val, int mode, void* stream);
/* dst_vals[i] = src_vals[slot[i]]: copies a sub-block of an assembled CSR (e.g. K, B, B^T of the
 * Taylor-Hood matrix) into its own CSR; the slot map is built once on the host. */
int sfem_csr_extract(int n, const int* slot, const double* src_vals, double* dst_vals, void* stream);
/* blocked (a | b) <-> interleaved (a0,b0,a1,b1,...) layouts of a two-component nodal field */
int sfem_vec_interleave2(int n, const double* a, const double* b, double* out, void* stream);
int sfem_vec_deinterleave2(int n, const double* in, double* a, double* b, void* stream);

/* ------------------------------------------------------------------ vectors ------------------ */
int sfem_vec_axpby(int n, double a, const double* x, double b, double* y, void* stream);   /* y = a x + b y */
int sfem_vec_dot(int n, const double* x, const double* y, double* h_out, void* stream);    /* syncs stream */
int sfem_vec_set(int n, double a, double* x, void* stream);                                /* x = a */
int sfem_vec_pointwise_mul(int n, double a, const double* d, const double* x, double* y, void* stream); /* y = a d.*x */
/* out[i] = flag[i] ? a[i] : (b ? b[i] : 0): Dirichlet values as initial guess (DirichletBC.apply on a vector,
 * solvers.py:30-31) and the Dirichlet rows of a lifted right-hand side; out may alias a or b */
int sfem_vec_select(int n, const unsigned char* flag, const double* a, const double* b, double* out, void* stream);
int sfem_vec_copy(int n, const double* x, double* y, void* stream);                          /* y = x */
/* out (n x n row-major) = inverse of a small CSR matrix (coarsest multigrid level, n <= 2048) */
int sfem_dense_inverse_csr(int n, const int* rowptr, const int* cols, const double* vals, double* out, void* stream);
int sfem_extract_diag_inv(int n, const int* rowptr, const int* cols, const double* vals, double* dinv, void* stream);
/* solvers.py:86-105,154-164,216-224: clamp non-finite to 0, then tiny negatives (|min|<1e-12) to 0;
 * h_stats[6] = {n_nonfinite, n_negative, min, max, mean, clamped(0/1)} (syncs stream) */
int sfem_postprocess_concentration(int n, double* c, int fix_nonfinite, double* h_stats, void* stream);

/* ------------------------------------------------------------------ multigrid ---------------- */
typedef struct sfem_mg* sfem_mg_t;
/* Geometric multigrid V-cycle with Chebyshev-Jacobi smoothing.  Level 0 is the system level.
 * Per level l: operator CSR (A_*[l]); for l < nlevels-1 the prolongation from level l+1
 * (P_*[l], n[l] x n[l+1]) and its transpose (R_*[l]).  coarse_inv: dense row-major inverse of the
 * coarsest operator (n[last]^2).  All arrays are host arrays of device pointers. */
sfem_mg_t sfem_mg_create(int nlevels, const int* h_n, const int* h_A_nnz,
                         const int* const* h_A_rowptr, const int* const* h_A_cols, const double* const* h_A_vals,
                         const int* h_P_nnz,
                         const int* const* h_P_rowptr, const int* const* h_P_cols, const double* const* h_P_vals,
                         const int* const* h_R_rowptr, const int* const* h_R_cols, const double* const* h_R_vals,
                         const double* coarse_inv, int cheb_degree, double eig_ratio, int nb);
/* nb = 1 or 2: number of right-hand sides a V-cycle carries, stored interleaved ([dof][nb]).
 * (re)compute D^-1, the Gershgorin bound of lambda_max(D^-1 A) and the Chebyshev coefficients of
 * every level from the current operator values -- all on the device, no host synchronisation */
int sfem_mg_setup(sfem_mg_t mg, void* stream);
/* set-up of the system level only: mu sweeps that keep the coarse levels of a nearby mu (preconditioner data) */
int sfem_mg_setup_fine(sfem_mg_t mg, void* stream);
int sfem_mg_vcycle(sfem_mg_t mg, const double* b, double* x, void* stream);   /* x = M^-1 b */
/* Multi-GPU: `mg` holds the row-partitioned levels (halo patterns attached to their matrices with
 * sfem_halo_attach BEFORE sfem_mg_create); below its last level the replicated hierarchy `tail` (n_tail
 * dofs on its finest level) is solved redundantly on every rank after a vector all-reduce.
 * P_last: n_own(last) x n_tail; R_last: n_tail x n_own(last), owned columns only. */
int sfem_mg_set_tail(sfem_mg_t mg, sfem_mg_t tail, int n_tail, int P_nnz, const int* P_rowptr, const int* P_cols,
                     const double* P_vals, int R_nnz, const int* R_rowptr, const int* R_cols, const double* R_vals);
/* Experiment knob: multigrid levels with at most `rows` unknowns (and the dense coarsest solve below them) run as ONE
 * fused thread-block-cluster kernel per V-cycle instead of ~7 launch-latency-bound launches per level.  0 = off, the
 * default (environment SFEM_TAIL_ROWS): correct to rounding but measured slower than the separate launches
 * (csrc/sfem_mg_tail.cu, profiles/r02_fused_tail.md).  Returns the previous setting. */
int sfem_mg_set_tail_rows(int rows);
int sfem_mg_lambda_max(sfem_mg_t mg, double* h_out);                            /* per level, host */
void sfem_mg_destroy(sfem_mg_t mg);

/* ------------------------------------------------------------------ Krylov ------------------- */
/* replaces: the sparse LU behind solve(a == L, u, bcs) (dolfin default linear solver).
 * The operator is CSR; the preconditioner is a multigrid handle (NULL = Jacobi).
 * h_info[4] = {iterations, final true relative residual ||b-Ax||/||b||, converged(1/0), estimate} */
int sfem_krylov_cg(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                   const double* b, double* x, double rtol, int maxit, double* h_info, void* stream);
int sfem_krylov_fgmres(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                       const double* b, double* x, double rtol, int restart, int maxit,
                       double* h_info, void* stream);
/* Batched multi-parameter PCG: the Robin-coefficient sweep of ONE geometry in one Krylov loop.
 * Solves (A0 + mu_c M) x_c = b0 + mu_c bM for c = 0..nb-1 (1 <= nb <= 16) with vectors interleaved by column.
 *   rowptr / cols      CSR pattern of the system level (n rows, nnz entries)
 *   lvl_vals0 / lvl_valsM   host arrays [nlevels] of device pointers: two value arrays per multigrid level on that
 *                      level's pattern -- A0_l = D K_l and M_l = M_Gamma,l, Dirichlet rows / columns eliminated
 *                      (identity rows in A0_l, zero rows in M_l); entry 0 is the system level (required); a NULL
 *                      entry l >= 1 makes that level use the handle's operator (assembled for mu_ref) for all columns
 *   h_mu[nb] (host), mu_ref   the coefficients (>= 0) and the coefficient the handle was assembled for
 *   mg                 nb = 1 handle with `nlevels` levels, set up for mu_ref: supplies patterns, transfers and the dense
 *                      inverse of the coarsest A(mu_ref); every column's smoothers run on its OWN operator on every
 *                      level, its coarsest system is solved by a Chebyshev iteration preconditioned with that inverse
 *   b0 / bM            lifted right-hand sides (b_c = b0 + mu_c bM), x0[n] the shared initial guess (Dirichlet values)
 *   X[n][nb]           solutions; every column iterates until ALL reach ||b_c - A_c x_c|| <= rtol ||b_c||
 * h_info[4 * nb]: per column {iterations, true relative residual, converged(1/0), recurrence estimate}.
 * replaces: the serial loop of sparse LU solves  no_advection_analysis_A.py:1306-1347 (run_mu_sweep: 20 mu values on
 * one mesh), no_advection_analysis_B.py:86-200 (mu x geometry), each a pure_diffusion_solver call  solvers.py:113-174 */
int sfem_krylov_cg_batch(int n, int nnz, const int* rowptr, const int* cols, int nlevels, const double* const* lvl_vals0,
                         const double* const* lvl_valsM, int nb, const double* h_mu, double mu_ref, sfem_mg_t mg,
                         const double* b0, const double* bM, const double* x0, double* X, double rtol, int maxit,
                         double* h_info, void* stream);
/* out[i] = X[i][c]: one column of an interleaved batch (the field handed back as a dolfin-style Function) */
int sfem_batch_column(int n, int nb, int c, const double* X, double* out, void* stream);
/* Taylor-Hood Stokes solver: preconditioned MINRES on the block form of the assembled matrix.
 * Unknown / right-hand-side layout: [velocity interleaved (ux0,uy0,ux1,uy1,...) : 2*n2 | pressure : nv].
 *   K   (n2 x n2)      scalar P2 stiffness block with the velocity Dirichlet rows/columns eliminated
 *   B   (nv x 2*n2)    divergence block, columns in interleaved velocity numbering; BT its transpose
 *   Mp  (nv x nv)      P1 pressure mass matrix (Schur-complement preconditioner, 4 Chebyshev steps)
 *   mg                 multigrid handle on K created with nb = 2 (one V-cycle serves both components)
 *   nz, zt_*, zidx, zw, Cc   optional coarse pressure correction  S^-1 += Z Cc Z^T  with Z the nz 1-D
 *                      hat functions in x (zidx/zw: left hat index and weight per pressure dof;
 *                      zt_*: CSR of Z^T), Cc (nz x nz, symmetric positive semi-definite); nz = 0: none
 * All arrays are caller-owned device arrays that must stay valid for the life of the handle; the
 * values may be rewritten between solves (re-assembly).  The iteration is replayed from CUDA graphs.
 * replaces: the sparse LU of solve(a == L, U, bcs) in stokes_solver, solvers.py:298
 * h_info[4] = {iterations, true relative residual, converged(1/0), preconditioned residual estimate} */
typedef struct sfem_stokes* sfem_stokes_t;
sfem_stokes_t sfem_stokes_create(int n2, int nv,
                                 int K_nnz, const int* K_rowptr, const int* K_cols, const double* K_vals,
                                 int B_nnz, const int* B_rowptr, const int* B_cols, const double* B_vals,
                                 const int* BT_rowptr, const int* BT_cols, const double* BT_vals,
                                 int Mp_nnz, const int* Mp_rowptr, const int* Mp_cols, const double* Mp_vals,
                                 sfem_mg_t mg,
                                 int nz, const int* zt_rowptr, const int* zt_cols, const double* zt_vals,
                                 const int* zidx, const double* zw, const double* Cc);
/* Row-partitioned variant for one rank of a multi-GPU solve (BASELINE config 5): n2 / nv = OWNED velocity / pressure
 * dofs; all matrices hold the owned rows, columns in the rank-local numbering [owned | hole | ghosts] planned by
 * sulcusfem/dist.py; n_alloc = length of b, x and every work vector, nv_alloc = length of a pressure work vector.
 * Halo patterns are attached to K, B^T and Mp with sfem_halo_attach; the Krylov scalars are all-reduced in-kernel. */
sfem_stokes_t sfem_stokes_create_part(int n2, int nv,
                                      int K_nnz, const int* K_rowptr, const int* K_cols, const double* K_vals,
                                      int B_nnz, const int* B_rowptr, const int* B_cols, const double* B_vals,
                                      const int* BT_rowptr, const int* BT_cols, const double* BT_vals,
                                      int Mp_nnz, const int* Mp_rowptr, const int* Mp_cols, const double* Mp_vals,
                                      sfem_mg_t mg,
                                      int nz, const int* zt_rowptr, const int* zt_cols, const double* zt_vals,
                                      const int* zidx, const double* zw, const double* Cc,
                                      int BT_nnz, long long n_alloc, long long nv_alloc);
int sfem_stokes_solve(sfem_stokes_t h, const double* b, double* x, double rtol, int maxit, double* h_info, void* stream);
/* Same solve started from a better vector x (e.g. the analytic channel flow of the inlet profile, solvers.py:252-258)
 * with the stopping level anchored to xref, the plain starting vector (Dirichlet values, zero elsewhere):
 * stops when the preconditioned residual is <= rtol * gamma(xref) -- exactly the absolute level sfem_stokes_solve
 * reaches from xref, so the accuracy is unchanged and only iterations are saved.  x and xref in the solver layout. */
int sfem_stokes_solve_from(sfem_stokes_t h, const double* b, double* x, const double* xref, double rtol, int maxit,
                           double* h_info, void* stream);
void sfem_stokes_destroy(sfem_stokes_t h);

/* ------------------------------------------------------------------ multi-GPU (one process per GPU) ---
 * Peer-memory mailboxes over NVLink: every rank allocates one device mailbox, exports it with CUDA IPC
 * (the 64-byte handles are all-gathered by the host, e.g. with torch.distributed) and maps its peers'.
 * Halo exchange and the Krylov all-reduces then run as kernels that store straight into peer memory
 * and synchronise through sequence flags -- no library collective on the data path.
 * replaces: nothing (the reference is serial); SURVEY 8(e): domain-decomposed refined meshes.
 *   mailbox_words   size of the mailbox in 8-byte words: sfem_dist_header_words(nranks, vec_cap) +
 *                   the halo channels laid out by the host (sulcusfem/partition.py)
 *   vec_cap         largest vector all-reduce (restriction onto the replicated coarse hierarchy) */
typedef struct sfem_dist* sfem_dist_t;
typedef struct sfem_halo* sfem_halo_t;
long long sfem_dist_header_words(int nranks, long long vec_cap);
sfem_dist_t sfem_dist_create(int rank, int nranks, long long mailbox_words, long long vec_cap);
int sfem_dist_ipc_handle(sfem_dist_t h, void* out64);
int sfem_dist_open_peers(sfem_dist_t h, const void* handles);
int sfem_dist_set_peer_pointer(sfem_dist_t h, int q, void* mailbox);   /* single-process emulation (tests) */
void* sfem_dist_mailbox(sfem_dist_t h);
int sfem_dist_activate(sfem_dist_t h);
int sfem_dist_error(sfem_dist_t h);
void sfem_dist_destroy(sfem_dist_t h);
/* Halo pattern of one level: HOST arrays of length nneigh (send_idx: send_ptr[nneigh] local owned
 * indices in the order of the receiver's ghost list); offsets in 8-byte words into the peer's / own
 * mailbox; cap = words per parity slot (>= 2 * count). */
sfem_halo_t sfem_halo_create(int nneigh, int n_own, int n_loc, const int* peer, const int* send_ptr, const int* send_idx,
                             const int* recv_cnt, const int* recv_off, const long long* peer_data_off,
                             const long long* peer_flag_off, const long long* my_data_off, const long long* my_flag_off,
                             const long long* cap);
void sfem_halo_destroy(sfem_halo_t h);
/* attach to the matrix whose rowptr lives at this device address: every SpMV-family launch on it first
 * fills the ghost entries of its input vector (vectors hold n_loc entries per right-hand side) */
int sfem_halo_attach(const int* rowptr, sfem_halo_t h);
int sfem_halo_exchange(sfem_halo_t h, double* x, int nb, int phase, void* stream);
int sfem_dist_allreduce_vec(sfem_dist_t h, double* x, int n, int phase, void* stream);
int sfem_dist_allreduce_scalars(sfem_dist_t h, double* vals, int K, void* stream);

/* ------------------------------------------------------------------ functionals -------------- */
/* Facet functionals of analysis.py, all groups in one launch.
 *   entries: ne facet entries sorted by group; ent_cell[ne], ent_local[ne] (local facet 0..2 of
 *   that cell: the cell whose trace and outward normal are used), grp_ptr[ngroups+1]
 *   out [ngroups][8] = { int -D grad c.n, int (u.n) c, int mu c, int c, int 1,
 *                        int |q|, int max(q,0), int max(-q,0) },  q = -D grad c.n + (u.n) c
 * replaces: the ~30 assemble(expr*ds/dS) calls of analysis.py:61-62,212-213,239-247,266-271,311,
 *           323-325,914,918,930-931 */
int sfem_facet_functionals(int ngroups, const int* grp_ptr, const int* ent_cell, const int* ent_local,
                           const double* geo, const int* celldofs, int nc,
                           const double* c, const double* ux, const double* uy,
                           double D, double mu_const, const double* mu_nodal,
                           double* out, void* stream);
/* Cell functionals: out[m][2] = { int_{cells with marker m} c dx, area }, m < nmarkers.
 * replaces: assemble(c*dx(i)), assemble(1*dx(i))  analysis.py:688-694,711-712 */
int sfem_cell_functionals(int nc, const double* geo, const int* celldofs, const int* cell_marker,
                          int nmarkers, const double* c, double* out, void* stream);

/* ------------------------------------------------------------------ point evaluation -------- */
/* Locate npts points and evaluate up to 4 nodal P1 (degree 1) or P2 (degree 2) fields there, one launch.
 *   pts [npts][2];  bin grid: nbx x nby bins of size hx x hy from (x0, y0); bin_ptr [nbx*nby+1], bin_cells =
 *   ascending cells whose bounding box overlaps the bin (sulcusfem/locator.py builds it once per mesh)
 *   geo [6][nc], celldofs [ndof][nc] SoA as in the assembly kernels; h_fields: HOST array of nfields device pointers
 *   tol: a point is inside a cell when min(barycentric) >= -tol; the lowest-numbered such candidate wins
 *   out [nfields][npts] (0 where outside), cell_out [npts] (-1 = in no cell)
 * replaces: mesh.bounding_box_tree().compute_first_entity_collision(Point) + c(Point) / u(Point) per sample,
 *           analysis.py:367-372, 405-410, 574-583, 616-625, 805-810 */
int sfem_eval_points(int degree, int npts, const double* pts, int nbx, int nby, double x0, double y0, double hx, double hy,
                     const int* bin_ptr, const int* bin_cells, const double* geo, int nc, const int* celldofs,
                     int nfields, const double* const* h_fields, double tol, double* out, int* cell_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SULCUSFEM_H */
