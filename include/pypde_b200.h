/*
 * pypde_b200 — C ABI of the sm_100a kernels behind pypde's Chebyshev
 * spectral-Galerkin time-step hot path.
 *
 * This is the drop-in boundary: every entry point replaces one call the
 * reference makes into its f2py Fortran modules / scipy.fftpack / scipy.sparse
 * (file:line given per function, relative to the reference root).
 *
 * Conventions
 *   - all data are float64 DEVICE pointers owned by the caller (no ownership
 *     transfer); 2-D arrays are row-major (n0, n1) with a leading dimension
 *     `ld*` counted in elements (>= n1); 1-D data is a (n, 1) array;
 *   - `axis` = 0: the operator acts along axis 0 (one independent problem per
 *     column), `axis` = 1: along axis 1 (one problem per row); `batch` is the
 *     extent of the other axis;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - every function returns 0 on success; on failure a non-zero code, and
 *     pde_last_error() holds the message (thread-local);
 *   - small coefficient tables (diagonals, stencils) are DEVICE pointers too,
 *     uploaded once by the host-side plan objects; opaque plan handles own
 *     only their constant tables.
 *   - nothing here falls back to the CPU: without a CUDA device every compute
 *     entry point fails with PDE_ERR_CUDA.
 */
#ifndef PYPDE_B200_H
#define PYPDE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PDE_OK 0
#define PDE_ERR_ARG 1
#define PDE_ERR_CUDA 2
#define PDE_ERR_UNSUPPORTED 3

/* ---- library ---------------------------------------------------------------- */
const char *pde_last_error(void);
int pde_version(void);
/* SM count and compute capability of the current device. */
int pde_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* Number of kernel launches issued by this library since the last reset
 * (bench.py's `gpu_launches`). */
long pde_launch_count(void);
void pde_launch_count_reset(void);

/* ---- DCT-I -------------------------------------------------------------------
 * Replaces scipy.fftpack.dctn(f, type=1, axes=(0,)) (pypde/bases/chebyshev.py:85-93)
 * and, with mode 1/2, the scale/sign/mass passes wrapped around it in
 * Chebyshev.forward_fft (:67-76, :143-149) and Chebyshev.backward_fft (:78-83).
 *   PDE_DCT_RAW : y_k = x_0 + (-1)^k x_{L-1} + 2 sum_{n=1}^{L-2} x_n cos(pi k n/(L-1))
 *   PDE_DCT_FWD : physical values on the Gauss-Lobatto grid -> Chebyshev coefficients
 *   PDE_DCT_BWD : Chebyshev coefficients -> physical values
 * Input entries n_in..L-1 along the axis are taken as zero (zero_pad,
 * pypde/bases/utils.py:89-110); only the first n_out outputs are written
 * (zero_unpad, :113-115).
 * algo: 0 = auto, 1 = dense cosine matrix (fp64 tensor-core GEMM),
 *       2 = shared-memory FFT (L-1 even with factors 2,3,5 only), 3 = Bluestein. */
#define PDE_DCT_RAW 0
#define PDE_DCT_FWD 1
#define PDE_DCT_BWD 2
typedef struct pde_dct_plan_s *pde_dct_plan_t;
int pde_dct_plan_create(pde_dct_plan_t *plan, int L, int algo);
int pde_dct_plan_destroy(pde_dct_plan_t plan);
int pde_dct_plan_algo(pde_dct_plan_t plan);
int pde_dct1(pde_dct_plan_t plan, int mode, const double *x, long ldx, int n_in,
             double *y, long ldy, int n_out, int batch, int axis, void *stream);

/* ---- Galerkin stencil maps ----------------------------------------------------
 * to_cheb: u = S v with S[k,k] = 1, S[k+2,k] = s[k]
 *   (GalerkinChebyshev.to_chebyshev, chebyshev.py:287-293; stencils :382-392, :426-436).
 *   v has M entries along the axis, u gets n_out >= 1 entries (entries beyond
 *   M+1 are zero: Galerkin-space zero padding for dealiasing, field.py:48-51).
 * from_cheb: v = (S^T S)^-1 S^T u, the S^T product followed by the offset-2
 *   tridiagonal solve (chebyshev.py:295-337 -> tdma.f90:55-106).  Tables:
 *   s[M], a[M-2] = sub-diagonal of S^T S, den[M], w[M-2] = Thomas denominators
 *   and c/den ratios, computed on the host exactly as tdma.f90:82-89 does. */
int pde_to_cheb(const double *s, const double *v, long ldv, int M,
                double *u, long ldu, int n_out, int batch, int axis, void *stream);
int pde_from_cheb(const double *s, const double *a, const double *den, const double *w,
                  const double *u, long ldu, int M, double *v, long ldv,
                  int batch, int axis, void *stream);
/* tdma alone: solve_tdma_1d/2d(a,b,c,d,k=2) (bases/linalg/tdma.py:92-100). */
int pde_tdma2_solve(const double *a, const double *den, const double *w,
                    const double *d, long ldd, int n, double *x, long ldx,
                    int batch, int axis, void *stream);

/* ---- Chebyshev derivative recurrence -----------------------------------------
 * differentiate_cheby.diff_1d/diff_2d (bases/fortran/differentiate_cheby.f90:1-53)
 * applied `order` times, then divided by `div` (grad(): `dvhat /= scale**deriv`,
 * field_operations.py:40-45; pass 1.0 for none).  c and dc are n long along the
 * axis, dc[n-1] = 0.  c and dc must not alias. */
int pde_cheb_diff(const double *c, long ldc, double *dc, long lddc, int n, int batch,
                  int axis, int order, double div, void *stream);

/* ---- banded matrix product ----------------------------------------------------
 * y = A x (accumulate = 0) or y += A x (accumulate = 1) along the axis for an
 * (n_out x n_in) matrix with `ndiag` diagonals: diags[d*n_out + r] = A[r, r+offsets[d]]
 * (PlanRHS.solve with the banded B, B@S matrices: solver/plans.py:54-74,
 * matrix.py:48-53).  offsets is a HOST array, ascending; products are summed
 * in ascending column order like the CSR mat-vec. */
int pde_banded_mul(const double *diags, const int *offsets, int ndiag,
                   const double *x, long ldx, int n_in, double *y, long ldy, int n_out,
                   int batch, int axis, int accumulate, void *stream);

/* ---- banded solves (in place on x) -------------------------------------------
 * fdma:   4 diagonals at offsets -2,0,2,4, pre-factored l,d,u1,u2 (Plan_fdma.FDMA_LU,
 *         solver/plans.py:226-234) -> solve_fdma_1d/2d (solver/linalg/fortran/fdma.f90:1-98).
 * twodma: diagonals 0,+2 -> solve_twodma_1d/2d (twodma.f90:1-61). */
int pde_fdma_solve(const double *l, const double *d, const double *u1, const double *u2,
                   double *x, long ldx, int n, int batch, int axis, void *stream);
int pde_twodma_solve(const double *d, const double *u, double *x, long ldx, int n, int batch,
                     int axis, void *stream);

/* ---- eigen-diagonalised Poisson core -------------------------------------------
 * solve_fdma_type2(A, C, lam, x, axis=0, singular) (fdma.f90:146-195 with
 * init_fdma :102-143): for column i solve (A + lam_i C) x_i = b_i.
 * The plan factors every column ONCE on the device (same operation order as
 * init_fdma) and keeps the l,d,u1,u2 tables (4 n m doubles); the reference
 * re-factors on every call.  Adiag/Cdiag: HOST arrays [4][n] holding the
 * diagonals at offsets -2,0,2,4 (entry r = M[r, r+off], 0 outside); lam: HOST [m]. */
typedef struct pde_poisson_plan_s *pde_poisson_plan_t;
int pde_poisson_plan_create(pde_poisson_plan_t *plan, const double *Adiag, const double *Cdiag,
                            const double *lam, int n, int m, int singular);
int pde_poisson_plan_destroy(pde_poisson_plan_t plan);
int pde_poisson_solve(pde_poisson_plan_t plan, double *x, long ldx, void *stream);
/* Copy one factor table of the plan into a caller-owned DEVICE buffer: which = 0..4 -> l, d, u1, u2, RN(1/d)
 * ((n x m) doubles, row r = entry of original row r, column = system), 5 -> the m ints that flag the
 * columns of the singular branch (row/column 0 dropped).  Used to lay the tables out for pde_pass_run. */
int pde_poisson_plan_export(pde_poisson_plan_t plan, int which, void *dst);

/* ---- dense fp64 contraction -----------------------------------------------------
 * C(m x n) = A(m x k) * B, B given as (k x n) [transB = 0] or (n x k) [transB = 1];
 * PlanRHS / PlanLHS "multiply" with the dense Hy, Qy of the Poisson plan
 * (templates/poisson.py:93-108): axis 0 -> C = Mat * X, axis 1 -> C = X * Mat^T. */
int pde_gemm_f64(int transB, const double *A, long lda, const double *B, long ldb,
                 double *C, long ldc, int m, int n, int k, void *stream);


/* ---- batched entry points -----------------------------------------------------------
 * The time stepper applies the same operator to several arrays at once (the U, V, T
 * fields; value and derivative of one field; ...).  These entry points take up to
 * PDE_MAX_JOBS arrays per launch (blockIdx.y / .z = job), which multiplies the number of
 * resident warps of the latency-bound sweeps and divides the number of launches.
 * Job arrays are HOST arrays of plain structs (copied into kernel parameters). */
#define PDE_MAX_JOBS 8

/* One sequence sweep job (pde_sweep).  Element i of sequence q is in[s][i*ldin[s] + q]
 * (axis 0) or in[s][q*ldin[s] + i] (axis 1); the result goes to out the same way
 * (out may alias in[0]).  tab[] are per-index DEVICE tables, see the op list. */
typedef struct {
    const double *in[5];
    long ldin[5];
    double *out;
    long ldout;
    const double *tab[6];
    const int *itab;
    int nseq;
    int flag;
    double sc;
} pde_sweep_job;

/* ops of pde_sweep (reference recurrences: see pde_cheb_diff, pde_from_cheb, pde_fdma_solve,
 * pde_twodma_solve, pde_poisson_solve; every solve is a forward and a backward sweep):
 *   DIFF        in[0] = c                 flag/sc: divide the result by sc
 *   TDMA_FWD    in[0] = u (in[1] = u too when tab[0] = stencil s is given: rhs = u_i + s_i u_{i+2});
 *               tab: 1 = a, 2 = den, 4 = RN(1/den) (optional, same bits, shorter critical path)
 *   TDMA_BWD    in[0] = g (usually == out), tab[3] = w
 *   FDMA_FWD    tab[0] = l;   FDMA_BWD  tab: 1 = d, 2 = u1, 3 = u2, 4 = RN(1/d) (optional)
 *   TWODMA_BWD  tab: 0 = d, 1 = u, 4 = RN(1/d) (optional)
 *   POISSON_*   driven by pde_poisson_solve (per-column tables as extra streams)
 * Kernel choice (results are bit-identical): axis-1 sweeps of even length >= 128 whose rows, pitches and
 * tables are 16-byte aligned run tile by tile (32 rows per CTA, coalesced 512-byte row segments); TDMA_BWD /
 * FDMA_FWD leave their first / last two indices unwritten, so the tiled form needs out == in[0] for them;
 * DIFF with a power-of-two sc multiplies by RN(1/sc).  Everything else takes the generic row-per-lane kernel. */
#define PDE_SWEEP_DIFF 0
#define PDE_SWEEP_TDMA_FWD 1
#define PDE_SWEEP_TDMA_BWD 2
#define PDE_SWEEP_FDMA_FWD 3
#define PDE_SWEEP_FDMA_BWD 4
#define PDE_SWEEP_TWODMA_BWD 5
int pde_sweep(int op, int axis, int n, int njobs, const pde_sweep_job *jobs, void *stream);

/* u = S v (pde_to_cheb) for several arrays */
typedef struct {
    const double *s;
    const double *v;
    long ldv;
    int M;
    double *u;
    long ldu;
    int n_out;
    int batch;
} pde_stencil_job;
int pde_to_cheb_multi(int axis, int njobs, const pde_stencil_job *jobs, void *stream);

/* y = A x or y += A x (pde_banded_mul) for several arrays */
typedef struct {
    const double *diags;
    int ndiag;
    int off[8];
    const double *x;
    long ldx;
    int n_in;
    double *y;
    long ldy;
    int n_out;
    int batch;
    int accumulate;
} pde_band_job;
int pde_banded_multi(int axis, int njobs, const pde_band_job *jobs, void *stream);

/* y = (((c0 x0) + c1 x1) + c2 x2) + c3 x3, rounded after every product and sum like the
 * NumPy expressions of navier/rbc2d.py:252-394 (a coefficient of exactly 1 is not multiplied);
 * y may alias any x.  All arrays are (n0 x n1) row-major with their own leading dimension. */
typedef struct {
    int nterm;
    const double *x[4];
    long ldx[4];
    double coef[4];
    double *y;
    long ldy;
    int n0, n1;
} pde_lincomb_job;
int pde_lincomb_multi(int njobs, const pde_lincomb_job *jobs, void *stream);

/* Pseudo-spectral products of one IMEX stage on the (dealiased) physical grid, n points
 * (conv_term / convective_term, pypde/field_operations.py:83-169, with the two calls of an
 * RK3 stage, navier/rbc2d.py:260-266, merged by linearity: ub = b u + c u_old):
 *   dxU <- ub dxU + wb dzU,  dxV <- ub dxV + wb dzV,  dxT <- ub dxT + wb dzT + wb dTbc
 * u_old / w_old may be NULL when c == 0; dTbc may be NULL. */
int pde_conv_products(long n, double b, double c, const double *u, const double *w, const double *u_old,
                      const double *w_old, double *dxU, const double *dzU, double *dxV, const double *dzV,
                      double *dxT, const double *dzT, const double *dTbc, void *stream);

/* The same products for `nmembers` independent runs whose arrays lie `stride` elements apart (ensemble of
 * equal grids; dTbc is shared). */
int pde_conv_products_members(long n, int nmembers, long stride, double b, double c, const double *u, const double *w,
                              const double *u_old, const double *w_old, double *dxU, const double *dzU, double *dxV,
                              const double *dzV, double *dxT, const double *dzT, const double *dTbc, void *stream);

/* pde_dct1 on several arrays of identical shape */
int pde_dct1_multi(pde_dct_plan_t plan, int mode, int njobs, const double *const *x, long ldx, int n_in,
                   double *const *y, long ldy, int n_out, int batch, int axis, void *stream);

/* Batched forms for ensembles of small grids (one launch for all members): operands come from DEVICE arrays
 * of `nbatch` pointers.  pde_dct1_batched needs a plan of the dense-matrix kind (pde_dct_plan_algo == 1);
 * pde_gemm_f64_batched takes each of A / B either shared (plain pointer) or per problem (device pointer array).
 * aligned != 0: the caller guarantees 16-byte aligned operands. */
int pde_dct1_batched(pde_dct_plan_t plan, int mode, int nbatch, const double *const *dev_x, long ldx, int n_in,
                     double *const *dev_y, long ldy, int n_out, int batch, int axis, int aligned, void *stream);
int pde_gemm_f64_batched(int transB, const double *A, const double *const *dev_A, long lda, const double *B,
                         const double *const *dev_B, long ldb, double *const *dev_C, long ldc, int m, int n, int k,
                         int nbatch, int aligned, void *stream);

/* ---- slab decomposition: pack / unpack of a bundle for the all-to-all transposes -------
 * bundle : (rows, K*cols) row-major, K arrays side by side (array k = columns k*cols .. k*cols+cols-1)
 * blocked: for every rank s (columns col_off[s] .. col_off[s+1]-1) a contiguous block (rows, K, w_s)
 * dir = 1: bundle -> blocked (pack before sending in the Y -> X transpose),
 * dir = 0: blocked -> bundle (unpack after receiving in the X -> Y transpose).
 * col_off is a HOST array of nranks+1 offsets (nranks <= 16). */
int pde_slab_repack(int dir, double *bundle, double *blocked, int rows, int K, int cols, int nranks,
                    const int *col_off, void *stream);


/* ---- fused axis passes ---------------------------------------------------------------
 * One launch applies a chain of 1-D operators (a small program) to every sequence of several
 * 2-D arrays: the axis passes of SURVEY.md §8(d).  Replaces, fused, the per-operator calls the
 * reference makes inside one stage of NavierStokes.update (navier/rbc2d.py:396-434):
 * to_chebyshev / from_chebyshev (bases/chebyshev.py:287-337 -> tdma.f90:55-106), the derivative
 * recurrence (differentiate_cheby.f90:28-53), PlanRHS banded products (solver/plans.py:54-74),
 * Plan_fdma solves (fdma.f90:1-98), Plan_Poisson column solves (fdma.f90:146-195) and the NumPy
 * axpy's between them (rbc2d.py:252-394).
 *
 * layout = axis the operators act along: PDE_PASS_COL (0): sequence q = column q, element i at
 * p + (i - start) * ld + q; PDE_PASS_ROW (1): sequence q = row q, element i at p + q * ld + (i - start).
 * An operand may be split along the sequence into nseg segments (segment s holds elements
 * start[s] .. start[s+1]-1 behind its own base pointer p[s] / leading dimension ld[s]): with peer
 * mappings of the other ranks' slabs as bases, the loads and stores of the row passes ARE the
 * distributed transposes of the slab decomposition.  ROW operands need even starts / leading
 * dimensions and 16-byte aligned bases.
 *
 * A sequence lives in shared memory as 16-byte units (x[2m], x[2m+1]); a launch works on
 * NUP = 32 << lg_segu units (sequences up to 2 NUP elements, zero padded).  Instructions:
 *   TABLES  stage n <= 4 recurrence tables p[k] (segment order, NUP double2 each) in the shared-memory slots
 *           k = 0..n-1; executed once per thread block and job
 *   LOAD    buffer <- operand[0..n), zero beyond
 *   STORE   operand[0..n) <- buffer           (flag ONLY_SEQ: only the sequence with global index off[0];
 *           flag BULK, ROW layout: cp.async.bulk shared -> global copies of 16 SEGU bytes, one per lane)
 *   AXPY    buffer <- buffer + f0 operand[0..n)   (flag SCALED: f1 buffer + f0 operand;
 *           flag STENCIL: the image operand_i + st_i operand_{i-2}, i < n + 2, of the n operand entries,
 *           with the element table st = p[7])
 *   LINCOMB buffer <- [buffer +] sum_k coef[k] G_k for nseg <= PDE_PASS_MAX_TERMS unsegmented operands
 *           p[k] / ld[k] with start[k] valid elements each (flag ACCUM keeps the buffer)
 *   SCALE   buffer <- f0 buffer
 *   SETZ0   element 0 of the sequence with global index off[0] <- 0
 *   POINT   y[m] = sum_{t<n} C_t[m] x[m + off[t]] on units, C_t = p[t] (double2 per unit, NUP entries)
 *           or 1 when p[t] is NULL; offsets in [-1, 2], all >= 0 or all <= 0
 *   DIFF    Chebyshev derivative recurrence, result times f0
 *   REC1    y[m] = T0[m] b[m] - T1[m] y[m-1]  (flag DESC: y[m+1]); T0 optional (= 1)
 *   REC2    x[m] = T0[m] b[m] - T1[m] x[m+1] - T2[m] x[m+2]
 * REC tables: table k is the shared-memory slot slot[k] (>= 0, staged by TABLES) or, with slot[k] < 0, the
 * global pointer p[k] (NULL = absent); segment order: the entry of unit lane*SEGU + j at [j*32 + lane]
 * (SEGU = 1 << lg_segu); flag PERSEQ (global tables): the tables of sequence q start ld[k] doubles after
 * those of sequence q-1.
 * Programs (<= 16 instructions) and job descriptors are DEVICE arrays (built once by the host-side stepper). */
#define PDE_PASS_COL 0
#define PDE_PASS_ROW 1
#define PDE_PASS_MAX_SEG 8
#define PDE_PASS_MAX_TERMS 6
#define PDE_PASS_LOAD 1
#define PDE_PASS_STORE 2
#define PDE_PASS_AXPY 3
#define PDE_PASS_SCALE 4
#define PDE_PASS_SETZ0 5
#define PDE_PASS_POINT 6
#define PDE_PASS_DIFF 7
#define PDE_PASS_REC1 8
#define PDE_PASS_REC2 9
#define PDE_PASS_TABLES 10
#define PDE_PASS_LINCOMB 11
#define PDE_PASS_F_DESC 1
#define PDE_PASS_F_PERSEQ 2
#define PDE_PASS_F_SCALED 4
#define PDE_PASS_F_STENCIL 8
#define PDE_PASS_F_ONLY_SEQ 16
#define PDE_PASS_F_ACCUM 32
#define PDE_PASS_F_BULK 64      /* ROW STORE: hand the row's 32 segments to the TMA bulk-copy engine (cp.async.bulk) */
typedef struct {
    int op;
    int n;
    int flags;
    int nseg;
    double f0, f1;
    double coef[PDE_PASS_MAX_TERMS];
    const void *p[PDE_PASS_MAX_SEG];
    long ld[PDE_PASS_MAX_SEG];
    int start[PDE_PASS_MAX_SEG + 1];
    int off[4];
    int slot[3];
} pde_pass_ins;
typedef struct {
    const pde_pass_ins *prog;   /* DEVICE pointer */
    int nins;
    int nseq;                   /* sequences of this job */
    int seq0;                   /* global index of sequence 0 (slab decomposition) */
    int pad_;
} pde_pass_job;
int pde_pass_run(int layout, int lg_segu, int njobs, int max_nseq, const pde_pass_job *dev_jobs, void *stream);
/* sequences per thread block (the strip width of the COL layout) for a given lg_segu */
int pde_pass_width(int lg_segu);

/* ---- peer memory of the slab decomposition (one process per GPU, one node) ---------------
 * Exchange buffers are cudaMalloc'ed by the library so that their IPC handles (64 bytes) can be
 * handed to the other ranks (through torch.distributed); pde_ipc_open maps a peer's buffer.
 * pde_peer_barrier is the device-side rendezvous between the ranks that separates the passes:
 * dev_peer_flags[s] = mapping of rank s's flag array (nranks 64-bit slots), dev_epoch = this
 * rank's epoch counter, dev_err is set when a peer did not arrive within ~4 s. */
int pde_ipc_alloc(void **ptr, long bytes, void *handle64);
int pde_ipc_open(const void *handle64, void **ptr);
int pde_ipc_close(void *ptr);
int pde_ipc_free(void *ptr);
int pde_peer_barrier(void *const *dev_peer_flags, void *dev_epoch, int rank, int nranks, int *dev_err, void *stream);

/* ---- layout helper ------------------------------------------------------------- */
/* out(n1 x n0) = in(n0 x n1)^T */
int pde_transpose(const double *in, long ldin, double *out, long ldout, int n0, int n1, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYPDE_B200_H */
