/* libtnb200 -- C ABI of the B200-native MPS hot path for TensorNetworks.jl.
 *
 * Every entry point is what a `ccall` from the reference's Julia code would bind for the hot path
 * (INTEGRATION.md shows the Julia side).  Conventions:
 *   - plain C: opaque handles, pointers and sizes; no C++/torch types.
 *   - complex data are interleaved (re, im) doubles, COLUMN-MAJOR, exactly a Julia Array{ComplexF64}
 *     / NumPy complex128 order='F'; `tn_cplx` is layout-compatible with `double _Complex`.
 *   - site / block indices are 1-based like the reference.
 *   - host buffers are borrowed for the duration of the call only; device memory is owned by the library
 *     and stays resident in HBM between calls.
 *   - every function returns 0 on success, a negative status otherwise; tn_last_error() gives the
 *     message (thread-local), which the Julia shim turns into error("...") like the reference does.
 *   - a tn_ctx is bound to one GPU and one CUDA stream and is not re-entrant across threads.
 * All paths below are relative to /root/reference/src.
 */
#ifndef TN_C_API_H
#define TN_C_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } tn_cplx;
typedef struct tn_ctx tn_ctx;
typedef struct tn_mps tn_mps;     /* GMPS of rank 1 (MPS) or 2 (MPO): structures/mps/gmps.jl:8-13 */
typedef struct tn_env tn_env;     /* ProjMPS: structures/mps/projmps.jl:1-8 */
typedef struct tn_gates tn_gates; /* GateList: structures/mps/gatelist.jl:8-13 */
typedef struct tn_envsum tn_envsum; /* ProjMPSSum: structures/mps/projmpssum.jl:1-4 */
typedef struct tn_imps tn_imps;   /* iGMPS of rank 1 in Vidal form: structures/mps/igmps.jl:8-17 */

/* kwargs of svd(): tensors.jl:170-172 (cutoff=0, maxdim=0 meaning unlimited, mindim=1) */
typedef struct { double cutoff; int64_t maxdim; int64_t mindim; } tn_trunc_t;
/* kwargs of KrylovKit.eigsolve as passed at algorithms/mps/dmrg.jl:51-53 */
typedef struct { int32_t krylovdim; int32_t maxiter; double tol; } tn_lanczos_t;

#define TN_OK 0
#define TN_ERR_INVALID (-1)   /* argument / shape error (the reference's error("...")) */
#define TN_ERR_CUDA (-2)
#define TN_ERR_INTERNAL (-3)

const char* tn_last_error(void);
int32_t tn_version(void);

/* ---- context -------------------------------------------------------------------------------- */
int32_t tn_ctx_create(int32_t device, tn_ctx** out);
int32_t tn_ctx_destroy(tn_ctx* ctx);
int32_t tn_sync(tn_ctx* ctx);
/* CUDA stream handle (cudaStream_t) the context launches on, for event timing by the caller */
int32_t tn_ctx_stream(tn_ctx* ctx, void** stream_out);
/* counters: kernels launched by this library (process-wide), matvecs and SVDs done by this ctx */
int32_t tn_counters(tn_ctx* ctx, int64_t* launches, int64_t* matvecs, int64_t* svds);

/* ---- GMPS container: structures/mps/gmps.jl, abstractmps.jl ---------------------------------- */
/* dims: N x (rank+2) row-major table of site tensor sizes; site_ptrs[i]: host tensor of site i+1 */
int32_t tn_mps_upload(tn_ctx* ctx, int32_t rank, int32_t d, int32_t N, const int64_t* dims,
                      const tn_cplx* const* site_ptrs, int32_t center, tn_mps** out);
int32_t tn_mps_free(tn_mps* m);
int32_t tn_mps_info(tn_mps* m, int32_t* rank, int32_t* d, int32_t* N, int32_t* center);
int32_t tn_mps_dims(tn_mps* m, int64_t* dims /* N x (rank+2) */);
int32_t tn_mps_download_site(tn_mps* m, int32_t site, tn_cplx* out);
int32_t tn_mps_upload_site(tn_mps* m, int32_t site, const int64_t* dims, const tn_cplx* data);
int32_t tn_mps_set_center(tn_mps* m, int32_t center);
int32_t tn_mps_maxbonddim(tn_mps* m, int64_t* out);                       /* abstractmps.jl:71-77 */
int32_t tn_mps_norm(tn_mps* m, tn_cplx* out);                             /* gmps.jl:29-37 */
int32_t tn_mps_normalize(tn_mps* m);                                      /* gmps.jl:46-51 */
int32_t tn_mps_movecenter(tn_mps* m, int32_t idx, tn_trunc_t trunc);      /* gmps.jl:90-112 (+ :60-82) */
/* replacesites!(psi, A, site, direction, normalize; kwargs) for a two-site tensor: gmps.jl:199-267 */
int32_t tn_mps_replacesites(tn_mps* m, const tn_cplx* theta_host, int32_t site, int32_t direction,
                            int32_t normalize, tn_trunc_t trunc);
int32_t tn_mps_applyop(tn_mps* m, int32_t site, const tn_cplx* op_host /* d x d */);   /* mps.jl:141-152 */
/* singular values across bond (site, site+1) -- the SVD inside entropy(): gmps.jl:184-189 */
int32_t tn_mps_bond_spectrum(tn_mps* m, int32_t site, double* out, int64_t cap, int64_t* k_out);
/* deepcopy(psi) (abstractmps.jl:95-96) and psi *= a (abstractmps.jl:99-109: the centre tensor, or site 1 when the centre is unset) */
int32_t tn_mps_copy(tn_mps* m, tn_mps** out);
int32_t tn_mps_scale(tn_mps* m, tn_cplx a);
/* applyMPO(O, psi; kwargs...) = O * psi for an MPO and an MPS: mpo.jl:105-143.  Exact site products (bonds w * chi), right-going
 * gauge sweep, then a truncating movecenter!(phi, 1; trunc).  Returns a new MPS handle (centre 1). */
int32_t tn_mpo_apply(tn_mps* O, tn_mps* psi, tn_trunc_t trunc, tn_mps** out);
/* Bond compression of an assembled MPO (or any GMPS): the two truncated-SVD sweeps at the end of MPO(st, H), mpo.jl:443-457
 * (and addMPOs, mpo.jl:296-311) -- right-going O[i] = U S, O[i+1] = V^H O[i+1]; left-going O[i] = S V^H, O[i-1] = O[i-1] U.
 * The reference's default there is cutoff = 1e-15.  Leaves the centre unset. */
int32_t tn_mpo_compress(tn_mps* m, tn_trunc_t trunc);
/* <psi| O_k |psi> for single-site operators: mps.jl:87-134 (one-site terms), qjmc.jl:170-220 */
int32_t tn_expect_local(tn_mps* m, int32_t nops, const int32_t* sites, const tn_cplx* ops_host, tn_cplx* out);

/* ---- truncated SVD: tensors.jl:168-227 -------------------------------------------------------- */
/* mat: m x n column-major on the host.  U: m x k, S: k, Vh: k x n written to the host buffers, which
 * must hold min(m,n) columns / rows (the rank is only known afterwards). */
int32_t tn_svd_trunc(tn_ctx* ctx, const tn_cplx* mat, int64_t m, int64_t n, tn_trunc_t trunc,
                     tn_cplx* U, double* S, tn_cplx* Vh, int64_t* k_out, int32_t* sweeps_out);

/* The split the MPS drivers use (replacesites!, moveleft!/moveright!: gmps.jl:60-82, 218-256): one orthonormal factor (the new
 * site tensor) and the other factor multiplied by S.  side = 1: U (m x k) orthonormal and Vh <- S V^H; side = 2: Vh (k x n)
 * orthonormal and U <- U S.  When the orthonormal factor sits on the long side of the matrix (always for square ones) the right
 * rotations of the Jacobi sweeps are not accumulated ("W-only": half the rotation flops) and the weighted factor is the projection
 * of the input on the orthonormal one.  repeat > 1 re-runs the factorisation on the resident matrix; ms_out (nullable) receives
 * the device time of the last run (factorisation + both gathers, no PCIe). */
int32_t tn_svd_trunc_split(tn_ctx* ctx, const tn_cplx* mat, int64_t m, int64_t n, tn_trunc_t trunc, int32_t side,
                           tn_cplx* U, double* S, tn_cplx* Vh, int64_t* k_out, int32_t* sweeps_out, int32_t repeat, double* ms_out);

/* B truncated SVDs of same-shape matrices in one batched factorisation (the gate / gauge-move SVDs of B QJMC trajectories that
 * advance in lockstep: every kernel launch of the single-problem pipeline covers all B problems).  mats: B consecutive m x n
 * column-major matrices on the host.  Problem b writes U at U + b*m*kmax (m x k_b), S at S + b*kmax, Vh at Vh + b*kmax*n
 * (k_b x n, leading dimension k_b), kmax = min(m, n), and its rank to k_out[b].  Same truncation rule as tn_svd_trunc. */
int32_t tn_svd_trunc_batched(tn_ctx* ctx, int32_t B, const tn_cplx* mats, int64_t m, int64_t n, tn_trunc_t trunc,
                             tn_cplx* U, double* S, tn_cplx* Vh, int64_t* k_out, int32_t* sweeps_out);

/* One Gram + rotation pass of the blocked Jacobi step on caller data (the two kernels of csrc/tn_jacobi.cu, which the factorisations
 * above run between the pair eigen-decompositions; exposed so that they can be checked on their own).  Z: rows x ncols column-major
 * on the host (leading dimension rows), ncols a multiple of 32; pairs: npairs x 2 indices of 32-column blocks (disjoint);
 * J: npairs column-major 64 x 64 matrices; skip: npairs flags or NULL.  G_out[p] = P_p^H P_p (64 x 64, P_p = the 64 columns of pair p,
 * evaluated BEFORE the update), Z_out = Z with the columns of every pair p (skip[p] == 0) replaced by P_p J_p. */
int32_t tn_jacobi_pair_pass(tn_ctx* ctx, const tn_cplx* Z, int64_t rows, int64_t ncols, const int32_t* pairs, int32_t npairs,
                            const tn_cplx* J, const int32_t* skip, tn_cplx* G_out, tn_cplx* Z_out);

/* One trailing update of the block Gram-Schmidt QR inside the factorisations above, on caller data (csrc/tn_jacobi.cu: jacobi_cross64 + the rank-64
 * update form of the rotation kernel).  Q: rows x ncols column-major on the host (ncols a multiple of 64); P = the 64 columns of panel `panel`,
 * T = all columns to its right.  C_out (64 x (ncols - 64 (panel + 1)), leading dimension 64) = P^H T; Q_out = Q with T replaced by T - P C. */
int32_t tn_jacobi_qr_update_pass(tn_ctx* ctx, const tn_cplx* Q, int64_t rows, int64_t ncols, int32_t panel, tn_cplx* C_out, tn_cplx* Q_out);

/* The pair schedule of the Jacobi sweeps for `nblocks` 32-column blocks in `groups` concurrent groups (csrc/tn_svd.cu "Split schedule"), for
 * inspection: (phase, task, step, p, q) per pair into out5 (5 ints per pair, `capacity` pairs); *npairs_out = nblocks (nblocks - 1) / 2, or 0
 * when this block count keeps the circle method.  Tasks of one phase run on separate streams and must own disjoint blocks.  Host function. */
int32_t tn_svd_split_schedule(int32_t nblocks, int32_t groups, int32_t* out5, int64_t capacity, int64_t* npairs_out);

/* The counter-based generator of QJMC throughput runs (uniforms == NULL in tn_qjmc_run / tn_qjmc_ensemble): Philox4x32-10 with key = seed and
 * counter = (trajectory, step, slot); tn_qjmc_uniform is the [0, 1) uniform the drivers draw for (seed, trajectory, step, slot) -- the
 * replacement of `rand()` in qjmc.jl:64,93-112 (Julia's global RNG cannot be reproduced; parity runs pass host-supplied uniforms instead).
 * Pure host functions (no device needed). */
int32_t tn_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4);
int32_t tn_qjmc_uniform(uint64_t seed, uint64_t trajectory, uint64_t step, uint64_t slot, double* out);

/* tuning switch (process-wide): 1 = QR-preconditioned Jacobi (default), 0 = plain Jacobi */
int32_t tn_svd_set_precond(int32_t mode);

/* ---- contraction: tensors.jl:9-18 restricted to the strided-GEMM form the hot path uses ------- */
/* C[m,n] = alpha * sum_k op(A)[m,k] op(B)[k,n]; each of m, n, k is a fused pair of tensor indices
 * (extent n0 x rest) with element strides (s0, s1); conj flags as the reference's conjx/conjy. */
typedef struct { int64_t n0, s0, s1; } tn_idx2_t;
int32_t tn_contract_strided(tn_ctx* ctx, int64_t M, int64_t N, int64_t K,
                            const tn_cplx* A, int64_t a_elems, tn_idx2_t am, tn_idx2_t ak, int32_t conjA,
                            const tn_cplx* B, int64_t b_elems, tn_idx2_t bk, tn_idx2_t bn, int32_t conjB,
                            tn_cplx* C, int64_t c_elems, tn_idx2_t cm, tn_idx2_t cn, tn_cplx alpha);

/* same contraction on DEVICE buffers (no copies), C = alpha*op(A)op(B) + beta*C, enqueued on the context's stream.
 * Used by the multi-GPU MPO-bond-sharded matvec, whose collectives run between the stages. */
int32_t tn_contract_strided_dev(tn_ctx* ctx, int64_t M, int64_t N, int64_t K,
                                const void* A_dev, tn_idx2_t am, tn_idx2_t ak, int32_t conjA,
                                const void* B_dev, tn_idx2_t bk, tn_idx2_t bn, int32_t conjB,
                                void* C_dev, tn_idx2_t cm, tn_idx2_t cn, tn_cplx alpha, tn_cplx beta);

/* ---- projected environments: structures/mps/projmps.jl, abstractprojmps.jl -------------------- */
/* ProjMPS(bra, mpo, ket; rank=2, coeff, center): projmps.jl:16-42.  mpo may be NULL (overlap <bra|ket>). */
int32_t tn_env_create(tn_ctx* ctx, tn_mps* bra, tn_mps* mpo, tn_mps* ket, tn_cplx coeff, int32_t center, tn_env** out);
int32_t tn_env_free(tn_env* e);
int32_t tn_env_buildleft(tn_env* e, int32_t idx);       /* projmps.jl:50-66 */
int32_t tn_env_buildright(tn_env* e, int32_t idx);      /* projmps.jl:74-95 */
int32_t tn_env_movecenter(tn_env* e, int32_t idx);      /* abstractprojmps.jl:60-81 */
int32_t tn_env_center(tn_env* e, int32_t* out);
int32_t tn_env_block_dims(tn_env* e, int32_t idx, int64_t* dims3);   /* (chi_bra, w, chi_ket) */
int32_t tn_env_block_download(tn_env* e, int32_t idx, tn_cplx* out); /* block(projV, idx): abstractprojmps.jl:43-46 */
/* projV[idx] = x (setindex!): abstractprojmps.jl:49-52.  dims3 = (chi_bra, w, chi_ket); data on the host. */
int32_t tn_env_block_upload(tn_env* e, int32_t idx, const int64_t* dims3, const tn_cplx* data);
int32_t tn_env_set_center(tn_env* e, int32_t center);
/* product(projV, A, direction, nsites=2): projmps.jl:103-145 rank-2 branch.  theta / out on the HOST,
 * shape (chi_l, d, d, chi_r) of sites (site, site+1) with site = direction ? center-1 : center. */
int32_t tn_env_product(tn_env* e, const tn_cplx* theta_host, int32_t direction, tn_cplx* out_host);
/* same with DEVICE buffers (no PCIe traffic), `reps` back-to-back applications (reps >= 1) */
int32_t tn_env_product_dev(tn_env* e, const void* theta_dev, int32_t direction, void* out_dev, int32_t reps);
/* instrumented variant for the roofline: CUDA-event time (ms, summed over reps) of the three contraction
 * stages L.theta / .W / .R of the matvec, measured on the context's stream. */
int32_t tn_env_product_profile(tn_env* e, const void* theta_dev, int32_t direction, void* out_dev, int32_t reps, double* stage_ms3);
int32_t tn_env_calculate(tn_env* e, tn_cplx* out);      /* projmps.jl:192-216 */

/* ---- projector branch and sums of projections ----------------------------------------------------------------- */
/* ProjMPS(V, psi; rank=2, squared=true, coeff, center) as built at algorithms/mps/dmrg.jl:144-145: the penalty
 * coeff * |V><V| in psi's local basis; product = coeff * phi <phi, A> with phi = conj(project(...)) (projmps.jl:135-143). */
int32_t tn_env_create_squared(tn_ctx* ctx, tn_mps* V, tn_mps* psi, tn_cplx coeff, int32_t center, tn_env** out);
/* product(projV, A, direction, nsites) for nsites = 1 or 2 (projmps.jl:103-145): the rank-2 branch of an environment with
 * an MPO layer, or the squared branch of tn_env_create_squared.  A / out on the host, shape (chi_l, d[, d], chi_r) of the
 * sites site .. site+nsites-1, site = direction ? center-nsites+1 : center. */
int32_t tn_env_product_n(tn_env* e, const tn_cplx* A_host, int32_t direction, int32_t nsites, tn_cplx* out_host);
/* project(projV, A, direction, nsites): projmps.jl:153-185, for ProjMPS(bra, ket) and ProjMPS(bra, mpo, ket).  The
 * reference never reads A, so it is not an argument.  out_host has the ket-side shape (chi_l, d[, d], chi_r). */
int32_t tn_env_project(tn_env* e, int32_t direction, int32_t nsites, tn_cplx* out_host);
/* ProjMPSSum(projVs; center): projmpssum.jl:11-19.  The members stay owned by their tn_env handles, which must outlive
 * the sum, live in the same context and share the ket MPS. */
int32_t tn_envsum_create(tn_ctx* ctx, int32_t n, tn_env* const* envs, int32_t center, tn_envsum** out);
int32_t tn_envsum_free(tn_envsum* s);
int32_t tn_envsum_movecenter(tn_envsum* s, int32_t idx);                                      /* projmpssum.jl:51-55 */
int32_t tn_envsum_calculate(tn_envsum* s, tn_cplx* out);                                      /* projmpssum.jl:97-108 */
int32_t tn_envsum_product(tn_envsum* s, const tn_cplx* A_host, int32_t direction, int32_t nsites, tn_cplx* out_host); /* :63-73 */
int32_t tn_envsum_project(tn_envsum* s, int32_t direction, int32_t nsites, tn_cplx* out_host);                         /* :81-91 */

/* ---- fused steps (what the drivers call) ------------------------------------------------------ */
/* One direction of a two-site DMRG sweep, algorithms/mps/dmrg.jl:35-63: for every bond movecenter!(Hs),
 * theta = psi[s]*psi[s+1], eigsolve (Lanczos, KrylovKit schedule), replacesites!(..., normalize=true);
 * then movecenter!(Hs, end).  direction 0 = left-to-right.  Returns the last Ritz value and maxbonddim. */
int32_t tn_dmrg_sweep(tn_mps* psi, tn_env* env, int32_t direction, tn_lanczos_t lanczos, tn_trunc_t trunc,
                      double* energy_out, int64_t* maxbond_out);
/* The same half sweep over a ProjMPSSum (several MPOs and / or squared MPS projections: dmrg.jl:128-154, excited states by
 * penalty) and for nsites = 1 or 2 (dmrg.jl:3; the one-site branch of replacesites! moves the centre untruncated). */
int32_t tn_dmrg_sweep_sum(tn_mps* psi, tn_envsum* Hs, int32_t direction, int32_t nsites, tn_lanczos_t lanczos, tn_trunc_t trunc,
                          double* energy_out, int64_t* maxbond_out);
/* One direction of a vmps sweep, algorithms/mps/vmps.jl:36-62: for every bond movecenter!(Vs), vec = conj(project(Vs, .)),
 * replacesites!(psi, vec, site1, direction; trunc) without normalisation; then movecenter!(Vs, end).  The cost
 * norm(psi)^2 - 2|calculate(Vs)| (vmps.jl:16-23) and the convergence logic stay host-side. */
int32_t tn_vmps_sweep(tn_mps* psi, tn_envsum* Vs, int32_t direction, int32_t nsites, tn_trunc_t trunc, int64_t* maxbond_out);
/* eigsolve alone on the two sites at the environment centre (host theta in/out); numops = H_eff applications */
int32_t tn_eigsolve(tn_env* e, const tn_cplx* theta0_host, int32_t direction, tn_lanczos_t lanczos,
                    double* eig_out, tn_cplx* theta_out_host, int32_t* numops_out);

/* ---- device-pointer entry points: for a caller that orchestrates the bond loop itself and only hands the dense pieces to the
 * library -- the multi-GPU MPO-bond-sharded DMRG sweep, whose H_eff application contains collectives, or a Julia
 * caller that keeps KrylovKit-style control with its own linear map. ------------------------------------------------------------ */
/* device address of psi[site]; valid until that site is next modified (replacesites, movecenter, gates, upload) */
int32_t tn_mps_site_ptr(tn_mps* m, int32_t site, void** dev_out);
/* tn_mps_replacesites with the two-site tensor already in device memory */
int32_t tn_mps_replacesites_dev(tn_mps* m, const void* theta_dev, int32_t site, int32_t direction, int32_t normalize, tn_trunc_t trunc);
/* psi[site] = x (abstractmps.jl setindex!) from a device buffer */
int32_t tn_mps_upload_site_dev(tn_mps* m, int32_t site, const int64_t* dims, const void* data_dev);
/* device-to-device copy ordered on the context's stream (completed on return) */
int32_t tn_memcpy_dev(tn_ctx* ctx, void* dst_dev, const void* src_dev, int64_t nbytes);
/* The truncated SVD in three calls for a caller that owns the Jacobi pair schedule (distributed sweeps over several GPUs,
 * tnb200/sharded.py: every rank begins on the same matrix, rotates the column-block pairs its schedule assigns to it, exchanges
 * column blocks with its peers, and finishes on the complete Z).  begin: init + the two QR steps; returns the device address of
 * Z = [W; V] (leading dimension ldz, zrows rows, nblocks blocks of 32 columns; block b starts at Z + b*32*ldz) and the convergence
 * threshold.  step: Gram -> EVD -> rotation over npairs disjoint block pairs (pairs: npairs x 2), returns the largest normalised
 * off-diagonal Gram entry seen.  finish: norms, sort, the reference's truncation rule (tensors.jl:201-215).  factors: U (m x k),
 * S (k doubles), Vh (k x n) into device buffers (any may be NULL); tn_mps_replacesites_factored = replacesites! with the factors
 * left in the context (gmps.jl:215-266). */
int32_t tn_svd_dist_begin(tn_ctx* ctx, const void* mat_dev, int64_t m, int64_t n, void** Z_dev, int64_t* ldz, int64_t* zrows,
                          int32_t* nblocks, double* tol);
int32_t tn_svd_dist_step(tn_ctx* ctx, const int32_t* pairs, int32_t npairs, double* offmax);
int32_t tn_svd_dist_finish(tn_ctx* ctx, tn_trunc_t trunc, int32_t sweeps, int64_t* k_out);
int32_t tn_svd_dist_factors(tn_ctx* ctx, void* U_dev, void* S_dev, void* Vh_dev);
int32_t tn_mps_replacesites_factored(tn_mps* m, int32_t site, int32_t direction, int32_t normalize);
/* KrylovKit eigsolve(f, x0, 1, :SR; krylovdim, maxiter, tol, ishermitian=true) (dmrg.jl:51-53) for a caller-supplied Hermitian
 * linear map on n-element complex128 device vectors.  apply(user, in_dev, out_dev) is entered with in_dev complete; all its writes
 * to out_dev must be complete (or enqueued on the context's stream) when it returns 0.  Non-zero aborts with TN_ERR_INVALID. */
typedef int32_t (*tn_apply_fn)(void* user, const void* in_dev, void* out_dev);
int32_t tn_eigsolve_fn(tn_ctx* ctx, int64_t n, const void* theta0_dev, void* theta_out_dev, tn_lanczos_t lanczos,
                       tn_apply_fn apply, void* user, double* eig_out, int32_t* numops_out);

/* GateList upload: nrows rows; counts[r] gates in row r; per gate (flattened in row order) the first
 * site, the number of sites (1 or 2) and a host pointer to the gate tensor (out1,in1[,out2,in2]),
 * column-major, as produced by trotterize(): gatelist.jl:75-121. */
int32_t tn_gates_upload(tn_ctx* ctx, int32_t d, int32_t nrows, const int32_t* counts, const int32_t* sites,
                        const int32_t* nsites, const tn_cplx* const* gate_ptrs, tn_gates** out);
int32_t tn_gates_free(tn_gates* g);
/* applygates!(psi, gates; cutoff, maxdim, mindim): gatelist.jl:191-227 (MPS rank 1 or MPO rank 2) */
int32_t tn_apply_gates(tn_mps* psi, tn_gates* gates, tn_trunc_t trunc);
/* applygates(psi, gates; error=true, kwargs...): gatelist.jl:191-223 -- also returns the product over the two-site gates of the
 * truncation fidelity |<Theta', Theta'_truncated>|^2 (gatelist.jl:160-168); one extra GEMM, dot product and host read per gate.
 * One-site gates contribute a factor 1 (the reference multiplies by |<O A, U>|^2 with U the gauge-moved tensor, a number that
 * depends on the phases the SVD happens to pick). */
int32_t tn_apply_gates_fidelity(tn_mps* psi, tn_gates* gates, tn_trunc_t trunc, double* fidelity_out);

/* ---- infinite TEBD (Vidal form): algorithms/mps/itebd.jl:71-119 for a two-site unit cell -------------------------------- */
/* iGMPS upload: L = 2 cell tensors (D_{i-1}, d, D_i) with D_2 = D_0, and the singular values to the LEFT of each tensor
 * (sing_ptrs[i]: D_{i-1} doubles), as the reference stores them (itebd.jl:78-83).  dims: L x 3 row-major. */
int32_t tn_imps_create(tn_ctx* ctx, int32_t d, int32_t L, const int64_t* dims, const tn_cplx* const* site_ptrs,
                       const double* const* sing_ptrs, tn_imps** out);
int32_t tn_imps_free(tn_imps* p);
int32_t tn_imps_dims(tn_imps* p, int64_t* dims /* L x 3 */);
/* tensors[site], singulars[site] and norms[site] (any output may be NULL) */
int32_t tn_imps_download(tn_imps* p, int32_t site, tn_cplx* tensor_out, double* singulars_out, double* norm_out);
/* nsteps passes of _itebd_apply_gates_mps!(psi, gate, mindim, maxdim, cutoff) (itebd.jl:71-119): per pass the gate
 * (o1,i1,o2,i2) on bond (1,2) then on bond (2,1); Theta = S_a G_a S_b G_b S_a, truncated SVD, G_a = S_a^-1 U, G_b = V^H S_a^-1,
 * S_b normalised with the log of its norm added to norms[b]. */
int32_t tn_itebd_apply_gate(tn_imps* p, const tn_cplx* gate_host, int32_t nsteps, tn_trunc_t trunc);

/* One QJMC trajectory, algorithms/mps/qjmc.jl:59-164: per step applygates!, normalize!, emission rates (single-site jump
 * operators jump_ops[k] d x d at jump_sites[k], rates scaled by jump_coeffs[k]^2), jump test and jump update.
 * classical != 0 (the reference's default, :88-112): jump when u1 > exp(-sum(rates) dt), channel from u2.
 * classical == 0 (:65-87): jump when u0 > norm(psi)^2 after the non-unitary gates, channel from u1.
 * uniforms: 3 per step, indexed [3*step + slot] (host) or NULL for the built-in
 * counter-based generator keyed by (seed, trajectory, step, slot).  Observables: <obs_op> on every site is
 * written to obs_out[(step/save_every - 1) * N + i] every save_every steps (obs_op may be NULL).
 * jumps_out / jumptimes_out: up to jump_cap records (1-based channel index, time). */
int32_t tn_qjmc_run(tn_mps* psi, tn_gates* gates, int32_t njump, const int32_t* jump_sites, const tn_cplx* jump_ops,
                    const double* jump_coeffs, int32_t steps, double dt, tn_trunc_t trunc, const double* uniforms,
                    uint64_t seed, uint64_t trajectory, const tn_cplx* obs_op, int32_t save_every, tn_cplx* obs_out,
                    int32_t* jumps_out, double* jumptimes_out, int32_t jump_cap, int32_t* njumps_out, int32_t classical);

/* inner(st, psi, oplist, phi): mps.jl:87-134.  out[t] = coeffs[t] * <psi| O_t |phi> for nterms operator strings
 * O_t = product of nops[t] single-site operators; op_sites (1-based, strictly ascending inside a term) and ops_host
 * (d x d each, column-major) are flattened over the terms in order.  Evaluated like the reference: overlap blocks
 * ProjMPS(psi, phi), the left block carried through the sites of the string (identity on the gaps), closed with the
 * right block -- no canonical-form shortcut.  tebd.jl:51,90 (energy) and the observers call this. */
int32_t tn_inner_oplist(tn_mps* psi, tn_mps* phi, int32_t nterms, const int32_t* nops, const int32_t* op_sites,
                        const tn_cplx* ops_host, const tn_cplx* coeffs, tn_cplx* out);

/* Ensemble of independent QJMC trajectories (the caller's loop around qjmc_simulation, examples/qjmc.jl:52; SURVEY 8(e)
 * trajectory-level parallelism).  All trajectories start from the same host MPS (dims: N x 3, site_ptrs) and the same gate
 * list (arguments as tn_gates_upload).  They are handed out dynamically to `nworkers` host threads, each owning a CUDA
 * stream + workspace on `device`, so that the small kernels of different trajectories overlap on the GPU.  Random numbers
 * come from the counter-based generator keyed by (seed, traj_ids[t], step) (traj_ids == NULL: t), i.e. the result of a
 * trajectory does not depend on the worker count or on which GPU / rank ran it.
 * Outputs for the trajectory at position t: obs_out[(t * (steps / save_every) + s) * N + i], njumps_out[t],
 * jumps_out / jumptimes_out[t * jump_cap + j] (any of the last three may be NULL). */
int32_t tn_qjmc_ensemble(int32_t device, int32_t nworkers, int32_t ntraj, const uint64_t* traj_ids,
                         int32_t d, int32_t N, const int64_t* dims, const tn_cplx* const* site_ptrs, int32_t center,
                         int32_t nrows, const int32_t* counts, const int32_t* gate_sites, const int32_t* gate_nsites,
                         const tn_cplx* const* gate_ptrs, int32_t njump, const int32_t* jump_sites, const tn_cplx* jump_ops,
                         const double* jump_coeffs, int32_t steps, double dt, tn_trunc_t trunc, uint64_t seed,
                         const tn_cplx* obs_op, int32_t save_every, tn_cplx* obs_out, int32_t* njumps_out,
                         int32_t* jumps_out, double* jumptimes_out, int32_t jump_cap, int32_t classical);

/* ---- MPO-bond-sharded H_eff application on several GPUs of ONE process (the library owns the NCCL
 * communicators) -----------------------------------------------------------------------------------------------------------
 * Replaces product(projV, A, ...) (structures/mps/projmps.jl:103-145) when the environment blocks of a large-chi bond are spread over
 * the box.  L (chi, w, chi), R (chi2, w2, chi2), M1 (w, d, d, w1), M2 (w1, d, d, w2): host buffers, borrowed for the call; device g
 * keeps rows [m0, m1) of the fused (a, w) index of L and an equal chunk of the fused (b', w2) index of R.  apply: theta (chi, d, d,
 * chi2) and out are host buffers; per call one local chi^3 stage, the small MPO stage, ncclReduceScatter, the second local chi^3
 * stage and ncclAllReduce run on one stream per device.  NCCL (libnccl.so.2) is loaded at run time; ngpus = 1 needs none. */
typedef struct tn_shard tn_shard;
int32_t tn_heff_sharded_create(int32_t ngpus, const int32_t* device_ids, int64_t chi, int64_t chi2, int32_t d, int64_t w, int64_t w1, int64_t w2,
                               const tn_cplx* L, const tn_cplx* R, const tn_cplx* M1, const tn_cplx* M2, tn_cplx coeff, tn_shard** out);
int32_t tn_heff_sharded_apply(tn_shard* h, const tn_cplx* theta, tn_cplx* out);
int32_t tn_heff_sharded_free(tn_shard* h);

#ifdef __cplusplus
}
#endif
#endif /* TN_C_API_H */
