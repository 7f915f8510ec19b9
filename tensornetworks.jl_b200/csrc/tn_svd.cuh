// Truncated SVD on the GPU: blocked one-sided Jacobi with on-device rank truncation.
#pragma once
#include "tn_common.cuh"
#include <map>
#include <vector>

namespace tn {

struct Trunc { double cutoff; long long maxdim; long long mindim; };

// Workspace + state of one factorisation (owned by a Ctx; reused across calls).
struct SvdWork {
  cplx* Z = nullptr; size_t Z_cap = 0;            // [W ; V] stacked, (rows + ncols) x ncols, ld >= rows + ncols
  int* skip = nullptr; size_t skip_cap = 0;        // per-pair flags: Gram block already diagonal, rotation skipped
  int* dtab = nullptr; size_t dtab_cap = 0;        // pair list of a caller-scheduled step (svd_dist_step)
  cplx* Gpart = nullptr; size_t G_cap = 0;        // split-K Gram partials
  cplx* J = nullptr; size_t J_cap = 0;            // per-pair 64x64 rotations
  double* sig = nullptr; int* perm = nullptr; size_t s_cap = 0;   // sorted singular values + permutation
  double* sig2 = nullptr;                          // unsorted squared column norms
  unsigned long long* offmax = nullptr;            // convergence measure (double bits)
  int* cflag = nullptr;                            // CholeskyQR3 early-termination flag of the current panel
  int* kout = nullptr;                             // device-side truncation rank
  std::map<int, int*> tables;                      // round-robin pair tables per block count
  // split pair schedule (tn_svd.cu, "Split schedule"): steps that fall apart into independent groups on auxiliary streams
  struct SplitSched* split = nullptr;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  // QR preconditioner (two GEMM-based QR factorisations: A = Q1 R1, R1^H = Q2 R2, Jacobi on X = R2^H)
  cplx* Q1 = nullptr; size_t Q1_cap = 0;           // rows x ncols_pad
  cplx* Q2 = nullptr; size_t Q2_cap = 0;           // ncols_pad x ncols_pad
  cplx* Ra = nullptr; size_t Ra_cap = 0;           // R of the current factorisation (ncols_pad^2)
  cplx* Rb = nullptr; size_t Rb_cap = 0;           // second-pass R / scratch
  cplx* Rc = nullptr; size_t Rc_cap = 0;           // product scratch
  cplx* small = nullptr;                           // 3 x 64x64: Rinv, Rtot, scratch
  cplx* Tg = nullptr; size_t Tg_cap = 0;           // gathered / scaled singular-vector block (ncols_pad x k)
  bool precond = false; int jrows = 0;             // Jacobi ran on an jrows x ncols_pad matrix
  // "W-only" factorisation (svd_factor with iso != 0): the caller needs ONE orthonormal factor (the isometry that becomes a site
  // tensor) and the other one multiplied by S.  Then the right rotations V' are never accumulated (half of the rotation flops):
  // with T = Q1 R1 (T = A or A^H, the tall orientation), R1^H = Q2 R2, X = R2^H and X V' = U' Sigma, the isometry is Q1 U' and
  // the other factor is S V_T^H = (Q1 U')^H T = U'^H R1 -- one GEMM with the saved R1, no Q2, no V'.
  bool wonly = false;
  cplx* R1 = nullptr; size_t R1_cap = 0;           // R1 of the first QR step (W-only mode)
  // Gauge moves that cannot truncate (svd_factor with need_values = false, cutoff = 0 and maxdim >= min(m, n)): any orthogonal
  // factorisation yields the same state, so the SVD is replaced by  1 = one QR step of the tall orientation (isometry Q, other
  // factor R)  or  2 = nothing at all when the isometry sits on the short side (identity, other factor = the matrix itself).
  int qr_mode = 0;
  const cplx* M0 = nullptr; long long ld0 = 0;     // the caller's matrix (qr_mode 2; valid until the gathers)
  // description of the last factorisation
  int m = 0, n = 0, rows = 0, ncols = 0, ncols_pad = 0, ldz = 0, nsv = 0, k = 0, sweeps = 0;
  bool transposed = false;
  double last_off = 0;
  // set by svd_factor when the calling thread takes part in a batching round (SvdBatcher): the factors live in a slot of a
  // shared batched workspace and the gathers read them through this single-problem view
  bool use_view = false;
  SvdWork* bview = nullptr;
};

// Factorises the m x n column-major matrix M (leading dimension ld) and applies the reference's
// truncation rule (src/tensors.jl:201-215).  Returns k; factors stay in `w` until gathered.
// iso: 0 = both singular-vector sets are accumulated (any combination of gathers); 1 = the caller gathers U plain and S V^H;
// 2 = V^H plain and U S (then the cheaper W-only factorisation is used when the isometry sits on the long side of M).
// need_values = false: the caller uses neither the singular values nor the canonical choice of the singular vectors (moveleft! /
// moveright! inside movecenter!, gmps.jl:60-112) -- see SvdWork::qr_mode.
int svd_factor(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso = 0, bool need_values = true);
// U (m x k, leading dim ldu), optionally multiplied by S on the right.
void svd_gather_U(SvdWork& w, cplx* U, long long ldu, bool times_S, cudaStream_t s);
// V^H (k x n, leading dim ldv), optionally multiplied by S on the left.
void svd_gather_Vh(SvdWork& w, cplx* Vh, long long ldv, bool times_S, cudaStream_t s);
// first k singular values (device -> device copy)
void svd_copy_S(SvdWork& w, double* S, cudaStream_t s);
void svd_free(SvdWork& w);
// Factorisation in three calls for a caller that owns the pair schedule (distributed sweeps): see tn_svd.cu.
int svd_dist_begin(SvdWork& w, const cplx* M, int m, int n, long long ld, cudaStream_t s);      // returns the number of 32-column blocks
double svd_dist_tol(const SvdWork& w);
double svd_dist_step(SvdWork& w, const int* pairs_host, int npairs, cudaStream_t s);
int svd_dist_finish(SvdWork& w, Trunc tr, int sweeps, cudaStream_t s);

// Workspace of a batched factorisation: B same-shape problems stacked side by side (tn_svd.cu, "Batched factorisation").
struct SvdBatch {
  int B = 0, m = 0, n = 0, rows = 0, ncols = 0, npad = 0, ldz = 0, jrows = 0, sweeps = 0;
  bool transposed = false, precond = false;
  bool wonly = false;                                // W-only factorisation (see SvdWork::wonly); R1 of every problem kept
  int qr_mode = 0;                                   // 1: only the first QR step was run (gauge moves that cannot truncate)
  cplx* R1 = nullptr; size_t R1_cap = 0;
  cplx* Z = nullptr; size_t Z_cap = 0;
  cplx* Q1 = nullptr; size_t Q1_cap = 0;
  cplx* Q2 = nullptr; size_t Q2_cap = 0;
  cplx* Ra = nullptr; size_t Ra_cap = 0;
  cplx* Rb = nullptr; size_t Rb_cap = 0;
  cplx* Rc = nullptr; size_t Rc_cap = 0;
  cplx* Gpart = nullptr; size_t G_cap = 0;
  cplx* J = nullptr; size_t J_cap = 0;
  cplx* small = nullptr; size_t small_cap = 0;      // per problem: Rinv, Rtot (64 x 64 each)
  cplx* Tg = nullptr; size_t Tg_cap = 0;
  int* skip = nullptr; size_t skip_cap = 0;
  int* cflag = nullptr; size_t cflag_cap = 0;
  int* kout = nullptr; size_t kout_cap = 0;
  double* sig = nullptr; double* sig2 = nullptr; int* perm = nullptr; size_t s_cap = 0;
  unsigned long long* offmax = nullptr;
  std::map<std::pair<int, int>, int*> tables;        // (B, column blocks) -> round-robin pair table of all problems
  std::vector<int> k;                                // truncation rank of every problem (host)
};
// Factorises B matrices Ms[b] (device pointers, each m x n column-major with leading dimension ld) and applies the
// reference's truncation rule to each; the ranks are left in w.k.
// iso / qr_only: as the iso / need_values = false arguments of svd_factor (qr_only requires iso on the long side).
void svd_batched_factor(SvdBatch& w, int B, const cplx* const* Ms, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso = 0, bool qr_only = false);
void svd_batched_gather_U(SvdBatch& w, int b, cplx* U, long long ldu, bool times_S, cudaStream_t s);
void svd_batched_gather_Vh(SvdBatch& w, int b, cplx* Vh, long long ldv, bool times_S, cudaStream_t s);
void svd_batched_copy_S(SvdBatch& w, int b, double* S, cudaStream_t s);
void svd_batched_free(SvdBatch& w);

// Batching rounds across host threads (QJMC ensembles): every worker thread runs the ordinary single-trajectory code; when it
// reaches svd_factor it parks its request, and once every active worker is parked the last one to arrive factorises all
// requests -- grouped by shape -- with svd_batched_factor and wakes the others, which then gather their own factors.
struct SvdBatcher;
SvdBatcher* svd_batcher_create(int nworkers);      // all nworkers count as active from the start
void svd_batcher_destroy(SvdBatcher* b);
void svd_batcher_attach(SvdBatcher* b);            // the calling thread's svd_factor calls join the rounds from now on
void svd_batcher_detach(cudaStream_t s);           // the calling thread leaves (may complete a round on stream s)
long long svd_batcher_rounds(SvdBatcher* b, long long* problems);
// tn_jacobi.cu: the Gram block G_p = P_p^H P_p and the in-place rotation Z(:, pair p) <- Z(:, pair p) J_p of column-block pairs
// (tab: 2 block indices per pair, 32 columns per block) as dedicated kernels
void jacobi_gram64(const cplx* Z, long long ldz, int rows, const int* tab, int npairs, cplx* G, int max_split, cudaStream_t s, const int* skip = nullptr);
void jacobi_rot64(cplx* Z, long long ldz, int rows, const int* tab, int npairs, const cplx* J, const int* skip, cudaStream_t s);
// T_b(:, tile i) -= P_b * C_b(:, tile i): the rank-64 trailing updates of the block Gram-Schmidt QR (P_b = column blocks ptab[2b], ptab[2b+1] of A;
// tile (b, i) = column blocks ttab[2 (b ntiles + i)], .. + 1 of T; C_b = C + cstride b, 64 x 64 ntiles coefficients, leading dimension ldc)
void jacobi_update64(const cplx* A, long long lda, const int* ptab, cplx* T, long long ldt, int rows, const int* ttab, int nprob, int ntiles,
                     const cplx* C, long long ldc, long long cstride, cudaStream_t s);
// C_b(:, tile i) = P_b^H T_b(:, tile i): the projection coefficients in front of jacobi_update64 (same tables / strides; C zeroed by the caller)
void jacobi_cross64(const cplx* A, long long lda, const int* ptab, const cplx* T, long long ldt, int rows, const int* ttab, int nprob, int ntiles,
                    cplx* C, long long ldc, long long cstride, int max_split, cudaStream_t s);
long long svd_split_schedule_dump(int nb, int groups, int* out5, long long cap5);   // tests: the split pair schedule of the Jacobi sweeps
void svd_set_precond(int mode);   // 1 (default): two-step QR preconditioning before the Jacobi sweeps; 0: plain Jacobi

}  // namespace tn
