// Truncated SVD on the GPU: blocked one-sided Jacobi with on-device rank truncation.
#pragma once
#include "tn_common.cuh"
#include <map>

namespace tn {

struct Trunc { double cutoff; long long maxdim; long long mindim; };

// Workspace + state of one factorisation (owned by a Ctx; reused across calls).
struct SvdWork {
  cplx* Z = nullptr; size_t Z_cap = 0;            // [W ; V] stacked, (rows + ncols) x ncols, ld = rows + ncols
  cplx* Gpart = nullptr; size_t G_cap = 0;        // split-K Gram partials
  cplx* J = nullptr; size_t J_cap = 0;            // per-pair 64x64 rotations
  double* sig = nullptr; int* perm = nullptr; size_t s_cap = 0;   // sorted singular values + permutation
  double* sig2 = nullptr;                          // unsorted squared column norms
  unsigned long long* offmax = nullptr;            // convergence measure (double bits)
  int* kout = nullptr;                             // device-side truncation rank
  std::map<int, int*> tables;                      // round-robin pair tables per block count
  // description of the last factorisation
  int m = 0, n = 0, rows = 0, ncols = 0, ncols_pad = 0, ldz = 0, nsv = 0, k = 0, sweeps = 0;
  bool transposed = false;
  double last_off = 0;
};

// Factorises the m x n column-major matrix M (leading dimension ld) and applies the reference's
// truncation rule (src/tensors.jl:201-215).  Returns k; factors stay in `w` until gathered.
int svd_factor(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s);
// U (m x k, leading dim ldu), optionally multiplied by S on the right.
void svd_gather_U(SvdWork& w, cplx* U, long long ldu, bool times_S, cudaStream_t s);
// V^H (k x n, leading dim ldv), optionally multiplied by S on the left.
void svd_gather_Vh(SvdWork& w, cplx* Vh, long long ldv, bool times_S, cudaStream_t s);
// first k singular values (device -> device copy)
void svd_copy_S(SvdWork& w, double* S, cudaStream_t s);
void svd_free(SvdWork& w);

}  // namespace tn
