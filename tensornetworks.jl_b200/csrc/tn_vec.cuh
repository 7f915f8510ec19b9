// Vector / elementwise kernels (HBM-bound): Lanczos dot/axpy/norm, scalings, gate mix, gathers.
#pragma once
#include "tn_common.cuh"

namespace tn {

// out[j] = <x_j, y> = sum conj(x_j[i]) * y[i], j < nx (nx <= 4); deterministic two-stage reduction.
// `partials` needs 4 * DOT_BLOCKS cplx.  Result written to device memory `out`.
constexpr int DOT_BLOCKS = 592;  // 4 x 148 SMs
void zdots(long long n, int nx, const cplx* const* x_host_ptrs, const cplx* y, cplx* out, cplx* partials, cudaStream_t s);
// y -= sum_j h[j] * x_j   (h in device memory); if real_only_diag >= 0, uses Re(h[diag]) for that j
void zsubproj(long long n, int nx, const cplx* const* x_host_ptrs, const cplx* h, cplx* y, cudaStream_t s);
// out = x * (1 / sqrt(Re(nrm2[0])))  (nrm2 = <x,x> in device memory)
void zscale_invnorm(long long n, const cplx* x, const cplx* nrm2, cplx* out, cudaStream_t s);
// out = sum_j c[j] * x_j  with host-side real coefficients
void zlincomb(long long n, int nx, const cplx* const* x_host_ptrs, const double* c_host, cplx* out, cudaStream_t s);
// x *= alpha (host complex scalar)
void zscal(long long n, cplx alpha, cplx* x, cudaStream_t s);
// y += alpha * x
void zaxpy(long long n, cplx alpha, const cplx* x, cplx* y, cudaStream_t s);
// y += alpha * h[0] * x   (h in device memory: a dot product that never visits the host)
void zaxpy_dev(long long n, cplx alpha, const cplx* h, const cplx* x, cplx* y, cudaStream_t s);

// out(l, m, r) = f(sl[l]) * in(l, m, r) * f(sr[r]) with real scale vectors (nullptr = 1) and f(x) = x or 1/x: the diagonal
// singular-value matrices of a Vidal-form MPS (itebd.jl:78-83, :101-102) applied without forming diag() or a GEMM
void scale_lr(const cplx* in, cplx* out, long long nl, long long nm, long long nr, const double* sl, bool invl, const double* sr, bool invr,
              cudaStream_t s);

// Two-site gate mix (reference gatelist.jl:149-154):
//   out(l, o1, [p1], o2, [p2], r) = sum_{i1,i2} G(o1,i1,o2,i2) * in(l, i1, [p1], i2, [p2], r)
// d = physical dim, inner = d for rank-2 (extra passive physical index per site) else 1.
void gate_mix2(const cplx* in, cplx* out, const cplx* G, long long chiL, int d, int inner, long long chiR, cudaStream_t s);
// One-site operator on the first physical index: out(l,o,[p],r) = sum_i O(o,i) in(l,i,[p],r)   (mps.jl:141-152)
void op_apply1(const cplx* in, cplx* out, const cplx* O, long long chiL, int d, int inner, long long chiR, cudaStream_t s);

}  // namespace tn
