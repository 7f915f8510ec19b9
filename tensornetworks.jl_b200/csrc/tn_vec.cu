#include "tn_vec.cuh"

namespace tn {
void count_launch(int n);

struct Ptrs4 { const cplx* p[4]; };

__global__ void __launch_bounds__(256) zdots_stage1(long long n, int nx, Ptrs4 xs, const cplx* __restrict__ y, cplx* __restrict__ partials) {
  double ar[4] = {0, 0, 0, 0}, ai[4] = {0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx b = y[i];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nx) {
        cplx a = xs.p[j][i];
        ar[j] += a.x * b.x + a.y * b.y;
        ai[j] += a.x * b.y - a.y * b.x;
      }
  }
  __shared__ double sh[8][8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double r = ar[j], im = ai[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    if (lane == 0) { sh[w][2 * j] = r; sh[w][2 * j + 1] = im; }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0;
    for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
    reinterpret_cast<double*>(partials)[(long long)blockIdx.x * 8 + threadIdx.x] = s;
  }
}
__global__ void zdots_stage2(int nblocks, int nx, const cplx* __restrict__ partials, cplx* __restrict__ out) {
  // one warp per output component (8 doubles); fixed summation order => deterministic
  int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* p = reinterpret_cast<const double*>(partials);
  double s = 0;
  for (int b = lane; b < nblocks; b += 32) s += p[(long long)b * 8 + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0 && (c >> 1) < nx) reinterpret_cast<double*>(out)[c] = s;
}
void zdots(long long n, int nx, const cplx* const* xp, const cplx* y, cplx* out, cplx* partials, cudaStream_t s) {
  TN_CHECK(nx >= 1 && nx <= 4, "zdots: 1..4 vectors");
  Ptrs4 xs{};
  for (int j = 0; j < nx; ++j) xs.p[j] = xp[j];
  int blocks = (int)std::min<long long>(DOT_BLOCKS, (n + 255) / 256);
  if (blocks < 1) blocks = 1;
  zdots_stage1<<<blocks, 256, 0, s>>>(n, nx, xs, y, partials);
  zdots_stage2<<<1, 256, 0, s>>>(blocks, nx, partials, out);
  TN_CUDA(cudaGetLastError());
  count_launch(2);
}

__global__ void __launch_bounds__(256) zsubproj_kernel(long long n, int nx, Ptrs4 xs, const cplx* __restrict__ h, cplx* __restrict__ y) {
  cplx hh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) hh[j] = j < nx ? h[j] : make_double2(0, 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx v = y[i];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nx) {
        cplx a = xs.p[j][i];
        v.x -= hh[j].x * a.x - hh[j].y * a.y;
        v.y -= hh[j].x * a.y + hh[j].y * a.x;
      }
    y[i] = v;
  }
}
void zsubproj(long long n, int nx, const cplx* const* xp, const cplx* h, cplx* y, cudaStream_t s) {
  Ptrs4 xs{};
  for (int j = 0; j < nx; ++j) xs.p[j] = xp[j];
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zsubproj_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, nx, xs, h, y);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

__global__ void __launch_bounds__(256) zscale_invnorm_kernel(long long n, const cplx* __restrict__ x, const cplx* __restrict__ nrm2, cplx* __restrict__ out) {
  double inv = nrm2[0].x > 0.0 ? 1.0 / sqrt(nrm2[0].x) : 0.0;   // invariant subspace (beta == 0): emit zeros, never NaN
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx v = x[i];
    out[i] = make_double2(v.x * inv, v.y * inv);
  }
}
void zscale_invnorm(long long n, const cplx* x, const cplx* nrm2, cplx* out, cudaStream_t s) {
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zscale_invnorm_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, x, nrm2, out);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

struct Coef4 { double c[4]; };
__global__ void __launch_bounds__(256) zlincomb_kernel(long long n, int nx, Ptrs4 xs, Coef4 c, cplx* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double r = 0, im = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nx) { cplx a = xs.p[j][i]; r += c.c[j] * a.x; im += c.c[j] * a.y; }
    out[i] = make_double2(r, im);
  }
}
void zlincomb(long long n, int nx, const cplx* const* xp, const double* ch, cplx* out, cudaStream_t s) {
  Ptrs4 xs{}; Coef4 c{};
  for (int j = 0; j < nx; ++j) { xs.p[j] = xp[j]; c.c[j] = ch[j]; }
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zlincomb_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, nx, xs, c, out);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

__global__ void __launch_bounds__(256) zscal_kernel(long long n, cplx a, cplx* __restrict__ x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx v = x[i];
    x[i] = make_double2(a.x * v.x - a.y * v.y, a.x * v.y + a.y * v.x);
  }
}
void zscal(long long n, cplx alpha, cplx* x, cudaStream_t s) {
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zscal_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, alpha, x);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}
__global__ void __launch_bounds__(256) zaxpy_kernel(long long n, cplx a, const cplx* __restrict__ x, cplx* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx v = x[i], w = y[i];
    y[i] = make_double2(w.x + a.x * v.x - a.y * v.y, w.y + a.x * v.y + a.y * v.x);
  }
}
void zaxpy(long long n, cplx alpha, const cplx* x, cplx* y, cudaStream_t s) {
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zaxpy_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, alpha, x, y);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

__global__ void __launch_bounds__(256) zaxpy_dev_kernel(long long n, cplx a, const cplx* __restrict__ h, const cplx* __restrict__ x, cplx* __restrict__ y) {
  const cplx hv = h[0];
  const double cr = a.x * hv.x - a.y * hv.y, ci = a.x * hv.y + a.y * hv.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    cplx v = x[i], w = y[i];
    y[i] = make_double2(w.x + cr * v.x - ci * v.y, w.y + cr * v.y + ci * v.x);
  }
}
void zaxpy_dev(long long n, cplx alpha, const cplx* h, const cplx* x, cplx* y, cudaStream_t s) {
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zaxpy_dev_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, alpha, h, x, y);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

__global__ void __launch_bounds__(256) scale_lr_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, long long nl, long long nlm, long long total,
                                                       const double* __restrict__ sl, int invl, const double* __restrict__ sr, int invr) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    double f = 1.0;
    if (sl) { double v = sl[e % nl]; f *= invl ? 1.0 / v : v; }
    if (sr) { double v = sr[e / nlm]; f *= invr ? 1.0 / v : v; }
    cplx x = in[e];
    out[e] = make_double2(x.x * f, x.y * f);
  }
}
void scale_lr(const cplx* in, cplx* out, long long nl, long long nm, long long nr, const double* sl, bool invl, const double* sr, bool invr,
              cudaStream_t s) {
  const long long total = nl * nm * nr;
  if (total <= 0) return;
  int blocks = (int)std::min<long long>(148 * 8, (total + 255) / 256);
  scale_lr_kernel<<<std::max(blocks, 1), 256, 0, s>>>(in, out, nl, nl * nm, total, sl, invl ? 1 : 0, sr, invr ? 1 : 0);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// ---- gate mix -------------------------------------------------------------------------------
// Tensor viewed as (l, i1, p1, i2, p2, r) with p1,p2 of extent `inner` (1 for an MPS).  One thread
// per (l, p1, p2, r) point applies the d^2 x d^2 gate to the d^2 (i1,i2) fibre: reads and writes are
// coalesced along l.  HBM-bound: 2 * 16 B per element.
template <int D>
__global__ void __launch_bounds__(256) gate_mix2_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, const cplx* __restrict__ G,
                                                         long long chiL, int inner, long long chiR) {
  __shared__ cplx g[D * D * D * D];
  for (int i = threadIdx.x; i < D * D * D * D; i += blockDim.x) g[i] = G[i];   // G(o1,i1,o2,i2) column-major
  __syncthreads();
  const long long s_i1 = chiL, s_p1 = chiL * D, s_i2 = s_p1 * inner, s_p2 = s_i2 * D, s_r = s_p2 * inner;
  const long long total = chiL * inner * inner * chiR;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long l = idx % chiL, rest = idx / chiL;
    int p1 = (int)(rest % inner); rest /= inner;
    int p2 = (int)(rest % inner); long long r = rest / inner;
    long long base = l + p1 * s_p1 + p2 * s_p2 + r * s_r;
    cplx v[D][D];
#pragma unroll
    for (int i2 = 0; i2 < D; ++i2)
#pragma unroll
      for (int i1 = 0; i1 < D; ++i1) v[i1][i2] = in[base + i1 * s_i1 + i2 * s_i2];
#pragma unroll
    for (int o2 = 0; o2 < D; ++o2)
#pragma unroll
      for (int o1 = 0; o1 < D; ++o1) {
        double xr = 0, xi = 0;
#pragma unroll
        for (int i2 = 0; i2 < D; ++i2)
#pragma unroll
          for (int i1 = 0; i1 < D; ++i1) {
            cplx c = g[o1 + D * (i1 + D * (o2 + D * i2))];
            xr += c.x * v[i1][i2].x - c.y * v[i1][i2].y;
            xi += c.x * v[i1][i2].y + c.y * v[i1][i2].x;
          }
        out[base + o1 * s_i1 + o2 * s_i2] = make_double2(xr, xi);
      }
  }
}
void gate_mix2(const cplx* in, cplx* out, const cplx* G, long long chiL, int d, int inner, long long chiR, cudaStream_t s) {
  long long total = chiL * inner * inner * chiR;
  int blocks = (int)std::min<long long>(148 * 8, (total + 255) / 256);
  blocks = std::max(blocks, 1);
  if (d == 2) gate_mix2_kernel<2><<<blocks, 256, 0, s>>>(in, out, G, chiL, inner, chiR);
  else if (d == 3) gate_mix2_kernel<3><<<blocks, 256, 0, s>>>(in, out, G, chiL, inner, chiR);
  else if (d == 4) gate_mix2_kernel<4><<<blocks, 256, 0, s>>>(in, out, G, chiL, inner, chiR);
  else throw Error(-1, "gate_mix2: physical dimension must be 2, 3 or 4");
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

template <int D>
__global__ void __launch_bounds__(256) op_apply1_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, const cplx* __restrict__ O,
                                                         long long chiL, long long tail) {
  __shared__ cplx o[D * D];
  if (threadIdx.x < D * D) o[threadIdx.x] = O[threadIdx.x];   // O(out,in) column-major
  __syncthreads();
  const long long total = chiL * tail;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long l = idx % chiL, r = idx / chiL;
    long long base = l + r * chiL * D;
    cplx v[D];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = in[base + i * chiL];
#pragma unroll
    for (int q = 0; q < D; ++q) {
      double xr = 0, xi = 0;
#pragma unroll
      for (int i = 0; i < D; ++i) { cplx c = o[q + D * i]; xr += c.x * v[i].x - c.y * v[i].y; xi += c.x * v[i].y + c.y * v[i].x; }
      out[base + q * chiL] = make_double2(xr, xi);
    }
  }
}
void op_apply1(const cplx* in, cplx* out, const cplx* O, long long chiL, int d, int inner, long long chiR, cudaStream_t s) {
  long long tail = (long long)inner * chiR;
  long long total = chiL * tail;
  int blocks = std::max(1, (int)std::min<long long>(148 * 8, (total + 255) / 256));
  if (d == 2) op_apply1_kernel<2><<<blocks, 256, 0, s>>>(in, out, O, chiL, tail);
  else if (d == 3) op_apply1_kernel<3><<<blocks, 256, 0, s>>>(in, out, O, chiL, tail);
  else if (d == 4) op_apply1_kernel<4><<<blocks, 256, 0, s>>>(in, out, O, chiL, tail);
  else throw Error(-1, "op_apply1: physical dimension must be 2, 3 or 4");
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

}  // namespace tn
