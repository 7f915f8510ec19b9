// Blocked one-sided (Hestenes) Jacobi SVD for complex FP64, sm_100a.
//
// Replaces `svd(x, idx; cutoff, maxdim, mindim)` (reference src/tensors.jl:168-227: LAPACK zgesdd +
// host-side truncation).  Algorithm, for an m x n matrix with m >= n (otherwise the conjugate
// transpose is factorised and the roles of U and V are swapped):
//   Z = [W ; V] with W = M, V = I.  Columns are grouped into blocks of 32; a round-robin tournament
//   pairs the blocks; for every pair (64 columns)
//     1. G = P^H P           (batched strided ZGEMM on the DMMA pipe, split-K, tn_zgemm.cu)
//     2. G = J Lambda J^H    (two-sided cyclic Jacobi of the 64 x 64 Hermitian Gram matrix, one CTA,
//                             entirely in shared memory)
//     3. [W;V](:,pair) *= J  (batched strided ZGEMM, in place)
//   until every pair's Gram matrix is diagonal to sqrt(m)*eps.  Then sigma_j = ||W(:,j)||, U = W/sigma,
//   sorted on device; the truncation rank is computed on device with the reference's exact rule.
#include "tn_svd.cuh"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace tn {
void count_launch(int n);

constexpr int JB = 32;        // column block
constexpr int JP = 2 * JB;    // pair width
constexpr int LDS_ = JP + 1;  // padded shared leading dimension

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { /* a * conj(b) */ return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }

__device__ __forceinline__ void rr_pair(int n, int step, int k, int& p, int& q) {
  // circle method: player n-1 fixed, the other n-1 rotate
  int a, b;
  if (k == 0) { a = n - 1; b = step; }
  else { a = (step + k) % (n - 1); b = (step - k + (n - 1)) % (n - 1); }
  p = min(a, b); q = max(a, b);
}

__global__ void __launch_bounds__(256, 1) jacobi_evd64_kernel(const cplx* __restrict__ Gpart, int nsplit, long long split_stride,
                                                               cplx* __restrict__ Jout, double tol, unsigned long long* offmax, int max_inner) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx* G = reinterpret_cast<cplx*>(sm_raw);     // G[row*LDS_ + col]
  cplx* J = G + JP * LDS_;
  __shared__ double r_cs[JB];
  __shared__ cplx r_s[JB];
  __shared__ int r_p[JB], r_q[JB];
  __shared__ int rotated;
  __shared__ double red[8];
  const int tid = threadIdx.x;
  const cplx* gp = Gpart + (long long)blockIdx.x * JP * JP;
  for (int e = tid; e < JP * JP; e += 256) {
    int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    for (int s = 0; s < nsplit; ++s) { cplx v = gp[s * split_stride + e]; xr += v.x; xi += v.y; }
    G[row * LDS_ + col] = make_double2(xr, xi);
    J[row * LDS_ + col] = make_double2(row == col ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  // symmetrise (the two triangles come from different DMMA accumulation orders) and measure off-diagonals
  double mx = 0;
  for (int e = tid; e < JP * JP; e += 256) {
    int row = e % JP, col = e / JP;
    if (row < col) {
      cplx a = G[row * LDS_ + col], b = G[col * LDS_ + row];
      cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row * LDS_ + col] = h;
      G[col * LDS_ + row] = make_double2(h.x, -h.y);
      double dd = G[row * LDS_ + row].x * G[col * LDS_ + col].x;
      double off = sqrt(h.x * h.x + h.y * h.y);
      if (dd > 0) mx = fmax(mx, off / sqrt(dd));
      else if (off > 0) mx = fmax(mx, 1.0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmax(mx, red[i]);
  if (tid == 0) atomicMax(offmax, (unsigned long long)__double_as_longlong(mx));
  cplx* jo = Jout + (long long)blockIdx.x * JP * JP;
  if (mx <= tol) {   // already orthogonal: identity rotation
    for (int e = tid; e < JP * JP; e += 256) jo[e] = make_double2((e % JP) == (e / JP) ? 1.0 : 0.0, 0.0);
    return;
  }
  __syncthreads();
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < JP - 1; ++step) {
      if (tid < JB) {
        int p, q; rr_pair(JP, step, tid, p, q);
        double a = G[p * LDS_ + p].x, b = G[q * LDS_ + q].x;
        cplx c = G[p * LDS_ + q];
        double absc = hypot(c.x, c.y);
        double cs = 1.0; cplx s = make_double2(0, 0);
        if (absc > tol * sqrt(fabs(a * b)) && absc > 0) {
          double zeta = (b - a) / (2.0 * absc);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          cs = 1.0 / sqrt(1.0 + t * t);
          double sn = cs * t;
          s = make_double2(sn * c.x / absc, sn * c.y / absc);
          rotated = 1;
        }
        r_cs[tid] = cs; r_s[tid] = s; r_p[tid] = p; r_q[tid] = q;
      }
      __syncthreads();
      // phase 1: columns p,q of G and J:  [x y] <- [x y] * R,  R = [[cs, s], [-conj(s), cs]]
#pragma unroll 4
      for (int it = 0; it < (JB * JP * 2) / 256; ++it) {
        int item = tid + it * 256;
        int k = item >> 7, rem = item & 127, row = rem & 63;
        cplx* Mx = (rem >> 6) ? J : G;
        double cs = r_cs[k]; cplx s = r_s[k];
        if (s.x == 0.0 && s.y == 0.0) continue;
        int p = r_p[k], q = r_q[k];
        cplx x = Mx[row * LDS_ + p], y = Mx[row * LDS_ + q];
        cplx ys = cmulc(y, s), xs = cmul(x, s);
        Mx[row * LDS_ + p] = make_double2(cs * x.x - ys.x, cs * x.y - ys.y);
        Mx[row * LDS_ + q] = make_double2(xs.x + cs * y.x, xs.y + cs * y.y);
      }
      __syncthreads();
      // phase 2: rows p,q of G:  [x; y] <- R^H [x; y] = [cs*x - s*y ; conj(s)*x + cs*y]
#pragma unroll 4
      for (int it = 0; it < (JB * JP) / 256; ++it) {
        int item = tid + it * 256;
        int k = item >> 6, col = item & 63;
        double cs = r_cs[k]; cplx s = r_s[k];
        if (s.x == 0.0 && s.y == 0.0) continue;
        int p = r_p[k], q = r_q[k];
        cplx x = G[p * LDS_ + col], y = G[q * LDS_ + col];
        cplx sy = cmul(s, y), cx = cmulc(x, s);   // cx = x*conj(s)
        G[p * LDS_ + col] = make_double2(cs * x.x - sy.x, cs * x.y - sy.y);
        G[q * LDS_ + col] = make_double2(cx.x + cs * y.x, cx.y + cs * y.y);
      }
      __syncthreads();
      if (tid < JB) {
        cplx s = r_s[tid];
        if (!(s.x == 0.0 && s.y == 0.0)) {
          int p = r_p[tid], q = r_q[tid];
          G[p * LDS_ + q] = make_double2(0, 0);
          G[q * LDS_ + p] = make_double2(0, 0);
          G[p * LDS_ + p].y = 0; G[q * LDS_ + q].y = 0;
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  for (int e = tid; e < JP * JP; e += 256) jo[e] = J[(e % JP) * LDS_ + (e / JP)];
}

// ---- preparation / finalisation kernels -----------------------------------------------------
__global__ void __launch_bounds__(256) svd_init_kernel(const cplx* __restrict__ M, long long ld, int m, int n, int transposed,
                                                        cplx* __restrict__ Z, int rows, int ncols, int ncols_pad, int ldz) {
  // Z(0:rows, j) = (transposed ? conj(M)^T : M)(:, j) for j < ncols, 0 for padding; Z(rows + i, j) = delta_ij
  long long total = (long long)ldz * ncols_pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int r = (int)(e % ldz), c = (int)(e / ldz);
    cplx v = make_double2(0, 0);
    if (r < rows) {
      if (c < ncols) {
        if (!transposed) v = M[r + (long long)c * ld];
        else { cplx t = M[c + (long long)r * ld]; v = make_double2(t.x, -t.y); }
      }
    } else if (r - rows == c) v = make_double2(1.0, 0.0);
    Z[e] = v;
  }
}

__global__ void __launch_bounds__(128) colnorm2_kernel(const cplx* __restrict__ Z, int rows, int ldz, double* __restrict__ sig2) {
  const cplx* col = Z + (long long)blockIdx.x * ldz;
  double s = 0;
  for (int r = threadIdx.x; r < rows; r += 128) { cplx v = col[r]; s += v.x * v.x + v.y * v.y; }
  __shared__ double sh[4];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) sig2[blockIdx.x] = sh[0] + sh[1] + sh[2] + sh[3];
}

// single-CTA bitonic sort (descending) of the first `ncols` squared norms; padding sorts last.
__global__ void __launch_bounds__(1024) sort_trunc_kernel(const double* __restrict__ sig2, int ncols, int npow2, double* __restrict__ sig,
                                                           int* __restrict__ perm, int nsv, double cutoff, long long maxdim, long long mindim,
                                                           int* __restrict__ kout) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double* key = reinterpret_cast<double*>(sm_raw);
  int* val = reinterpret_cast<int*>(key + npow2);
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    key[i] = i < ncols ? sig2[i] : -1.0;
    val[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          bool desc = ((i & k) == 0);
          double a = key[i], b = key[ixj];
          int va = val[i], vb = val[ixj];
          // total order: larger key first, ties by smaller index (deterministic)
          bool a_first = (a > b) || (a == b && va < vb);
          if (desc ? !a_first : a_first) { key[i] = b; key[ixj] = a; val[i] = vb; val[ixj] = va; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) { sig[i] = sqrt(fmax(key[i], 0.0)); perm[i] = val[i]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    // reference rule, src/tensors.jl:201-215 (the S==0 test at :205 is a no-op)
    long long n = nsv;
    long long mind = mindim < n ? mindim : n;
    long long maxd = (maxdim == 0 || maxdim > n) ? n : maxdim;
    if (maxd == 0) maxd = 1;
    if (cutoff != 0.0) {
      double tot = 0;
      for (long long i = 0; i < n; ++i) tot += fmax(key[i], 0.0);
      double run = 0; long long keep = 0;
      for (long long i = n - 1; i >= 0; --i) {      // reverse(cumsum(reverse(S2))) / sum(S2)
        run += fmax(key[i], 0.0);
        if (run / tot > cutoff) { keep = i + 1; break; }
      }
      if (keep == 0) keep = 1;
      if (keep < maxd) maxd = keep;
    }
    kout[0] = (int)(maxd > mind ? maxd : mind);
  }
}

// out[r, j] = src[r, perm[j]] * scale_j                (transpose_out = 0)
// out[j, r] = conj(src[r, perm[j]]) * scale_j          (transpose_out = 1)
// scale: 0 -> 1, 1 -> sigma_j, 2 -> 1/sigma_j (0 if sigma_j == 0)
__global__ void __launch_bounds__(256) gather_kernel(const cplx* __restrict__ src, int lds, int rows, int k, const int* __restrict__ perm,
                                                      const double* __restrict__ sig, int scale_mode, int transpose_out,
                                                      cplx* __restrict__ out, long long ldo) {
  long long total = (long long)rows * k;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int r, j;
    if (!transpose_out) { r = (int)(e % rows); j = (int)(e / rows); }
    else { j = (int)(e % k); r = (int)(e / k); }
    cplx v = src[r + (long long)perm[j] * lds];
    double sc = 1.0;
    if (scale_mode == 1) sc = sig[j];
    else if (scale_mode == 2) sc = sig[j] > 0 ? 1.0 / sig[j] : 0.0;
    if (!transpose_out) out[r + (long long)j * ldo] = make_double2(v.x * sc, v.y * sc);
    else out[j + (long long)r * ldo] = make_double2(v.x * sc, -v.y * sc);
  }
}

// ---- host driver ------------------------------------------------------------------------------
template <class T>
static void ensure(T*& p, size_t& cap, size_t need, cudaStream_t s) {
  if (need <= cap) return;
  if (p) TN_CUDA(cudaFreeAsync(p, s));
  TN_CUDA(cudaMallocAsync((void**)&p, need * sizeof(T), s));
  cap = need;
}

static const int* pair_table(SvdWork& w, int nb, cudaStream_t s) {
  auto it = w.tables.find(nb);
  if (it != w.tables.end()) return it->second;
  int steps = nb - 1, np = nb / 2;
  std::vector<int> h((size_t)steps * np * 2);
  for (int st = 0; st < steps; ++st)
    for (int k = 0; k < np; ++k) {
      int a, b;
      if (nb == 2) { a = 0; b = 1; }
      else if (k == 0) { a = nb - 1; b = st; }
      else { a = (st + k) % (nb - 1); b = (st - k + (nb - 1)) % (nb - 1); }
      h[((size_t)st * np + k) * 2 + 0] = std::min(a, b);
      h[((size_t)st * np + k) * 2 + 1] = std::max(a, b);
    }
  int* d = nullptr;
  TN_CUDA(cudaMalloc((void**)&d, h.size() * sizeof(int)));
  TN_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TN_CUDA(cudaStreamSynchronize(s));
  w.tables[nb] = d;
  return d;
}

int svd_factor(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s) {
  TN_CHECK(m >= 1 && n >= 1, "svd: empty matrix");
  w.m = m; w.n = n;
  w.transposed = m < n;
  w.rows = w.transposed ? n : m;
  w.ncols = w.transposed ? m : n;
  w.nsv = w.ncols;
  w.ncols_pad = ((w.ncols + JP - 1) / JP) * JP;
  w.ldz = w.rows + w.ncols_pad;
  TN_CHECK(w.ncols_pad <= 8192, "svd: more than 8192 columns is not supported yet");
  const int nb = w.ncols_pad / JB, np = nb / 2, steps = nb - 1;
  ensure(w.Z, w.Z_cap, (size_t)w.ldz * w.ncols_pad, s);
  // split-K so that the Gram GEMMs fill the 148 SMs
  int ksplit = std::max(1, std::min(16, (2 * 148 + np - 1) / np));
  int kchunk = ((w.rows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (w.rows + kchunk - 1) / kchunk;
  ensure(w.Gpart, w.G_cap, (size_t)ksplit * np * JP * JP, s);
  ensure(w.J, w.J_cap, (size_t)np * JP * JP, s);
  if (w.s_cap < (size_t)w.ncols_pad) {
    if (w.sig) { TN_CUDA(cudaFreeAsync(w.sig, s)); TN_CUDA(cudaFreeAsync(w.perm, s)); TN_CUDA(cudaFreeAsync(w.sig2, s)); }
    TN_CUDA(cudaMallocAsync((void**)&w.sig, w.ncols_pad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.sig2, w.ncols_pad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.perm, w.ncols_pad * sizeof(int), s));
    w.s_cap = w.ncols_pad;
  }
  if (!w.offmax) { TN_CUDA(cudaMalloc((void**)&w.offmax, 8)); TN_CUDA(cudaMalloc((void**)&w.kout, 4)); }
  const int* tab = pair_table(w, nb, s);

  {
    long long total = (long long)w.ldz * w.ncols_pad;
    int blocks = (int)std::min<long long>(148 * 8, (total + 255) / 256);
    svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Z, w.rows, w.ncols, w.ncols_pad, w.ldz);
    TN_CUDA(cudaGetLastError());
    count_launch(1);
  }
  static bool evd_cfg = false;
  const int evd_smem = 2 * JP * LDS_ * (int)sizeof(cplx);
  if (!evd_cfg) { TN_CUDA(cudaFuncSetAttribute(jacobi_evd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, evd_smem)); evd_cfg = true; }

  const double tol = std::sqrt((double)w.rows) * 2.220446049250313e-16;
  const long long colblk = (long long)JB * w.ldz;
  const int max_sweeps = 60;
  // One inner sweep per visit gives the same number of outer sweeps as a full inner diagonalisation
  // (measured with tools/jacobi_emul.py) at 1/8 of the cost and with less rounding accumulated in V;
  // a single pair (n <= 64) has no outer parallelism to trade, so it is diagonalised fully.
  const int inner_sweeps = (np == 1) ? 12 : 1;
  w.sweeps = 0;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    TN_CUDA(cudaMemsetAsync(w.offmax, 0, 8, s));
    for (int st = 0; st < steps; ++st) {
      const int* tb = tab + (size_t)st * np * 2;
      Idx2 cols{JB, (long long)w.ldz, colblk, tb, 2};
      GemmDesc g{};
      g.M = JP; g.N = JP; g.K = w.rows;
      g.A = w.Z; g.am = cols; g.ak = idx1(1); g.conjA = 1;
      g.B = w.Z; g.bk = idx1(1); g.bn = cols; g.conjB = 0;
      g.C = w.Gpart; g.cm = idx1(1); g.cn = idx1(JP);
      g.alpha = make_double2(1, 0); g.beta = make_double2(0, 0);
      g.batch = np; g.bsA = 0; g.bsB = 0; g.bsC = (long long)JP * JP;
      g.ksplit = ksplit; g.kchunk = kchunk; g.ssC = (long long)np * JP * JP;
      zgemm_auto(g, s);
      jacobi_evd64_kernel<<<np, 256, evd_smem, s>>>(w.Gpart, ksplit, (long long)np * JP * JP, w.J, tol, w.offmax, inner_sweeps);
      TN_CUDA(cudaGetLastError());
      count_launch(1);
      GemmDesc a{};
      a.M = w.ldz; a.N = JP; a.K = JP;
      a.A = w.Z; a.am = idx1(1); a.ak = cols; a.conjA = 0;
      a.B = w.J; a.bk = idx1(1); a.bn = idx1(JP); a.conjB = 0;
      a.C = w.Z; a.cm = idx1(1); a.cn = cols;
      a.alpha = make_double2(1, 0); a.beta = make_double2(0, 0);
      a.batch = np; a.bsA = 0; a.bsB = (long long)JP * JP; a.bsC = 0;
      a.ksplit = 1; a.kchunk = JP; a.ssC = 0;
      zgemm_auto(a, s);
    }
    unsigned long long bits = 0;
    TN_CUDA(cudaMemcpyAsync(&bits, w.offmax, 8, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    double off; std::memcpy(&off, &bits, 8);
    w.last_off = off;
    w.sweeps = sweep + 1;
    if (off <= tol) break;
  }
  colnorm2_kernel<<<w.ncols_pad, 128, 0, s>>>(w.Z, w.rows, w.ldz, w.sig2);
  int npow2 = 64; while (npow2 < w.ncols_pad) npow2 <<= 1;
  static bool sort_cfg = false;
  if (!sort_cfg) { TN_CUDA(cudaFuncSetAttribute(sort_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12)); sort_cfg = true; }
  sort_trunc_kernel<<<1, 1024, npow2 * 12, s>>>(w.sig2, w.ncols_pad, npow2, w.sig, w.perm, w.nsv, tr.cutoff, tr.maxdim, tr.mindim, w.kout);
  TN_CUDA(cudaGetLastError());
  count_launch(2);
  int k = 0;
  TN_CUDA(cudaMemcpyAsync(&k, w.kout, 4, cudaMemcpyDeviceToHost, s));
  TN_CUDA(cudaStreamSynchronize(s));
  w.k = k;
  return k;
}

static void gather(const cplx* src, int lds, int rows, int k, const int* perm, const double* sig, int mode, int tr, cplx* out, long long ldo, cudaStream_t s) {
  long long total = (long long)rows * k;
  int blocks = std::max(1, (int)std::min<long long>(148 * 8, (total + 255) / 256));
  gather_kernel<<<blocks, 256, 0, s>>>(src, lds, rows, k, perm, sig, mode, tr, out, ldo);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

void svd_gather_U(SvdWork& w, cplx* U, long long ldu, bool times_S, cudaStream_t s) {
  if (!w.transposed) gather(w.Z, w.ldz, w.m, w.k, w.perm, w.sig, times_S ? 0 : 2, 0, U, ldu, s);          // W / sigma
  else gather(w.Z + w.rows, w.ldz, w.m, w.k, w.perm, w.sig, times_S ? 1 : 0, 0, U, ldu, s);                // V' (* sigma)
}
void svd_gather_Vh(SvdWork& w, cplx* Vh, long long ldv, bool times_S, cudaStream_t s) {
  if (!w.transposed) gather(w.Z + w.rows, w.ldz, w.n, w.k, w.perm, w.sig, times_S ? 1 : 0, 1, Vh, ldv, s); // conj(V)^T (* sigma)
  else gather(w.Z, w.ldz, w.n, w.k, w.perm, w.sig, times_S ? 0 : 2, 1, Vh, ldv, s);                        // conj(W'/sigma)^T
}
void svd_copy_S(SvdWork& w, double* S, cudaStream_t s) {
  TN_CUDA(cudaMemcpyAsync(S, w.sig, (size_t)w.k * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void svd_free(SvdWork& w) {
  if (w.Z) cudaFree(w.Z);
  if (w.Gpart) cudaFree(w.Gpart);
  if (w.J) cudaFree(w.J);
  if (w.sig) { cudaFree(w.sig); cudaFree(w.sig2); cudaFree(w.perm); }
  if (w.offmax) { cudaFree(w.offmax); cudaFree(w.kout); }
  for (auto& kv : w.tables) cudaFree(kv.second);
  w = SvdWork{};
}

}  // namespace tn
