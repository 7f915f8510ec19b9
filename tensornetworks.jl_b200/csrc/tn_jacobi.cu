// The two GEMM-shaped kernels of the blocked Jacobi step (tn_svd.cu), specialised for column-block pairs (2 x 32 columns).
//
// The truncated SVD (reference src/tensors.jl:168-227 -> LAPACK zgesdd) spends two thirds of a sweep in
//   G_p = P_p^H P_p                 (P_p = the 64 columns of pair p, rows x 64)         and
//   Z(:, pair p) <- Z(:, pair p) J_p   (rows x 64 times 64 x 64, in place).
// Through the general strided GEMM (tn_zgemm.cu) both ran at 20-23 TFLOP/s of the 33 the DMMA pipe sustains: with K = 64 the
// rotation has eight k-tiles per CTA tile, so the pipeline fill and the epilogue are a third of a tile's life, and the Gram
// block was computed in full (both triangles) from a panel that was loaded twice (once as A, once as B).  Here:
//   * jacobi_gram64: one staged copy of the panel serves both operands; only the 36 of 64 8x8 tiles on or above the diagonal
//     (in a cyclic-diagonal numbering that gives every warp nine tiles) are computed, the mirror image is written by the
//     epilogue, so the consumer still sees the full Hermitian block;
//   * jacobi_rot64: J_p stays resident in shared memory, a CTA walks several 64-row blocks of its pair with ONE continuous
//     cp.async ring (the loads of the next row block are in flight while the current one finishes: no drain / refill per
//     tile), two CTAs per SM.
// Fragment conventions are those of tn_zgemm.cu: complex elements interleaved in shared memory, [k][m] layout with a leading
// dimension == 2 (mod 8) 16-byte units, complex MAC = 4 real DMMA.8x8x4.
#include "tn_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace tn {
void count_launch(int n);

namespace {

constexpr int JB = 32;        // column block
constexpr int JP = 2 * JB;    // pair width

__device__ __forceinline__ void cpa16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ double neg_bits(double x) { return __longlong_as_double(__double_as_longlong(x) ^ (long long)0x8000000000000000ull); }   // integer pipe, not FP64
__device__ __forceinline__ void dmma2(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- G = P^H P --------------------------------------------------------------------------------------------------------------
constexpr int G_BK = 16, G_LD = JP + 2, G_STAGES = 4, G_THREADS = 128;
constexpr int G_STAGE = G_BK * G_LD;                    // complex elements per stage
constexpr int G_SMEM = G_STAGES * G_STAGE * 16;         // 67 584 B -> three CTAs per SM

// grid = (k-splits, pairs).  CTA (split, pair) accumulates rows [split * kchunk, (split + 1) * kchunk) of the pair's panel.
// Warp w owns the 8x8 tiles (i, (i + d) mod 8), i in {2w, 2w + 1}, d in 0..3, and (w, w + 4): every unordered pair of the eight
// 8-column groups exactly once.  atomic != 0: contributions are added with red.global.add.f64 (G zeroed by the launcher).
__global__ void __launch_bounds__(G_THREADS, 3) jacobi_gram64_kernel(const cplx* __restrict__ Z, long long ldz, int rows, int kchunk,
                                                                     const int* __restrict__ tab, cplx* __restrict__ G, int atomic,
                                                                     const int* __restrict__ skip) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Ps = reinterpret_cast<cplx*>(smem_raw);
  const int pair = blockIdx.y;
  if (skip != nullptr && skip[pair] != 0) return;     // e.g. third CholeskyQR pass of a panel that is orthonormal already
  const int k_begin = blockIdx.x * kchunk, k_end = min(rows, k_begin + kchunk);
  if (k_begin >= k_end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const long long c0 = (long long)tab[2 * pair] * JB, c1 = (long long)tab[2 * pair + 1] * JB;
  // load slots: row lk of the tile, columns lc + 8 i (i < 4: first block, i >= 4: second block); 16 consecutive threads read 256 contiguous bytes
  const int lk = tid & 15, lc = tid >> 4;
  const cplx* pa = Z + (c0 + lc) * ldz + k_begin + lk;
  const cplx* pb = Z + (c1 + lc) * ldz + k_begin + lk;
  int krow = k_begin + lk;
  auto load_stage = [&](int stage) {
    cplx* ps = Ps + stage * G_STAGE + lk * G_LD + lc;
    const bool p = krow < k_end;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cpa16(ps + 8 * i, p ? pa + (long long)(8 * i) * ldz : Z, p);
      cpa16(ps + 32 + 8 * i, p ? pb + (long long)(8 * i) * ldz : Z, p);
    }
    pa += G_BK; pb += G_BK; krow += G_BK;
  };

  double re[9][2], im[9][2];
#pragma unroll
  for (int i = 0; i < 9; ++i) { re[i][0] = re[i][1] = 0.0; im[i][0] = im[i][1] = 0.0; }

  const int ktiles = (k_end - k_begin + G_BK - 1) / G_BK;
#pragma unroll
  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s);
    cpa_commit();
  }
  // fragment columns of this warp: groups (2w + d) mod 8, d = 0..4, then w and w + 4
  int fcol[7];
#pragma unroll
  for (int d = 0; d < 5; ++d) fcol[d] = (((2 * warp + d) & 7) << 3) + g;
  fcol[5] = (warp << 3) + g; fcol[6] = ((warp + 4) << 3) + g;

  for (int kt = 0; kt < ktiles; ++kt) {
    cpa_wait<G_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + G_STAGES - 1;
      if (nk < ktiles) load_stage(nk % G_STAGES);
      cpa_commit();
    }
    const cplx* ps = Ps + (kt % G_STAGES) * G_STAGE;
#pragma unroll
    for (int kk = 0; kk < G_BK; kk += 4) {
      const cplx* row = ps + (kk + t) * G_LD;
      double fr[7], fi[7];
#pragma unroll
      for (int d = 0; d < 7; ++d) { cplx v = row[fcol[d]]; fr[d] = v.x; fi[d] = v.y; }
      // G_ij += conj(a) b :  re += ar br + ai bi,  im += ar bi - ai br
#pragma unroll
      for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int d = 0; d < 4; ++d) dmma2(re[ii * 4 + d], fr[ii], fr[ii + d]);
      dmma2(re[8], fr[5], fr[6]);
#pragma unroll
      for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int d = 0; d < 4; ++d) dmma2(im[ii * 4 + d], fr[ii], fi[ii + d]);
      dmma2(im[8], fr[5], fi[6]);
#pragma unroll
      for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int d = 0; d < 4; ++d) dmma2(re[ii * 4 + d], fi[ii], fi[ii + d]);
      dmma2(re[8], fi[5], fi[6]);
#pragma unroll
      for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int d = 0; d < 4; ++d) dmma2(im[ii * 4 + d], neg_bits(fi[ii]), fr[ii + d]);
      dmma2(im[8], neg_bits(fi[5]), fr[6]);
    }
  }
  cpa_wait<0>();

  // epilogue: tile (bi, bj) holds G[bi*8 + g][bj*8 + 2t + q]; off-diagonal tiles also write their mirror image
  cplx* Gp = G + (long long)pair * JP * JP;
  auto put = [&](int m, int n, double xr, double xi) {
    cplx* p = Gp + m + (long long)JP * n;
    if (atomic) { atomicAdd(&p->x, xr); atomicAdd(&p->y, xi); }
    else *p = make_double2(xr, xi);
  };
#pragma unroll
  for (int idx = 0; idx < 9; ++idx) {
    int bi, bj;
    if (idx < 8) { bi = 2 * warp + (idx >> 2); bj = (bi + (idx & 3)) & 7; }
    else { bi = warp; bj = warp + 4; }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int m = bi * 8 + g, n = bj * 8 + 2 * t + q;
      put(m, n, re[idx][q], im[idx][q]);
      if (bi != bj) put(n, m, re[idx][q], -im[idx][q]);
    }
  }
}

// ---- C = P^H T (projection coefficients of the block Gram-Schmidt QR) ---------------------------------------------------------
constexpr int X_BK = 8, X_LD = 2 * JP + 2, X_STAGES = 4, X_THREADS = 128;
constexpr int X_STAGE = X_BK * X_LD;
constexpr int X_SMEM = X_STAGES * X_STAGE * 16;         // 66 560 B -> three CTAs per SM

// grid = (k-splits, items), item = (problem b, tile i) = b * ntiles + i.  C_b(:, tile i) (+)= P_b^H T_b(:, tile i) over the rows of the split:
// P_b = column blocks atab[2 b], atab[2 b + 1] of A; the tile = column blocks ttab[2 item], ttab[2 item + 1] of T; the 64 x 64 result goes to
// C + cstride b + 64 ldc i with leading dimension ldc.  A stage holds the k-tile of the panel (columns 0..63) next to that of the tile (64..127);
// warp w owns the row groups 2w, 2w + 1 of the result and all eight column groups.
__global__ void __launch_bounds__(X_THREADS, 3) jacobi_cross64_kernel(const cplx* __restrict__ A, long long lda, const int* __restrict__ atab,
                                                                      const cplx* __restrict__ T, long long ldt, int rows, int kchunk,
                                                                      const int* __restrict__ ttab, int ntiles, cplx* __restrict__ C, long long ldc,
                                                                      long long cstride, int atomic) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Ps = reinterpret_cast<cplx*>(smem_raw);
  const int item = blockIdx.y, prob = item / ntiles, tile = item - prob * ntiles;
  const int k_begin = blockIdx.x * kchunk, k_end = min(rows, k_begin + kchunk);
  if (k_begin >= k_end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  // load slots: row lk of the k-tile, columns lc + 16 i of the 128-column stage (i < 4: panel, i >= 4: tile)
  const int lk = tid & 7, lc = tid >> 3;
  const cplx* sa[2]; const cplx* st[2];             // column lc of the two panel blocks / the two tile blocks; columns lc + 16 follow at 16 ld
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    sa[b] = A + ((long long)atab[2 * prob + b] * JB + lc) * lda + k_begin + lk;
    st[b] = T + ((long long)ttab[2 * item + b] * JB + lc) * ldt + k_begin + lk;
  }
  int krow = k_begin + lk;
  auto load_stage = [&](int stage) {
    cplx* ps = Ps + stage * X_STAGE + lk * X_LD + lc;
    const bool p = krow < k_end;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      cpa16(ps + 32 * b, p ? sa[b] : A, p);
      cpa16(ps + 32 * b + 16, p ? sa[b] + 16 * lda : A, p);
      cpa16(ps + JP + 32 * b, p ? st[b] : A, p);
      cpa16(ps + JP + 32 * b + 16, p ? st[b] + 16 * ldt : A, p);
      sa[b] += X_BK; st[b] += X_BK;
    }
    krow += X_BK;
  };
  double re[2][8][2], im[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { re[i][j][0] = re[i][j][1] = 0.0; im[i][j][0] = im[i][j][1] = 0.0; }
  const int ktiles = (k_end - k_begin + X_BK - 1) / X_BK;
#pragma unroll
  for (int s = 0; s < X_STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s);
    cpa_commit();
  }
  const int a_frag = warp * 16 + g, b_frag = JP + g;
  for (int kt = 0; kt < ktiles; ++kt) {
    cpa_wait<X_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + X_STAGES - 1;
      if (nk < ktiles) load_stage(nk % X_STAGES);
      cpa_commit();
    }
    const cplx* ps = Ps + (kt % X_STAGES) * X_STAGE;
#pragma unroll
    for (int kk = 0; kk < X_BK; kk += 4) {
      const cplx* row = ps + (kk + t) * X_LD;
      double ar[2], ai[2], nai[2], br[8], bi[8];
#pragma unroll
      for (int i = 0; i < 2; ++i) { cplx v = row[a_frag + 8 * i]; ar[i] = v.x; ai[i] = v.y; nai[i] = neg_bits(v.y); }
#pragma unroll
      for (int j = 0; j < 8; ++j) { cplx v = row[b_frag + 8 * j]; br[j] = v.x; bi[j] = v.y; }
      // C_ij += conj(a) b :  re += ar br + ai bi,  im += ar bi - ai br
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma2(re[i][j], ar[i], br[j]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma2(im[i][j], ar[i], bi[j]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma2(re[i][j], ai[i], bi[j]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma2(im[i][j], nai[i], br[j]);
    }
  }
  cpa_wait<0>();
  cplx* Cp = C + cstride * prob + (long long)JP * ldc * tile;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        cplx* p = Cp + (warp * 16 + i * 8 + g) + ldc * (j * 8 + 2 * t + q);
        if (atomic) { atomicAdd(&p->x, re[i][j][q]); atomicAdd(&p->y, im[i][j][q]); }
        else *p = make_double2(re[i][j][q], im[i][j][q]);
      }
}

// ---- Z(:, pair) <- Z(:, pair) J ---------------------------------------------------------------------------------------------
constexpr int R_BM = 64, R_BK = 8, R_LDA = R_BM + 2, R_LDB = JP + 2, R_STAGES = 5, R_THREADS = 128;
constexpr int R_ASTAGE = R_BK * R_LDA;                              // complex elements per stage
constexpr int R_SMEM = (JP * R_LDB + R_STAGES * R_ASTAGE) * 16;     // 109 824 B -> two CTAs per SM

// grid = (CTAs per item, items).  CTA (x, item) owns the 64-row blocks [nrb * x / gridDim.x, nrb * (x + 1) / gridDim.x) of the item and
// all 64 of its output columns.  2 x 2 warps, warp tile 32 x 32.  The k-tiles of consecutive row blocks form one stream v = 0, 1, ...
// (row block v / 8, columns 8 (v % 8) ..) through the cp.async ring; the accumulators are stored and cleared whenever v % 8 == 7.
//   UPDATE = false: Z(:, out blocks) = A(:, a blocks) * J          (Jacobi rotation / CholeskyQR apply: A = Z, a blocks = out blocks, in place)
//   UPDATE = true : Z(:, out blocks) -= A(:, a blocks) * J         (rank-64 trailing update of the block Gram-Schmidt QR: the panel A is
//                                                                  shared by all items, J = a 64 x 64 block of the coefficient matrix)
// a blocks of item i = atab[astride * (i / adiv)], atab[astride * (i / adiv) + 1]; out blocks = tab[2 i], tab[2 i + 1];
// J_i = J + jstride_hi * (i / adiv) + jstride * (i % adiv), element (k, n) at k + ldj n.
template <bool UPDATE>
__global__ void __launch_bounds__(R_THREADS, 2) jacobi_rot64_kernel(const cplx* A, long long lda, const int* __restrict__ atab, int astride, int adiv,
                                                                    cplx* Z, long long ldz, int rows, const int* __restrict__ tab,
                                                                    const cplx* __restrict__ J, long long ldj, long long jstride, long long jstride_hi,
                                                                    const int* __restrict__ skip) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Js = reinterpret_cast<cplx*>(smem_raw);       // [k][n], leading dimension R_LDB
  cplx* As = Js + JP * R_LDB;                         // ring of [k][m] tiles
  const int pair = blockIdx.y;
  if (skip != nullptr && skip[pair] != 0) return;     // Gram block already diagonal: J = identity
  const int nrb = (rows + R_BM - 1) / R_BM;
  const int rb0 = (int)((long long)nrb * blockIdx.x / gridDim.x), rb1 = (int)((long long)nrb * (blockIdx.x + 1) / gridDim.x);
  if (rb0 >= rb1) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  const int hi = pair / adiv, lo = pair - hi * adiv;
  const long long abase0 = (long long)atab[(long long)astride * hi] * JB * lda, abase1 = (long long)atab[(long long)astride * hi + 1] * JB * lda;
  const long long base0 = (long long)tab[2 * pair] * JB * ldz, base1 = (long long)tab[2 * pair + 1] * JB * ldz;

  // J (column-major 64 x 64 in global memory) -> Js[k][n]; part of the first cp.async group
  {
    const cplx* Jg = J + jstride_hi * hi + jstride * lo;
#pragma unroll 4
    for (int e = tid; e < JP * JP; e += R_THREADS) cpa16(Js + (e & 63) * R_LDB + (e >> 6), Jg + (e & 63) + ldj * (e >> 6), true);
  }
  // load slots: row lm of the block, columns lq + 2 i of the k-tile
  const int lm = tid & 63, lq = tid >> 6;
  const int nv = (rb1 - rb0) * 8;
  auto load_tile = [&](int v) {
    const int rb = rb0 + (v >> 3), kt = v & 7;
    const int row = rb * R_BM + lm;
    const bool p = row < rows;
    const cplx* src = A + (kt < 4 ? abase0 : abase1) + (long long)((kt & 3) * 8 + lq) * lda + row;
    cplx* dst = As + (v % R_STAGES) * R_ASTAGE + lq * R_LDA + lm;
#pragma unroll
    for (int i = 0; i < 4; ++i) cpa16(dst + 2 * i * R_LDA, p ? src + (long long)(2 * i) * lda : A, p);
  };

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }

#pragma unroll
  for (int s = 0; s < R_STAGES - 1; ++s) {
    if (s < nv) load_tile(s);
    cpa_commit();
  }
  const int a_frag = wm * 32 + g, b_frag = wn * 32 + g;

  for (int v = 0; v < nv; ++v) {
    cpa_wait<R_STAGES - 2>();
    __syncthreads();
    {
      const int nk = v + R_STAGES - 1;
      if (nk < nv) load_tile(nk);
      cpa_commit();
    }
    const int kt = v & 7;
    const cplx* as = As + (v % R_STAGES) * R_ASTAGE;
    const cplx* bs = Js + kt * R_BK * R_LDB;
#pragma unroll
    for (int kk = 0; kk < R_BK; kk += 4) {
      double ar[4], ai[4], br[4], bi[4], nbi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { cplx x = as[(kk + t) * R_LDA + a_frag + i * 8]; ar[i] = x.x; ai[i] = x.y; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { cplx x = bs[(kk + t) * R_LDB + b_frag + j * 8]; br[j] = x.x; bi[j] = x.y; nbi[j] = neg_bits(x.y); }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma2(cre[i][j], ar[i], br[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma2(cim[i][j], ar[i], bi[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma2(cre[i][j], ai[i], nbi[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma2(cim[i][j], ai[i], br[j]);
    }
    if (kt == 7) {
      // every k-tile of this row block has been read into shared memory: write the block (in place for the rotation) and start the next one
      const int rb = rb0 + (v >> 3);
      cplx* Cb = Z + (wn == 0 ? base0 : base1);          // the warp's 32 columns are exactly one column block
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = rb * R_BM + wm * 32 + i * 8 + g;
        if (UPDATE) {
          // read-modify-write: all eight loads of this row first (one L2 round trip per i, not one per element -- ncu showed the
          // element-by-element form stalled on the long scoreboard with the DMMA pipe 26 % active)
          cplx old[4][2];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) old[j][q] = (m < rows) ? __ldcg(Cb + (long long)(j * 8 + 2 * t + q) * ldz + m) : make_double2(0, 0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (m < rows) Cb[(long long)(j * 8 + 2 * t + q) * ldz + m] = make_double2(old[j][q].x - cre[i][j][q], old[j][q].y - cim[i][j][q]);
              cre[i][j][q] = 0.0; cim[i][j][q] = 0.0;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (m < rows) Cb[(long long)(j * 8 + 2 * t + q) * ldz + m] = make_double2(cre[i][j][q], cim[i][j][q]);
              cre[i][j][q] = 0.0; cim[i][j][q] = 0.0;
            }
        }
      }
    }
  }
  cpa_wait<0>();
}

int sm_count() {
  static int sms[32] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = sms[dev & 31];
  if (v == 0) { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) v = n; else v = 148; }
  return v;
}

}  // namespace

// G[p] (64 x 64, column-major, full Hermitian block) = P_p^H P_p over rows [0, rows) of Z, P_p = column blocks tab[2p], tab[2p+1] (32 columns each).
void jacobi_gram64(const cplx* Z, long long ldz, int rows, const int* tab, int npairs, cplx* G, int max_split, cudaStream_t s, const int* skip) {
  if (npairs <= 0 || rows <= 0) return;
  TN_CHECK(npairs <= 65535, "jacobi_gram64: too many pairs");
  static DeviceOnce cfg;
  cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_gram64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM)); });
  // about one wave of the 3 x SMs CTA slots; at least 64 rows per CTA
  int ksplit = std::max(1, std::min(std::min(max_split, rows / 64), (3 * sm_count()) / npairs));
  int kchunk = ((rows + ksplit - 1) / ksplit + G_BK - 1) / G_BK * G_BK;
  ksplit = (rows + kchunk - 1) / kchunk;
  if (ksplit > 1) TN_CUDA(cudaMemsetAsync(G, 0, (size_t)npairs * JP * JP * sizeof(cplx), s));
  jacobi_gram64_kernel<<<dim3(ksplit, npairs), G_THREADS, G_SMEM, s>>>(Z, ldz, rows, kchunk, tab, G, ksplit > 1 ? 1 : 0, skip);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// Z(0:rows, pair p) <- Z(0:rows, pair p) * J[p] in place (J[p] column-major 64 x 64); pairs with skip[p] != 0 are left alone.
void jacobi_rot64(cplx* Z, long long ldz, int rows, const int* tab, int npairs, const cplx* J, const int* skip, cudaStream_t s) {
  if (npairs <= 0 || rows <= 0) return;
  TN_CHECK(npairs <= 65535, "jacobi_rot64: too many pairs");
  static DeviceOnce cfg;
  cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_rot64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM)); });
  const int nrb = (rows + R_BM - 1) / R_BM;
  const int per_pair = std::max(1, std::min(nrb, (2 * sm_count()) / npairs));
  jacobi_rot64_kernel<false><<<dim3(per_pair, npairs), R_THREADS, R_SMEM, s>>>(Z, ldz, tab, 2, 1, Z, ldz, rows, tab, J, JP, 0, (long long)JP * JP, skip);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// Rank-64 updates T_b(0:rows, tile i) -= P_b(0:rows, :) * C_b(:, tile i) for b < nprob, i < ntiles: P_b = the two column blocks ptab[2 b], ptab[2 b + 1]
// of A (leading dimension lda), tile (b, i) = the column blocks ttab[2 (b ntiles + i)], ttab[.. + 1] of T (leading dimension ldt),
// C_b = C + cstride b: 64 x (64 ntiles) coefficients with leading dimension ldc.
void jacobi_update64(const cplx* A, long long lda, const int* ptab, cplx* T, long long ldt, int rows, const int* ttab, int nprob, int ntiles,
                     const cplx* C, long long ldc, long long cstride, cudaStream_t s) {
  const long long items = (long long)nprob * ntiles;
  if (items <= 0 || rows <= 0) return;
  TN_CHECK(items <= 65535, "jacobi_update64: too many tiles");
  static DeviceOnce cfg;
  cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_rot64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM)); });
  const int nrb = (rows + R_BM - 1) / R_BM;
  const int per_tile = std::max(1, std::min(nrb, (int)((2 * sm_count() + items - 1) / items)));
  jacobi_rot64_kernel<true><<<dim3(per_tile, (unsigned)items), R_THREADS, R_SMEM, s>>>(A, lda, ptab, 2, ntiles, T, ldt, rows, ttab, C, ldc, (long long)JP * ldc, cstride,
                                                                                          nullptr);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// Projection coefficients C_b(:, tile i) = P_b^H T_b(:, tile i) (64 x 64 each; tables and strides as jacobi_update64).  C must be zero on
// entry when the k range is split (contributions are added with red.global.add.f64).
void jacobi_cross64(const cplx* A, long long lda, const int* ptab, const cplx* T, long long ldt, int rows, const int* ttab, int nprob, int ntiles,
                    cplx* C, long long ldc, long long cstride, int max_split, cudaStream_t s) {
  const long long items = (long long)nprob * ntiles;
  if (items <= 0 || rows <= 0) return;
  TN_CHECK(items <= 65535, "jacobi_cross64: too many tiles");
  static DeviceOnce cfg;
  cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_cross64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM)); });
  int ksplit = std::max(1, std::min(std::min(max_split, rows / 64), (int)((3 * sm_count()) / items)));
  int kchunk = ((rows + ksplit - 1) / ksplit + X_BK - 1) / X_BK * X_BK;
  ksplit = (rows + kchunk - 1) / kchunk;
  jacobi_cross64_kernel<<<dim3(ksplit, (unsigned)items), X_THREADS, X_SMEM, s>>>(A, lda, ptab, T, ldt, rows, kchunk, ttab, ntiles, C, ldc, cstride, ksplit > 1 ? 1 : 0);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

}  // namespace tn
