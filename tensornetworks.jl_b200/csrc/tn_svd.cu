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
#include "tn_rounds.h"
#include <tuple>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>
#include <cstdlib>

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

// Hermitian storage of the Gram matrix: only G[r][c] with r <= c is kept up to date
__device__ __forceinline__ cplx herm_get(const cplx* G, int r, int c) {
  if (r < c) return G[r * LDS_ + c];
  cplx v = G[c * LDS_ + r];
  return make_double2(v.x, -v.y);
}
__device__ __forceinline__ void herm_put(cplx* G, int r, int c, cplx v) {
  if (r < c) G[r * LDS_ + c] = v;
  else G[c * LDS_ + r] = make_double2(v.x, -v.y);
}
constexpr int NBLK = JB * (JB + 1) / 2;   // 2x2-block pairs (a <= b) of the 32 rotations of a step

// One CTA (1024 threads) per column-block pair.  Per rotation step: 32 threads compute the 32 disjoint
// rotations, then every thread applies both the row- and the column-rotation to one 2x2 block of G
// (G' = R_a^H G_ab R_b) and the column rotation to two (row, pair) items of J: two barriers per step.
constexpr int EVD_THREADS = 1024;
__global__ void __launch_bounds__(EVD_THREADS, 1) jacobi_evd64_kernel(const cplx* __restrict__ Gpart, int nsplit, long long split_stride,
                                                                       cplx* __restrict__ Jout, double tol, unsigned long long* offmax, int max_inner, int nact,
                                                                       int* __restrict__ skip_out) {
  // nact (even, <= 64): only the leading nact columns of the pair can be non-zero (a single zero-padded pair); the
  // round-robin then runs over nact columns (nact-1 steps) instead of 64
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx* G = reinterpret_cast<cplx*>(sm_raw);     // G[row*LDS_ + col]
  cplx* J = G + JP * LDS_;
  __shared__ double r_cs[JB];
  __shared__ cplx r_s[JB];
  __shared__ int r_p[JB], r_q[JB];
  __shared__ int rotated;
  __shared__ double red[32];
  __shared__ unsigned char blk_a[NBLK], blk_b[NBLK];
  const int tid = threadIdx.x;
  { int a = tid >> 5, b = tid & 31; if (a <= b) { int i = a * JB - a * (a - 1) / 2 + (b - a); blk_a[i] = (unsigned char)a; blk_b[i] = (unsigned char)b; } }
  const cplx* gp = Gpart + (long long)blockIdx.x * JP * JP;
  for (int e = tid; e < JP * JP; e += EVD_THREADS) {
    int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    for (int s = 0; s < nsplit; ++s) { cplx v = gp[s * split_stride + e]; xr += v.x; xi += v.y; }
    G[row * LDS_ + col] = make_double2(xr, xi);
    J[row * LDS_ + col] = make_double2(row == col ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  // symmetrise (the two triangles come from different DMMA accumulation orders) and measure off-diagonals
  double mx = 0;
  for (int e = tid; e < JP * JP; e += EVD_THREADS) {
    int row = e % JP, col = e / JP;
    if (row < col) {
      cplx a = G[row * LDS_ + col], b = G[col * LDS_ + row];
      cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row * LDS_ + col] = h;
      G[col * LDS_ + row] = make_double2(h.x, -h.y);
      double dd = G[row * LDS_ + row].x * G[col * LDS_ + col].x;
      double off2 = h.x * h.x + h.y * h.y;
      if (dd > 0) mx = fmax(mx, off2 / dd);
      else if (off2 > 0) mx = fmax(mx, 1.0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < EVD_THREADS / 32; ++i) mx = fmax(mx, red[i]);
  mx = sqrt(mx);
  if (tid == 0) atomicMax(offmax, (unsigned long long)__double_as_longlong(mx));
  cplx* jo = Jout + (long long)blockIdx.x * JP * JP;
  if (tid == 0 && skip_out) skip_out[blockIdx.x] = (mx <= tol) ? 1 : 0;   // the rotation GEMM skips converged pairs
  if (mx <= tol) {   // already orthogonal: identity rotation
    for (int e = tid; e < JP * JP; e += EVD_THREADS) jo[e] = make_double2((e % JP) == (e / JP) ? 1.0 : 0.0, 0.0);
    return;
  }
  const double tol2 = tol * tol;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int step = 0; step < nact - 1; ++step) {
      if (tid < JB) {
        int p = 0, q = 0;
        if (tid < nact / 2) rr_pair(nact, step, tid, p, q);
        double a = G[p * LDS_ + p].x, b = G[q * LDS_ + q].x;
        cplx c = G[p * LDS_ + q];
        double absc2 = c.x * c.x + c.y * c.y;
        double cs = 1.0; cplx s = make_double2(0, 0);
        if (tid < nact / 2 && absc2 > tol2 * fabs(a * b) && absc2 > 0) {
          // t = sign(dl)|c| / (|dl| + sqrt(dl^2 + |c|^2)), dl = (b-a)/2;  with h = |dl| + sqrt(dl^2+|c|^2):
          // cs = h / sqrt(h^2 + |c|^2),  s = sign(dl) c / sqrt(h^2 + |c|^2)   (one sqrt, one rsqrt, no division:
          // FP64 latency on this part is ~50 cycles per dependent op, so the length of this chain sets the step time)
          double dl = 0.5 * (b - a);
          double h = fabs(dl) + sqrt(fma(dl, dl, absc2));
          double q = rsqrt(fma(h, h, absc2));
          cs = h * q;
          double sg = dl >= 0 ? q : -q;
          s = make_double2(sg * c.x, sg * c.y);
          rotated = 1;
        }
        r_cs[tid] = cs; r_s[tid] = s; r_p[tid] = p; r_q[tid] = q;
      }
      __syncthreads();
      if (tid < NBLK) {
        // G block (pair a rows, pair b cols), a <= b only: G' = R_a^H (G R_b),  R = [[cs, s], [-conj(s), cs]].  G is
        // Hermitian, so only the blocks on or above the block diagonal are updated (528 of 1024: the FP64 issue rate of
        // the one SM bounds this kernel); element (r, c) lives at G[min][max], conjugated when r > c.
        const int a = blk_a[tid], b = blk_b[tid];
        double ca = r_cs[a], cb = r_cs[b]; cplx sa = r_s[a], sb = r_s[b];
        bool ra = !(sa.x == 0.0 && sa.y == 0.0), rb = !(sb.x == 0.0 && sb.y == 0.0);
        if ((ra || rb) && b < nact / 2) {
          int pa = r_p[a], qa = r_q[a], pb = r_p[b], qb = r_q[b];
          cplx g00, g01, g10, g11;
          if (a == b) {
            g00 = make_double2(G[pa * LDS_ + pa].x, 0.0); g11 = make_double2(G[qa * LDS_ + qa].x, 0.0);
            g01 = G[pa * LDS_ + qa]; g10 = make_double2(g01.x, -g01.y);
          } else {
            g00 = herm_get(G, pa, pb); g01 = herm_get(G, pa, qb); g10 = herm_get(G, qa, pb); g11 = herm_get(G, qa, qb);
          }
          // T = G R_b
          cplx t00, t01, t10, t11;
          { cplx y = cmulc(g01, sb), x = cmul(g00, sb); t00 = make_double2(cb * g00.x - y.x, cb * g00.y - y.y); t01 = make_double2(x.x + cb * g01.x, x.y + cb * g01.y); }
          { cplx y = cmulc(g11, sb), x = cmul(g10, sb); t10 = make_double2(cb * g10.x - y.x, cb * g10.y - y.y); t11 = make_double2(x.x + cb * g11.x, x.y + cb * g11.y); }
          // G' = R_a^H T : row0 = ca*T0 - sa*T1 ; row1 = conj(sa)*T0 + ca*T1
          cplx n00, n01, n10, n11;
          { cplx y = cmul(sa, t10), x = cmulc(t00, sa); n00 = make_double2(ca * t00.x - y.x, ca * t00.y - y.y); n10 = make_double2(x.x + ca * t10.x, x.y + ca * t10.y); }
          { cplx y = cmul(sa, t11), x = cmulc(t01, sa); n01 = make_double2(ca * t01.x - y.x, ca * t01.y - y.y); n11 = make_double2(x.x + ca * t11.x, x.y + ca * t11.y); }
          if (a == b) {
            G[pa * LDS_ + pa] = make_double2(n00.x, 0.0); G[qa * LDS_ + qa] = make_double2(n11.x, 0.0);
            G[pa * LDS_ + qa] = n01;
          } else {
            herm_put(G, pa, pb, n00); herm_put(G, pa, qb, n01); herm_put(G, qa, pb, n10); herm_put(G, qa, qb, n11);
          }
        }
      }
      else {
        // J columns: 64 rows x 32 pairs = 2048 items, handled by the 496 threads that have no G block (about four
        // items each, the same FP64 work as one G block: the two halves of the CTA finish the step together)
        for (int item = tid - NBLK; item < JP * JB; item += EVD_THREADS - NBLK) {
          int k = item >> 6, row = item & 63;
          double cs = r_cs[k]; cplx s = r_s[k];
          if (s.x == 0.0 && s.y == 0.0) continue;
          int p = r_p[k], q = r_q[k];
          cplx x = J[row * LDS_ + p], y = J[row * LDS_ + q];
          cplx ys = cmulc(y, s), xs = cmul(x, s);
          J[row * LDS_ + p] = make_double2(cs * x.x - ys.x, cs * x.y - ys.y);
          J[row * LDS_ + q] = make_double2(xs.x + cs * y.x, xs.y + cs * y.y);
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
  }
  __syncthreads();
  for (int e = tid; e < JP * JP; e += EVD_THREADS) jo[e] = J[(e % JP) * LDS_ + (e / JP)];
}


// ---- second-generation pair EVD ---------------------------------------------------------------------------------------
// Same mathematics as jacobi_evd64_kernel (parallel-order two-sided Jacobi on the 64 x 64 Hermitian Gram block, rotations
// accumulated in J), reorganised around the measured bottleneck: the old kernel spent 1.8 us per rotation step, 2n of which
// are strictly sequential per outer sweep.  Here
//   * the J accumulation is taken off the critical path: J' = J R_step is independent row by row, so a second group of warps
//     owns 8 rows of J each and replays the rotation parameters from a shared ring (one entry per step, written by the G
//     group) -- it never takes part in the per-step barriers, only waits for the step counter;
//   * the G group is 288 threads (two 2x2 blocks of the upper block triangle per thread, independent instruction streams)
//     synchronised with a named barrier that the J warps do not join;
//   * all index tables of a step (pair -> columns) are built once per kernel, not per step.
constexpr int EVD2_G = 288;                    // threads updating G (9 warps)
constexpr int EVD2_J = 256;                    // threads accumulating J (8 warps x 8 rows)
constexpr int EVD2_THREADS = EVD2_G + EVD2_J;
constexpr int EVD2_RING = (JP - 1) * JB;       // rotation parameters of one inner sweep
constexpr size_t EVD2_SMEM = 2 * (size_t)JP * LDS_ * sizeof(cplx) + (size_t)EVD2_RING * (sizeof(double) + sizeof(cplx)) + (size_t)(JP - 1) * JP;

__global__ void __launch_bounds__(EVD2_THREADS, 1) jacobi_evd64v2_kernel(const cplx* __restrict__ Gpart, int nsplit, long long split_stride,
                                                                         cplx* __restrict__ Jout, double tol, unsigned long long* offmax, int max_inner, int nact,
                                                                         int* __restrict__ skip_out) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx* G = reinterpret_cast<cplx*>(sm_raw);     // G[row*LDS_ + col], upper triangle (row <= col) kept up to date
  cplx* J = G + JP * LDS_;
  double* ring_c = reinterpret_cast<double*>(J + JP * LDS_);            // [step][k] cosine
  cplx* ring_s = reinterpret_cast<cplx*>(ring_c + EVD2_RING);           // [step][k] sine (complex)
  unsigned char* pq = reinterpret_cast<unsigned char*>(ring_s + EVD2_RING);   // [step][2k], [step][2k+1]: the columns of pair k
  __shared__ double red[EVD2_THREADS / 32];
  __shared__ unsigned char blk_a[NBLK], blk_b[NBLK];
  __shared__ volatile int ready;               // steps of the current inner sweep whose parameters are in the ring
  __shared__ int rotated;
  const int tid = threadIdx.x;
  const int nsteps = nact - 1, npairs = nact / 2;
  for (int e = tid; e < JB * JB; e += EVD2_THREADS) {
    int a = e >> 5, b = e & 31;
    if (a <= b) { int i = a * JB - a * (a - 1) / 2 + (b - a); blk_a[i] = (unsigned char)a; blk_b[i] = (unsigned char)b; }
  }
  for (int e = tid; e < nsteps * JB; e += EVD2_THREADS) {
    int st = e / JB, k = e - st * JB, p = 0, q = 0;
    if (k < npairs) rr_pair(nact, st, k, p, q);
    pq[st * JP + 2 * k] = (unsigned char)p; pq[st * JP + 2 * k + 1] = (unsigned char)q;
  }
  if (tid == 0) { ready = 0; rotated = 0; }
  const cplx* gp = Gpart + (long long)blockIdx.x * JP * JP;
  for (int e = tid; e < JP * JP; e += EVD2_THREADS) {
    int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    for (int s = 0; s < nsplit; ++s) { cplx v = gp[s * split_stride + e]; xr += v.x; xi += v.y; }
    G[row * LDS_ + col] = make_double2(xr, xi);
    J[row * LDS_ + col] = make_double2(row == col ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  // symmetrise (the two triangles come from different DMMA accumulation orders) and measure the off-diagonals
  double mx = 0;
  for (int e = tid; e < JP * JP; e += EVD2_THREADS) {
    int row = e % JP, col = e / JP;
    if (row < col) {
      cplx a = G[row * LDS_ + col], b = G[col * LDS_ + row];
      cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row * LDS_ + col] = h;
      double dd = G[row * LDS_ + row].x * G[col * LDS_ + col].x;
      double off2 = h.x * h.x + h.y * h.y;
      if (dd > 0) mx = fmax(mx, off2 / dd);
      else if (off2 > 0) mx = fmax(mx, 1.0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < EVD2_THREADS / 32; ++i) mx = fmax(mx, red[i]);
  mx = sqrt(mx);
  if (tid == 0) atomicMax(offmax, (unsigned long long)__double_as_longlong(mx));
  cplx* jo = Jout + (long long)blockIdx.x * JP * JP;
  if (tid == 0 && skip_out) skip_out[blockIdx.x] = (mx <= tol) ? 1 : 0;   // the rotation GEMM skips converged pairs
  if (mx <= tol) {   // already orthogonal: identity rotation
    for (int e = tid; e < JP * JP; e += EVD2_THREADS) jo[e] = make_double2((e % JP) == (e / JP) ? 1.0 : 0.0, 0.0);
    return;
  }
  const double tol2 = tol * tol;
  const bool ggroup = tid < EVD2_G;
  const int jt = tid - EVD2_G, jwarp = jt >> 5, lane = tid & 31;
  for (int sweep = 0; sweep < max_inner; ++sweep) {
    if (ggroup) {
      for (int step = 0; step < nsteps; ++step) {
        const unsigned char* pqs = pq + step * JP;
        if (tid < JB) {
          const int p = pqs[2 * tid], q = pqs[2 * tid + 1];
          double a = G[p * LDS_ + p].x, b = G[q * LDS_ + q].x;
          cplx c = G[p * LDS_ + q];
          double absc2 = c.x * c.x + c.y * c.y;
          double cs = 1.0; cplx s = make_double2(0, 0);
          if (tid < npairs && absc2 > tol2 * fabs(a * b) && absc2 > 0) {
            double dl = 0.5 * (b - a);
            double h = fabs(dl) + sqrt(fma(dl, dl, absc2));
            double qq = rsqrt(fma(h, h, absc2));
            cs = h * qq;
            double sg = dl >= 0 ? qq : -qq;
            s = make_double2(sg * c.x, sg * c.y);
            rotated = 1;
          }
          ring_c[step * JB + tid] = cs; ring_s[step * JB + tid] = s;
          __syncwarp();
          if (tid == 0) { __threadfence_block(); ready = step + 1; }
        }
        asm volatile("bar.sync 1, %0;" :: "n"(EVD2_G) : "memory");
        const double* rc = ring_c + step * JB; const cplx* rs = ring_s + step * JB;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int blk = tid + it * EVD2_G;
          if (blk < NBLK) {
            // G block (pair a rows, pair b cols), a <= b only: G' = R_a^H (G R_b),  R = [[cs, s], [-conj(s), cs]]
            const int a = blk_a[blk], b = blk_b[blk];
            const double ca = rc[a], cb = rc[b]; const cplx sa = rs[a], sb = rs[b];
            const bool ra = !(sa.x == 0.0 && sa.y == 0.0), rb = !(sb.x == 0.0 && sb.y == 0.0);
            if ((ra || rb) && b < npairs) {
              const int pa = pqs[2 * a], qa = pqs[2 * a + 1], pb = pqs[2 * b], qb = pqs[2 * b + 1];
              cplx g00, g01, g10, g11;
              if (a == b) {
                g00 = make_double2(G[pa * LDS_ + pa].x, 0.0); g11 = make_double2(G[qa * LDS_ + qa].x, 0.0);
                g01 = G[pa * LDS_ + qa]; g10 = make_double2(g01.x, -g01.y);
              } else {
                g00 = herm_get(G, pa, pb); g01 = herm_get(G, pa, qb); g10 = herm_get(G, qa, pb); g11 = herm_get(G, qa, qb);
              }
              cplx t00, t01, t10, t11;
              { cplx y = cmulc(g01, sb), x = cmul(g00, sb); t00 = make_double2(cb * g00.x - y.x, cb * g00.y - y.y); t01 = make_double2(x.x + cb * g01.x, x.y + cb * g01.y); }
              { cplx y = cmulc(g11, sb), x = cmul(g10, sb); t10 = make_double2(cb * g10.x - y.x, cb * g10.y - y.y); t11 = make_double2(x.x + cb * g11.x, x.y + cb * g11.y); }
              cplx n00, n01, n10, n11;
              { cplx y = cmul(sa, t10), x = cmulc(t00, sa); n00 = make_double2(ca * t00.x - y.x, ca * t00.y - y.y); n10 = make_double2(x.x + ca * t10.x, x.y + ca * t10.y); }
              { cplx y = cmul(sa, t11), x = cmulc(t01, sa); n01 = make_double2(ca * t01.x - y.x, ca * t01.y - y.y); n11 = make_double2(x.x + ca * t11.x, x.y + ca * t11.y); }
              if (a == b) {
                G[pa * LDS_ + pa] = make_double2(n00.x, 0.0); G[qa * LDS_ + qa] = make_double2(n11.x, 0.0);
                G[pa * LDS_ + qa] = n01;
              } else {
                herm_put(G, pa, pb, n00); herm_put(G, pa, qb, n01); herm_put(G, qa, pb, n10); herm_put(G, qa, qb, n11);
              }
            }
          }
        }
        asm volatile("bar.sync 1, %0;" :: "n"(EVD2_G) : "memory");
      }
    } else {
      // J warps: rows [8 jwarp, 8 jwarp + 8), lane = pair; they trail the G group by polling the step counter
      for (int step = 0; step < nsteps; ++step) {
        if (lane == 0) { while (ready <= step) { } }
        __syncwarp();
        const double cs = ring_c[step * JB + lane]; const cplx s = ring_s[step * JB + lane];
        if (!(s.x == 0.0 && s.y == 0.0)) {
          const int p = pq[step * JP + 2 * lane], q = pq[step * JP + 2 * lane + 1];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int row = jwarp * 8 + r;
            cplx x = J[row * LDS_ + p], y = J[row * LDS_ + q];
            cplx ys = cmulc(y, s), xs = cmul(x, s);
            J[row * LDS_ + p] = make_double2(cs * x.x - ys.x, cs * x.y - ys.y);
            J[row * LDS_ + q] = make_double2(xs.x + cs * y.x, xs.y + cs * y.y);
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();                 // both groups have finished this inner sweep: the ring can be reused
    const int any = rotated;
    __syncthreads();
    if (tid == 0) { ready = 0; rotated = 0; }
    __syncthreads();
    if (!any) break;
  }
  for (int e = tid; e < JP * JP; e += EVD2_THREADS) jo[e] = J[(e % JP) * LDS_ + (e / JP)];
}


// ---- what did NOT pay (measured on a B200, profiles/r02_evd_experiments.md) --------------------------------------------------
// A third kernel moved the critical path to a leader warp (the 32 elements the next step needs are updated in registers and the
// next parameters derived at once, FP32-seeded and re-normalised in FP64, while bulk warps update the other blocks) and kept J in
// registers (one lane shuffle per value and step instead of 128 KB of shared-memory traffic).  Bit-for-bit parity, but no gain:
// 105 us per 63-step launch against 111 us -- the step is bound by instruction issue (~6800 warp instructions per step over four
// schedulers; the select / move overhead of the shuffled J costs what its shared-memory traffic did), ~20 us of every launch are
// fixed cost (load, symmetrise, tables, store), and polling J warps slowed the launch 4.5x until they joined the step barrier.
// The kernel was removed again; the numbers are in the profile note.

// ---- small matrices: the whole factorisation in one CTA -------------------------------------------------------------------
// rotation R = [[cs, s], [-conj(s), cs]] that annihilates the off-diagonal c between the diagonal entries a, b: the angle is seeded
// in FP32 (scale-free ratio |dl| / |c| after the even part of the exponent of |c|^2 is removed) and cs^2 + |s|^2 = 1 is restored in
// FP64 by one Newton step.  Unitary to double precision; the angle is accurate to ~1e-7 relative, which leaves Jacobi's quadratic
// convergence intact down to 1e-7 * off.  ~10 dependent FP64 instructions instead of ~21 plus two MUFU expansions.
__device__ __forceinline__ void jacobi_params(double a, double b, cplx c, double absc2, double& cs, cplx& s) {
  const double dl = 0.5 * (b - a);
  const int ex = ((__double2hiint(absc2) >> 20) & 0x7ff) - 1023;
  const int h = ex >> 1;
  const double sc = __hiloint2double((1023 - h) << 20, 0);
  const float cx = (float)(c.x * sc), cy = (float)(c.y * sc);
  const float cn2 = fmaf(cx, cx, cy * cy);
  const float rc = rsqrtf(cn2);                                  // 1 / |cn|
  const float z = fminf(fabsf((float)(dl * sc)) * rc, 1e30f);    // |zeta| = |dl| / |c|
  const float t = z < 1e4f ? __fdividef(1.0f, z + sqrtf(fmaf(z, z, 1.0f))) : __fdividef(0.5f, z);
  const float c0 = rsqrtf(fmaf(t, t, 1.0f));
  const float m0 = (dl >= 0 ? t : -t) * c0 * rc;                 // s0 = m0 * cn
  const double csd = (double)c0, sx = (double)(m0 * cx), sy = (double)(m0 * cy);
  const double e = fma(csd, csd, fma(sx, sx, sy * sy)) - 1.0;    // |e| ~ 1e-7
  const double f = fma(e, fma(e, 0.375, -0.5), 1.0);             // (1 + e)^(-1/2) up to O(e^3)
  cs = csd * f; s = make_double2(sx * f, sy * f);
}

// One-sided (Hestenes) Jacobi on a matrix with at most 64 columns and 128 rows, entirely in shared memory: one warp per column
// pair of the round-robin step (lane = row), three dot products by warp shuffles, the rotation applied to the two columns of W
// and of V, one block barrier per step.  Then sigma_j = ||W_j||, the sort and the reference's truncation rule (tensors.jl:201-215),
// and W / V go back to the ordinary workspace layout Z = [W ; V] so that the gathers do not change.  Replaces ~15 launches
// (init, 3 x (memset, Gram GEMM, pair EVD, rotation GEMM), column norms, sort) of the general path: the bond of the reference's own
// example (examples/dmrg.jl, 22 x 22) spent 450 of its 920 us there.
constexpr int SMALL_MAX_ROWS = 128;
__global__ void __launch_bounds__(1024, 1) small_svd_kernel(const cplx* __restrict__ M, long long ld, int transposed, int rows, int ncols,
                                                            cplx* __restrict__ Z, int ldz, double* __restrict__ sig, int* __restrict__ perm,
                                                            int nsv, double cutoff, long long maxdim, long long mindim, int* __restrict__ kout,
                                                            int* __restrict__ sweeps_out, double tol) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int nact = (ncols + 1) & ~1, npairs = nact / 2;
  cplx* W = reinterpret_cast<cplx*>(sm_raw);                  // rows x nact, column-major, leading dimension rows
  cplx* V = W + (size_t)rows * nact;                          // nact x nact
  __shared__ double s2[JP];
  __shared__ unsigned long long offbits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < rows * nact; e += blockDim.x) {
    const int r = e % rows, c = e / rows;
    cplx v = make_double2(0, 0);
    if (c < ncols) {
      if (!transposed) v = M[r + (long long)c * ld];
      else { const cplx t = M[c + (long long)r * ld]; v = make_double2(t.x, -t.y); }
    }
    W[e] = v;
  }
  for (int e = tid; e < nact * nact; e += blockDim.x) V[e] = make_double2((e % nact) == (e / nact) ? 1.0 : 0.0, 0.0);
  if (tid == 0) offbits = 0ull;
  __syncthreads();
  const double tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < 60; ++sweep) {
    for (int step = 0; step < nact - 1; ++step) {
      if (warp < npairs) {
        int p, q;
        rr_pair(nact, step, warp, p, q);
        cplx* wp = W + (size_t)p * rows; cplx* wq = W + (size_t)q * rows;
        cplx xp[4], xq[4];
        double a = 0, b = 0, cr = 0, ci = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = lane + 32 * i;
          if (r < rows) {
            xp[i] = wp[r]; xq[i] = wq[r];
            a = fma(xp[i].x, xp[i].x, fma(xp[i].y, xp[i].y, a));
            b = fma(xq[i].x, xq[i].x, fma(xq[i].y, xq[i].y, b));
            cr = fma(xp[i].x, xq[i].x, fma(xp[i].y, xq[i].y, cr));       // conj(xp) * xq
            ci = fma(xp[i].x, xq[i].y, fma(-xp[i].y, xq[i].x, ci));
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
          cr += __shfl_xor_sync(0xffffffffu, cr, o); ci += __shfl_xor_sync(0xffffffffu, ci, o);
        }
        const double absc2 = fma(cr, cr, ci * ci);
        if (absc2 > tol2 * a * b && absc2 > 1e-280) {
          if (lane == 0) atomicMax(&offbits, (unsigned long long)__double_as_longlong(absc2 / (a * b)));
          double cs; cplx s;
          jacobi_params(a, b, make_double2(cr, ci), absc2, cs, s);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            if (r < rows) {
              const cplx ys = cmulc(xq[i], s), xs = cmul(xp[i], s);
              wp[r] = make_double2(cs * xp[i].x - ys.x, cs * xp[i].y - ys.y);
              wq[r] = make_double2(xs.x + cs * xq[i].x, xs.y + cs * xq[i].y);
            }
          }
          cplx* vp = V + (size_t)p * nact; cplx* vq = V + (size_t)q * nact;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = lane + 32 * i;
            if (r < nact) {
              const cplx x = vp[r], y = vq[r];
              const cplx ys = cmulc(y, s), xs = cmul(x, s);
              vp[r] = make_double2(cs * x.x - ys.x, cs * x.y - ys.y);
              vq[r] = make_double2(xs.x + cs * y.x, xs.y + cs * y.y);
            }
          }
        }
      }
      __syncthreads();
    }
    const unsigned long long ob = offbits;
    __syncthreads();
    if (tid == 0) offbits = 0ull;
    __syncthreads();
    if (ob == 0ull) { ++sweep; break; }             // no pair was above the threshold in this sweep
  }
  // squared column norms, descending order (ties by index; padding columns last), truncation rank
  if (tid < JP) {
    double v = -1.0;
    if (tid < ncols) { v = 0; const cplx* wc = W + (size_t)tid * rows; for (int r = 0; r < rows; ++r) v = fma(wc[r].x, wc[r].x, fma(wc[r].y, wc[r].y, v)); }
    s2[tid] = v;
  }
  __syncthreads();
  if (tid < JP) {
    const double v = s2[tid];
    int rank = 0;
    for (int i = 0; i < JP; ++i) { const double u = s2[i]; rank += (u > v || (u == v && i < tid)) ? 1 : 0; }
    sig[rank] = sqrt(fmax(v, 0.0)); perm[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    // reference rule, src/tensors.jl:201-215 (the S==0 test at :205 is a no-op)
    const long long n = nsv;
    const long long mind = mindim < n ? mindim : n;
    long long maxd = (maxdim == 0 || maxdim > n) ? n : maxdim;
    if (maxd == 0) maxd = 1;
    if (cutoff != 0.0) {
      double tot = 0;
      for (long long i = 0; i < n; ++i) tot += sig[i] * sig[i];
      double run = 0; long long keep = 0;
      for (long long i = n - 1; i >= 0; --i) { run += sig[i] * sig[i]; if (run / tot > cutoff) { keep = i + 1; break; } }
      if (keep == 0) keep = 1;
      if (keep < maxd) maxd = keep;
    }
    kout[0] = (int)(maxd > mind ? maxd : mind);
    sweeps_out[0] = sweep;
  }
  // back to the workspace layout: Z(0:rows, j) = W(:, j), Z(rows + i, j) = V(i, j)
  for (int e = tid; e < rows * nact; e += blockDim.x) { const int r = e % rows, c = e / rows; Z[r + (long long)c * ldz] = W[e]; }
  for (int e = tid; e < nact * nact; e += blockDim.x) { const int r = e % nact, c = e / nact; Z[rows + r + (long long)c * ldz] = V[e]; }
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
  // one CTA per problem of a batched factorisation (grid = 1 otherwise): problem b owns entries [b * ncols, (b + 1) * ncols)
  sig2 += (long long)blockIdx.x * ncols; sig += (long long)blockIdx.x * ncols; perm += (long long)blockIdx.x * ncols; kout += blockIdx.x;
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
// scale: 0 -> 1, 1 -> sigma_j, 2 -> 1/sigma_j (0 if sigma_j == 0), 3 -> 1/sigma_j^2
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
    else if (scale_mode == 3) sc = sig[j] > 0 ? 1.0 / (sig[j] * sig[j]) : 0.0;
    if (!transpose_out) out[r + (long long)j * ldo] = make_double2(v.x * sc, v.y * sc);
    else out[j + (long long)r * ldo] = make_double2(v.x * sc, -v.y * sc);
  }
}

// ---- QR preconditioner kernels ----------------------------------------------------------------
// Shifted Cholesky of one 64x64 Gram matrix (sum of split-K partials) in shared memory, inverse of the
// triangular factor, and accumulation of the panel's R over the CholeskyQR passes:
//   G (+ shift I) = R^H R ;  Rinv = R^-1 ;  Rtot <- R * Rtot      (pass 0: Rtot = R, with the shift)
// Exactly-zero columns (padding, product states) are kept as null columns: pivot 1, row/column 0, and their
// row of Rtot is zeroed after the last pass.  A failed pivot in later passes leaves that column untouched.
constexpr int CHOL_THREADS = 1024;
__global__ void __launch_bounds__(CHOL_THREADS, 1) chol_inv64_kernel(const cplx* __restrict__ Gpart, int nsplit, long long split_stride,
                                                             cplx* __restrict__ Rinv_out, cplx* __restrict__ Rtot, int pass, int last_pass,
                                                             double shift_factor, int* __restrict__ done_flag) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx (*G)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw);                                    // G, then R in its upper triangle
  cplx (*Ri)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw + sizeof(cplx) * JP * (JP + 1));    // R^-1, then old Rtot
  __shared__ double red[CHOL_THREADS / 32];
  __shared__ int nullcol[JP];
  __shared__ double rrow[JP];       // 1 / R(j,j) per row (0 = null / failed pivot)
  __shared__ double rdinv[JP];      // reciprocal diagonal of R (divisions cost ~600 cycles of FP64 latency: do each once)
  const int tid = threadIdx.x;
  // batched factorisations (svd_batched_factor) launch one CTA per problem: CTA b works on the b-th 64 x 64 slot of every buffer
  Gpart += (long long)blockIdx.x * JP * JP; Rinv_out += (long long)blockIdx.x * JP * JP; Rtot += (long long)blockIdx.x * JP * JP;
  if (done_flag != nullptr) done_flag += blockIdx.x;
  // CholeskyQR3 with early termination: the second pass marks the panel as finished when its Gram matrix was already
  // within 1e-3 of the identity (one Cholesky pass on such a panel leaves O(eps) orthogonality error); the third pass'
  // Gram / Cholesky / apply launches then return at once.
  if (done_flag != nullptr) {
    if (pass == 0) { if (tid == 0) *done_flag = 0; }
    else if (pass == 2 && *done_flag != 0) return;
  }
  double fro = 0;
  for (int e = tid; e < JP * JP; e += CHOL_THREADS) {
    int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    for (int s = 0; s < nsplit; ++s) { cplx v = Gpart[s * split_stride + e]; xr += v.x; xi += v.y; }
    G[row][col] = make_double2(xr, xi);
    fro += xr * xr + xi * xi;
  }
  for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
  if ((tid & 31) == 0) red[tid >> 5] = fro;
  __syncthreads();
  fro = 0; for (int i = 0; i < CHOL_THREADS / 32; ++i) fro += red[i];
  const double shift = pass == 0 ? shift_factor * sqrt(fro) : 0.0;
  if (tid < JP) nullcol[tid] = (G[tid][tid].x <= 0.0) ? 1 : 0;
  __syncthreads();
  double dev = 0;                                        // max |G - I| over the non-null part (second pass only)
  for (int e = tid; e < JP * JP; e += CHOL_THREADS) {   // symmetrise + shift
    int row = e % JP, col = e / JP;
    if (row < col) {
      cplx a = G[row][col], b = G[col][row];
      cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row][col] = h; G[col][row] = make_double2(h.x, -h.y);
      dev = fmax(dev, fmax(fabs(h.x), fabs(h.y)));
    } else if (row == col) {
      if (!nullcol[row]) dev = fmax(dev, fabs(G[row][col].x - 1.0));
      G[row][col].x += shift; G[row][col].y = 0;
    }
  }
  if (done_flag != nullptr && pass == 1) {
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    __syncthreads();                                     // red[] was read above
    if ((tid & 31) == 0) red[tid >> 5] = dev;
    __syncthreads();
    dev = 0; for (int i = 0; i < CHOL_THREADS / 32; ++i) dev = fmax(dev, red[i]);
    if (dev < 1e-3) { last_pass = 1; if (tid == 0) *done_flag = 1; }
  }
  __syncthreads();
  // Right-looking Cholesky G = R^H R, one barrier per column: step j only READS row j (final since step j-1) and
  // applies the rank-1 update G(i,c) -= conj(G(j,i)) G(j,c) / G(j,j) to the trailing upper triangle, every thread
  // deriving 1/G(j,j) itself (no broadcast barrier); the rows are scaled by 1/sqrt(G(j,j)) in one pass at the end.
  // Thread t owns column c = t % 64 and rows (t / 64) + 16q.  1024 threads = 8 warps per scheduler: the ~50-cycle
  // dependent-issue latency of FP64 on this part is hidden by the other warps instead of being paid per operation.
  const int lane = tid & 31, wrp = tid >> 5;
  const int cc = 2 * wrp + (lane >> 4), part = lane & 15;     // back substitution: 2 columns per warp, 16 threads per column
  {
    const int c = tid & 63, i0 = tid >> 6;
    for (int j = 0; j < JP; ++j) {
      const double dd = G[j][j].x;
      const bool bad = nullcol[j] || !(dd > 0.0) || !isfinite(dd);
      const double ri = bad ? 0.0 : rsqrt(dd);
      const double sc = ri * ri;
      if (tid == 0) { rrow[j] = ri; rdinv[j] = bad ? 1.0 : ri; }
      if (!bad && c > j) {
        const cplx gjc = G[j][c];
#pragma unroll
        for (int q = 0; q < JP / (CHOL_THREADS / 64); ++q) {
          const int i = i0 + (CHOL_THREADS / 64) * q;
          if (i > j && i <= c) {
            const cplx gji = G[j][i];
            const double ar = gji.x * gjc.x + gji.y * gjc.y, ai = gji.x * gjc.y - gji.y * gjc.x;   // conj(G(j,i)) G(j,c)
            cplx v = G[i][c];
            v.x = fma(-sc, ar, v.x); v.y = fma(-sc, ai, v.y);
            G[i][c] = v;
          }
        }
      }
      __syncthreads();
    }
    for (int e = tid; e < JP * JP; e += CHOL_THREADS) {
      const int row = e / JP, col = e % JP;
      if (col < row) { G[row][col] = make_double2(0, 0); continue; }
      const double ri = rrow[row];
      if (ri == 0.0) G[row][col] = make_double2(col == row ? 1.0 : 0.0, 0.0);          // null / failed pivot: R(j,j) = 1, R(j,j+1:) = 0
      else { cplx v = G[row][col]; G[row][col] = (col == row) ? make_double2(v.x * ri, 0.0) : make_double2(v.x * ri, v.y * ri); }
    }
  }
  __syncthreads();
  // R^-1 by back substitution: column cc of the inverse, rows i = cc .. 0; the row's dot product is split over the 16
  // threads of the column group (<= 4 terms each).  Each warp owns 2 columns, so only warp-level synchronisation is needed.
  {
    for (int i = JP - 1 - part; i > cc; i -= 16) Ri[i][cc] = make_double2(0, 0);
    __syncwarp();
    const int cmax = 2 * wrp + 1;                       // largest column handled by this warp
    for (int i = cmax; i >= 0; --i) {
      double ar0 = 0, ai0 = 0;
      if (i <= cc) {
        for (int k = i + 1 + part; k <= cc; k += 16) { cplx a = G[i][k], b = Ri[k][cc]; ar0 += a.x * b.x - a.y * b.y; ai0 += a.x * b.y + a.y * b.x; }
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { ar0 += __shfl_xor_sync(0xffffffffu, ar0, o); ai0 += __shfl_xor_sync(0xffffffffu, ai0, o); }
      if (part == 0 && i <= cc) {
        double dinv = rdinv[i];
        Ri[i][cc] = make_double2(((i == cc ? 1.0 : 0.0) - ar0) * dinv, -ai0 * dinv);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < JP * JP; e += CHOL_THREADS) { int row = e % JP, col = e / JP; Rinv_out[e] = Ri[row][col]; }
  __syncthreads();
  for (int e = tid; e < JP * JP; e += CHOL_THREADS) { int row = e % JP, col = e / JP; Ri[row][col] = pass == 0 ? make_double2(row == col ? 1.0 : 0.0, 0.0) : Rtot[e]; }
  __syncthreads();
  for (int e = tid; e < JP * JP; e += CHOL_THREADS) {   // Rtot <- R * Rtot
    int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    double yr = 0, yi = 0;
    int k = row;
    for (; k + 1 < JP; k += 2) {
      cplx a = G[row][k], b = Ri[k][col], a2 = G[row][k + 1], b2 = Ri[k + 1][col];
      xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x;
      yr += a2.x * b2.x - a2.y * b2.y; yi += a2.x * b2.y + a2.y * b2.x;
    }
    if (k < JP) { cplx a = G[row][k], b = Ri[k][col]; xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x; }
    xr += yr; xi += yi;
    if (last_pass && nullcol[row]) { xr = 0; xi = 0; }
    Rtot[e] = make_double2(xr, xi);
  }
}

// ---- blocked variant of chol_inv64_kernel ------------------------------------------------------------------------------------------
// Same inputs, outputs and null-column / failed-pivot semantics, but the 64 pivots are processed in four 16-wide blocks: the diagonal
// block is factorised and inverted by ONE warp (warp-level synchronisation only), the block row R12 = R11^-H G12 and the trailing update
// G22 -= R12^H R12 are small GEMMs over the whole CTA -- 3 block barriers per 16 pivots instead of one per pivot -- and the inverse
// is assembled block column by block column.  ncu on the one-barrier-per-column kernel (profiles/r02_ncu_full_summary.json): 108 us
// per call, stalled on barriers / shared-memory latency at 25 % FP64 utilisation; it is half of the QR phase at n = 2048 and 70 % of it
// at n = 512 (64 calls per 512 x 512 factorisation).
constexpr int CB = 16;                       // pivot block
constexpr int CHOL2_THREADS = 512;
__global__ void __launch_bounds__(CHOL2_THREADS, 1) chol_inv64b_kernel(const cplx* __restrict__ Gpart, cplx* __restrict__ Rinv_out, cplx* __restrict__ Rtot,
                                                                       int pass, int last_pass, double shift_factor, int* __restrict__ done_flag) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx (*G)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw);                                    // G, then R in its upper triangle
  cplx (*Ri)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw + sizeof(cplx) * JP * (JP + 1));    // R^-1
  cplx (*Ro)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw + 2 * sizeof(cplx) * JP * (JP + 1)); // old Rtot
  __shared__ double red[CHOL2_THREADS / 32];
  __shared__ int nullcol[JP];
  __shared__ double rrow[JP];       // 1 / R(j,j) per row (0 = null / failed pivot)
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  Gpart += (long long)blockIdx.x * JP * JP; Rinv_out += (long long)blockIdx.x * JP * JP; Rtot += (long long)blockIdx.x * JP * JP;
  if (done_flag != nullptr) done_flag += blockIdx.x;
  if (done_flag != nullptr) {
    if (pass == 0) { if (tid == 0) *done_flag = 0; }
    else if (pass == 2 && *done_flag != 0) return;
  }
  double fro = 0;
  for (int e = tid; e < JP * JP; e += CHOL2_THREADS) {
    const int row = e % JP, col = e / JP;
    const cplx v = Gpart[e];
    G[row][col] = v;
    fro += v.x * v.x + v.y * v.y;
    Ri[row][col] = make_double2(0, 0);
  }
  for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
  if (lane == 0) red[wrp] = fro;
  __syncthreads();
  fro = 0; for (int i = 0; i < CHOL2_THREADS / 32; ++i) fro += red[i];
  const double shift = pass == 0 ? shift_factor * sqrt(fro) : 0.0;
  if (tid < JP) nullcol[tid] = (G[tid][tid].x <= 0.0) ? 1 : 0;
  __syncthreads();
  double dev = 0;                                        // max |G - I| over the non-null part (second pass only)
  for (int e = tid; e < JP * JP; e += CHOL2_THREADS) {   // symmetrise + shift
    const int row = e % JP, col = e / JP;
    if (row < col) {
      const cplx a = G[row][col], b = G[col][row];
      const cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row][col] = h; G[col][row] = make_double2(h.x, -h.y);
      dev = fmax(dev, fmax(fabs(h.x), fabs(h.y)));
    } else if (row == col) {
      if (!nullcol[row]) dev = fmax(dev, fabs(G[row][col].x - 1.0));
      G[row][col].x += shift; G[row][col].y = 0;
    }
  }
  if (done_flag != nullptr && pass == 1) {
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    __syncthreads();                                     // red[] was read above
    if (lane == 0) red[wrp] = dev;
    __syncthreads();
    dev = 0; for (int i = 0; i < CHOL2_THREADS / 32; ++i) dev = fmax(dev, red[i]);
    if (dev < 1e-3) { last_pass = 1; if (tid == 0) *done_flag = 1; }
  }
  __syncthreads();
  for (int kb = 0; kb < JP / CB; ++kb) {
    const int b0 = kb * CB, b1 = b0 + CB;
    if (wrp == 0) {
      // (1) right-looking Cholesky of the diagonal block by one warp: pivot j reads row j (final), updates the trailing upper triangle
      for (int j = b0; j < b1; ++j) {
        const double dd = G[j][j].x;
        const bool bad = nullcol[j] || !(dd > 0.0) || !isfinite(dd);
        const double ri = bad ? 0.0 : rsqrt(dd);
        const double sc = ri * ri;
        if (lane == 0) rrow[j] = ri;
        if (!bad) {
          const int nt = b1 - 1 - j;                     // trailing rows / columns inside the block
          for (int e = lane; e < nt * nt; e += 32) {
            const int i = j + 1 + e / nt, c = j + 1 + e % nt;
            if (i <= c) {
              const cplx gji = G[j][i], gjc = G[j][c];
              const double ar = gji.x * gjc.x + gji.y * gjc.y, ai = gji.x * gjc.y - gji.y * gjc.x;   // conj(G(j,i)) G(j,c)
              cplx v = G[i][c];
              v.x = fma(-sc, ar, v.x); v.y = fma(-sc, ai, v.y);
              G[i][c] = v;
            }
          }
        }
        __syncwarp();
      }
      // rows of the diagonal block scaled to R11 (null / failed pivot: unit row)
      for (int e = lane; e < CB * CB; e += 32) {
        const int row = b0 + e / CB, col = b0 + e % CB;
        if (col < row) continue;
        const double ri = rrow[row];
        if (ri == 0.0) G[row][col] = make_double2(col == row ? 1.0 : 0.0, 0.0);
        else { const cplx v = G[row][col]; G[row][col] = (col == row) ? make_double2(v.x * ri, 0.0) : make_double2(v.x * ri, v.y * ri); }
      }
      __syncwarp();
      // (2) inverse of the 16 x 16 upper-triangular R11 by back substitution: lane = column (lanes 16..31 idle)
      if (lane < CB) {
        const int cc = b0 + lane;
        for (int i = cc; i >= b0; --i) {
          double ar = 0, ai = 0;
          for (int k = i + 1; k <= cc; ++k) { const cplx a = G[i][k], b = Ri[k][cc]; ar += a.x * b.x - a.y * b.y; ai += a.x * b.y + a.y * b.x; }
          const double dinv = rrow[i] == 0.0 ? 1.0 : rrow[i];
          Ri[i][cc] = make_double2(((i == cc ? 1.0 : 0.0) - ar) * dinv, -ai * dinv);
        }
      }
    }
    __syncthreads();
    // (3) R12 = R11^-H G12 (rows b0..b1, columns >= b1); rows of null / failed pivots are zero
    const int nc = JP - b1;
    for (int e = tid; e < CB * nc; e += CHOL2_THREADS) {
      const int r = b0 + e % CB, c = b1 + e / CB;
      double xr = 0, xi = 0;
      if (rrow[r] != 0.0)
        for (int k = b0; k <= r; ++k) { const cplx a = Ri[k][r], g = G[k][c]; xr += a.x * g.x + a.y * g.y; xi += a.x * g.y - a.y * g.x; }   // conj(Rinv(k,r)) G(k,c)
      Ro[r][c] = make_double2(xr, xi);                  // scratch: another thread still needs the old G(r, c) for a row below r
    }
    __syncthreads();
    for (int e = tid; e < CB * nc; e += CHOL2_THREADS) { const int r = b0 + e % CB, c = b1 + e / CB; G[r][c] = Ro[r][c]; }
    __syncthreads();
    // (4) trailing update G22 -= R12^H R12 (upper triangle)
    for (int e = tid; e < nc * nc; e += CHOL2_THREADS) {
      const int i = b1 + e % nc, c = b1 + e / nc;
      if (i > c) continue;
      double xr = 0, xi = 0;
#pragma unroll 4
      for (int k = b0; k < b1; ++k) { const cplx a = G[k][i], b = G[k][c]; xr += a.x * b.x + a.y * b.y; xi += a.x * b.y - a.y * b.x; }
      cplx v = G[i][c]; v.x -= xr; v.y -= xi; G[i][c] = v;
    }
    __syncthreads();
  }
  for (int e = tid; e < JP * JP; e += CHOL2_THREADS) { const int row = e / JP, col = e % JP; if (col < row) G[row][col] = make_double2(0, 0); }
  __syncthreads();
  // (5) off-diagonal blocks of R^-1, block column by block column: Rinv_ij = -Rinv_ii sum_{i < l <= j} R_il Rinv_lj  (i = j-1 .. 0)
  for (int step = 1; step < JP / CB; ++step) {
    // all block columns j >= step handle their block row i = j - step at once (they only need rows > i of their own column)
    for (int e = tid; e < (JP / CB - step) * CB * CB; e += CHOL2_THREADS) {
      const int jb = step + e / (CB * CB), ib = jb - step, r = ib * CB + (e % (CB * CB)) % CB, c = jb * CB + (e % (CB * CB)) / CB;
      double xr = 0, xi = 0;
      for (int k = (ib + 1) * CB; k <= c; ++k) { const cplx a = G[r][k], b = Ri[k][c]; xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x; }
      Ro[c][r] = make_double2(xr, xi);                  // T = R_i,(i+1..j) Rinv_(i+1..j),j  (scratch, stored transposed)
    }
    __syncthreads();
    for (int e = tid; e < (JP / CB - step) * CB * CB; e += CHOL2_THREADS) {
      const int jb = step + e / (CB * CB), ib = jb - step, r = ib * CB + (e % (CB * CB)) % CB, c = jb * CB + (e % (CB * CB)) / CB;
      double xr = 0, xi = 0;
      for (int k = r; k < (ib + 1) * CB; ++k) { const cplx a = Ri[r][k], t = Ro[c][k]; xr += a.x * t.x - a.y * t.y; xi += a.x * t.y + a.y * t.x; }
      Ri[r][c] = make_double2(-xr, -xi);
    }
    __syncthreads();
  }
  for (int e = tid; e < JP * JP; e += CHOL2_THREADS) { const int row = e % JP, col = e / JP; Rinv_out[e] = Ri[row][col]; }
  // (6) Rtot <- R * Rtot_old  (pass 0: Rtot = R)
  if (pass != 0) {
    for (int e = tid; e < JP * JP; e += CHOL2_THREADS) Ro[e % JP][e / JP] = Rtot[e];     // the accumulated R of the earlier passes
    __syncthreads();
  }
  for (int e = tid; e < JP * JP; e += CHOL2_THREADS) {
    const int row = e % JP, col = e / JP;
    double xr = 0, xi = 0;
    if (pass == 0) { if (col >= row) { xr = G[row][col].x; xi = G[row][col].y; } }
    else {
      for (int k = row; k < JP; ++k) { const cplx a = G[row][k], b = Ro[k][col]; xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x; }
    }
    if (last_pass && nullcol[row]) { xr = 0; xi = 0; }
    Rtot[e] = make_double2(xr, xi);
  }
}

// ---- register-blocked variant (default) ---------------------------------------------------------------------------------------------
// Same inputs, outputs and null-column / failed-pivot semantics.  The blocked kernel above still spends ~72 us per call (ncu launch
// list of a 2048^2 factorisation: 256 calls = 18 of the 52 ms of the QR phase, and half of the QR phase at n = 512): its 16 x 16 diagonal
// blocks are factorised and inverted by one warp through shared memory, a chain of ~20 k cycles per block.  Here 256 threads own one
// 4 x 4 block of G each IN REGISTERS (thread (ti, tj): rows 4 ti.., columns 4 tj..).  Per 4 pivots: the diagonal thread factorises its
// block and inverts the 4 x 4 triangle with fully unrolled register arithmetic (the only sequential part: four rsqrt chains), the
// threads of the block row form R12 = R11^-H G12 in registers and publish it, every trailing thread applies the rank-4 update to its
// own registers -- two block barriers per 4 pivots and no shared-memory round trip on the pivot chain.  R^-1 is then assembled by
// recursive doubling ([A B; 0 C]^-1 = [A^-1, -A^-1 B C^-1; 0, C^-1], block sizes 4, 8, 16, 32) with the whole CTA.
constexpr int CHOL3_THREADS = 256;
__global__ void __launch_bounds__(CHOL3_THREADS, 1) chol_inv64c_kernel(const cplx* __restrict__ Gpart, cplx* __restrict__ Rinv_out, cplx* __restrict__ Rtot,
                                                                       int pass, int last_pass, double shift_factor, int* __restrict__ done_flag) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  cplx (*G)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw);                                    // G, then R in its upper triangle
  cplx (*Ri)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw + sizeof(cplx) * JP * (JP + 1));    // R^-1
  cplx (*Ro)[JP + 1] = reinterpret_cast<cplx (*)[JP + 1]>(sm_raw + 2 * sizeof(cplx) * JP * (JP + 1)); // scratch / old Rtot
  __shared__ double red[CHOL3_THREADS / 32];
  __shared__ int nullcol[JP];
  __shared__ int rowbad[JP];                          // null column or failed pivot: R(j, :) = e_j
  __shared__ cplx Pn[4][JP];                          // R12 of the current pivot block, column 4 tj + b stored at tj + 16 b
  __shared__ cplx X11[4][4];                          // inverse of the current diagonal block
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  Gpart += (long long)blockIdx.x * JP * JP; Rinv_out += (long long)blockIdx.x * JP * JP; Rtot += (long long)blockIdx.x * JP * JP;
  if (done_flag != nullptr) done_flag += blockIdx.x;
  if (done_flag != nullptr) {
    if (pass == 0) { if (tid == 0) *done_flag = 0; }
    else if (pass == 2 && *done_flag != 0) return;
  }
  double fro = 0;
  for (int e = tid; e < JP * JP; e += CHOL3_THREADS) {
    const int row = e % JP, col = e / JP;
    const cplx v = Gpart[e];
    G[row][col] = v;
    fro += v.x * v.x + v.y * v.y;
    Ri[row][col] = make_double2(0, 0);
  }
  for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
  if (lane == 0) red[wrp] = fro;
  __syncthreads();
  fro = 0; for (int i = 0; i < CHOL3_THREADS / 32; ++i) fro += red[i];
  const double shift = pass == 0 ? shift_factor * sqrt(fro) : 0.0;
  if (tid < JP) nullcol[tid] = (G[tid][tid].x <= 0.0) ? 1 : 0;
  __syncthreads();
  double dev = 0;                                        // max |G - I| over the non-null part (second pass only)
  for (int e = tid; e < JP * JP; e += CHOL3_THREADS) {   // symmetrise + shift
    const int row = e % JP, col = e / JP;
    if (row < col) {
      const cplx a = G[row][col], b = G[col][row];
      const cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      G[row][col] = h; G[col][row] = make_double2(h.x, -h.y);
      dev = fmax(dev, fmax(fabs(h.x), fabs(h.y)));
    } else if (row == col) {
      if (!nullcol[row]) dev = fmax(dev, fabs(G[row][col].x - 1.0));
      G[row][col].x += shift; G[row][col].y = 0;
    }
  }
  if (done_flag != nullptr && pass == 1) {
    for (int o = 16; o > 0; o >>= 1) dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    __syncthreads();                                     // red[] was read above
    if (lane == 0) red[wrp] = dev;
    __syncthreads();
    dev = 0; for (int i = 0; i < CHOL3_THREADS / 32; ++i) dev = fmax(dev, red[i]);
    if (dev < 1e-3) { last_pass = 1; if (tid == 0) *done_flag = 1; }
  }
  __syncthreads();

  // ---- factorisation: thread (ti, tj) keeps G(4 ti + a, 4 tj + b) in registers (only tj >= ti matters) ----
  const int ti = tid >> 4, tj = tid & 15;
  cplx B[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) B[a][b] = G[4 * ti + a][4 * tj + b];
  __syncthreads();                                       // G is rewritten below (R) while other threads may still be loading
  for (int kb = 0; kb < JP / 4; ++kb) {
    if (ti == kb && tj == kb) {
      // (1) diagonal block: R11 (rows of null / failed pivots become unit rows) and X = R11^-1, all in registers
      double dinv[4]; bool bad[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double dd = B[a][a].x;
        bad[a] = nullcol[4 * kb + a] || !(dd > 0.0) || !isfinite(dd);
        const double ri = bad[a] ? 0.0 : rsqrt(dd);
        dinv[a] = bad[a] ? 1.0 : ri;
        B[a][a] = make_double2(bad[a] ? 1.0 : dd * ri, 0.0);
#pragma unroll
        for (int b = a + 1; b < 4; ++b) B[a][b] = bad[a] ? make_double2(0, 0) : make_double2(B[a][b].x * ri, B[a][b].y * ri);
#pragma unroll
        for (int a2 = a + 1; a2 < 4; ++a2)
#pragma unroll
          for (int b2 = a2; b2 < 4; ++b2) {              // G(a2, b2) -= conj(R(a, a2)) R(a, b2)
            const cplx u = B[a][a2], v = B[a][b2];
            B[a2][b2].x -= u.x * v.x + u.y * v.y; B[a2][b2].y -= u.x * v.y - u.y * v.x;
          }
      }
      cplx X[4][4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int a = 3; a >= 0; --a) {
          if (a > b) { X[a][b] = make_double2(0, 0); continue; }
          if (a == b) { X[a][b] = make_double2(dinv[a], 0.0); continue; }
          double xr = 0, xi = 0;
#pragma unroll
          for (int k = a + 1; k <= b; ++k) { xr += B[a][k].x * X[k][b].x - B[a][k].y * X[k][b].y; xi += B[a][k].x * X[k][b].y + B[a][k].y * X[k][b].x; }
          X[a][b] = make_double2(-xr * dinv[a], -xi * dinv[a]);
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        rowbad[4 * kb + a] = bad[a] ? 1 : 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          G[4 * kb + a][4 * kb + b] = (b >= a) ? B[a][b] : make_double2(0, 0);
          Ri[4 * kb + a][4 * kb + b] = X[a][b];
          X11[a][b] = X[a][b];
        }
      }
    }
    __syncthreads();
    if (ti == kb && tj > kb) {
      // (2) R12 = R11^-H G12: row r = sum_{k <= r} conj(X(k, r)) G12(k, :); rows of null / failed pivots are zero
      cplx Nw[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const bool zr = rowbad[4 * kb + r] != 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double xr = 0, xi = 0;
#pragma unroll
          for (int k = 0; k <= r; ++k) { const cplx x = X11[k][r]; xr += x.x * B[k][c].x + x.y * B[k][c].y; xi += x.x * B[k][c].y - x.y * B[k][c].x; }
          Nw[r][c] = zr ? make_double2(0, 0) : make_double2(xr, xi);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { B[r][c] = Nw[r][c]; G[4 * kb + r][4 * tj + c] = Nw[r][c]; Pn[r][tj + 16 * c] = Nw[r][c]; }
    }
    __syncthreads();
    if (ti > kb && tj >= ti) {
      // (3) trailing update G22 -= R12^H R12 on the thread's own block
      cplx L[4][4], Rr[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a) { L[k][a] = Pn[k][ti + 16 * a]; Rr[k][a] = Pn[k][tj + 16 * a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          double xr = 0, xi = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { xr += L[k][a].x * Rr[k][b].x + L[k][a].y * Rr[k][b].y; xi += L[k][a].x * Rr[k][b].y - L[k][a].y * Rr[k][b].x; }
          B[a][b].x -= xr; B[a][b].y -= xi;
        }
    }
    // no barrier here: step kb + 1 starts in the registers of thread (kb + 1, kb + 1); Pn / X11 are rewritten only after the next barrier
    // ... except X11 and rowbad, written in (1) of the next step while (2)/(3) of this step never read them after the barrier above
  }
  __syncthreads();
  for (int e = tid; e < JP * JP; e += CHOL3_THREADS) { const int row = e / JP, col = e % JP; if (col < row) G[row][col] = make_double2(0, 0); }
  __syncthreads();
  // ---- R^-1 by recursive doubling: the diagonal 4 x 4 blocks are in Ri; merge blocks of size h into 2h ----
  // A dependent FP64 operation costs ~50 cycles on this part and the CTA has two warps per scheduler, so every dot product below runs
  // four output columns x two k-phases = 16 independent accumulation chains per thread (a task = one row, four consecutive columns).
  for (int h = 4; h < JP; h *= 2) {
    const int nprob = JP / (2 * h), q4 = h / 4;
    // T = B C^-1 (B = R(rows, cols), C^-1 = Ri(cols, cols) upper triangular: its lower triangle holds zeros)
    for (int e = tid; e < nprob * h * q4; e += CHOL3_THREADS) {
      const int pb = e / (h * q4), r = (e % (h * q4)) / q4, c = 4 * (e % q4);
      const int r0 = pb * 2 * h, c0 = r0 + h;
      double xr[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, xi[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int k = 0; k <= c + 3; k += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if (k + ph > c + 3) continue;
          const cplx a = G[r0 + r][c0 + k + ph];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const cplx b = Ri[c0 + k + ph][c0 + c + j]; xr[ph][j] += a.x * b.x - a.y * b.y; xi[ph][j] += a.x * b.y + a.y * b.x; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Ro[r0 + r][c0 + c + j] = make_double2(xr[0][j] + xr[1][j], xi[0][j] + xi[1][j]);
    }
    __syncthreads();
    // Ri(rows, cols) = -A^-1 T (A^-1 = Ri(rows, rows) upper triangular)
    for (int e = tid; e < nprob * h * q4; e += CHOL3_THREADS) {
      const int pb = e / (h * q4), r = (e % (h * q4)) / q4, c = 4 * (e % q4);
      const int r0 = pb * 2 * h, c0 = r0 + h;
      double xr[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, xi[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int k = r; k < h; k += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if (k + ph >= h) continue;
          const cplx a = Ri[r0 + r][r0 + k + ph];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const cplx t = Ro[r0 + k + ph][c0 + c + j]; xr[ph][j] += a.x * t.x - a.y * t.y; xi[ph][j] += a.x * t.y + a.y * t.x; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Ri[r0 + r][c0 + c + j] = make_double2(-(xr[0][j] + xr[1][j]), -(xi[0][j] + xi[1][j]));
    }
    __syncthreads();
  }
  for (int e = tid; e < JP * JP; e += CHOL3_THREADS) { const int row = e % JP, col = e / JP; Rinv_out[e] = Ri[row][col]; }
  // ---- Rtot <- R * Rtot_old  (pass 0: Rtot = R); both factors are upper triangular: k runs over [row, col] only ----
  if (pass != 0) {
    for (int e = tid; e < JP * JP; e += CHOL3_THREADS) Ro[e % JP][e / JP] = Rtot[e];
    __syncthreads();
  }
  for (int e = tid; e < JP * (JP / 4); e += CHOL3_THREADS) {
    const int row = e / (JP / 4), col = 4 * (e % (JP / 4));
    double xr[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, xi[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (pass == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (col + j >= row) { xr[0][j] = G[row][col + j].x; xi[0][j] = G[row][col + j].y; }
    } else if (col + 3 >= row) {
      for (int k = row; k <= col + 3; k += 2) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if (k + ph > col + 3) continue;
          const cplx a = G[row][k + ph];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const cplx b = Ro[k + ph][col + j]; xr[ph][j] += a.x * b.x - a.y * b.y; xi[ph][j] += a.x * b.y + a.y * b.x; }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double vr = xr[0][j] + xr[1][j], vi = xi[0][j] + xi[1][j];
      if (last_pass && nullcol[row]) { vr = 0; vi = 0; }
      Rtot[row + (long long)JP * (col + j)] = make_double2(vr, vi);
    }
  }
}

// dst(c, r) = conj(src(r, c)): dst is cols x rows (ldd), src rows x cols (lds); optional zero fill beyond (rows_valid, cols_valid)
__global__ void __launch_bounds__(256) conj_transpose_kernel(const cplx* __restrict__ src, long long lds, int rows, int cols,
                                                              cplx* __restrict__ dst, long long ldd) {
  long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % cols), r = (int)(e / cols);
    cplx v = src[r + (long long)c * lds];
    dst[c + (long long)r * ldd] = make_double2(v.x, -v.y);
  }
}
__global__ void __launch_bounds__(256) set_identity_kernel(cplx* __restrict__ dst, long long ld, int n) {
  long long total = (long long)n * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int r = (int)(e % n), c = (int)(e / n);
    dst[r + (long long)c * ld] = make_double2(r == c ? 1.0 : 0.0, 0.0);
  }
}

// ---- host driver ------------------------------------------------------------------------------

// Panel Cholesky + inverse over `nb` Gram blocks (TN_SVD_CHOL=1 selects the one-barrier-per-column kernel, 2 the 16-wide blocked one).
static void launch_chol(int nb, const cplx* Gp, cplx* Rinv, cplx* Rtot, int pass, int last, double shift_factor, int* flag, cudaStream_t s) {
  static int ver = -1;
  if (ver < 0) { const char* e = getenv("TN_SVD_CHOL"); ver = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 3); }
  if (ver == 3) {
    static DeviceOnce cfg3;
    const int smem = 3 * JP * (JP + 1) * (int)sizeof(cplx);
    cfg3.run([&] { TN_CUDA(cudaFuncSetAttribute(chol_inv64c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); });
    chol_inv64c_kernel<<<nb, CHOL3_THREADS, smem, s>>>(Gp, Rinv, Rtot, pass, last, shift_factor, flag);
  } else if (ver == 1) {
    static DeviceOnce cfg1;
    const int smem = 2 * JP * (JP + 1) * (int)sizeof(cplx);
    cfg1.run([&] { TN_CUDA(cudaFuncSetAttribute(chol_inv64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); });
    chol_inv64_kernel<<<nb, CHOL_THREADS, smem, s>>>(Gp, 1, (long long)JP * JP, Rinv, Rtot, pass, last, shift_factor, flag);
  } else {
    static DeviceOnce cfg2;
    const int smem = 3 * JP * (JP + 1) * (int)sizeof(cplx);
    cfg2.run([&] { TN_CUDA(cudaFuncSetAttribute(chol_inv64b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); });
    chol_inv64b_kernel<<<nb, CHOL2_THREADS, smem, s>>>(Gp, Rinv, Rtot, pass, last, shift_factor, flag);
  }
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// Launches the pair EVD over `npairs` Gram blocks (TN_SVD_EVD=1 selects the first-generation kernel).
static void launch_evd(int npairs, const cplx* Gp, cplx* Jp, double tol, unsigned long long* offmax, int inner, int nact, int* skip, cudaStream_t s) {
  static int ver = -1;
  if (ver < 0) { const char* e = getenv("TN_SVD_EVD"); ver = (e && e[0] == '1') ? 1 : 2; }
  if (ver == 1) {
    static DeviceOnce cfg1;
    const int smem = 2 * JP * LDS_ * (int)sizeof(cplx);
    cfg1.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_evd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); });
    jacobi_evd64_kernel<<<npairs, EVD_THREADS, smem, s>>>(Gp, 1, 0, Jp, tol, offmax, inner, nact, skip);
  } else {
    static DeviceOnce cfg2;
    cfg2.run([&] { TN_CUDA(cudaFuncSetAttribute(jacobi_evd64v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EVD2_SMEM)); });
    jacobi_evd64v2_kernel<<<npairs, EVD2_THREADS, EVD2_SMEM, s>>>(Gp, 1, 0, Jp, tol, offmax, inner, nact, skip);
  }
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

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

static int g_precond_mode = -1;   // -1: read TN_SVD_PRECOND once (default on)
void svd_set_precond(int mode) { g_precond_mode = mode ? 1 : 0; }
static bool precond_enabled() {
  if (g_precond_mode < 0) { const char* e = getenv("TN_SVD_PRECOND"); g_precond_mode = (e && e[0] == '0') ? 0 : 1; }
  return g_precond_mode == 1;
}

// Upper bound on the split-K factors of the SVD's Gram / projection GEMMs.  Splitting trades CTA efficiency (short k
// loops, atomic epilogues) for latency; a single stream wants the latency, many concurrent streams (QJMC ensembles) are
// throughput-bound and can lower it (TN_SVD_MAXSPLIT).
static int max_split() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_MAXSPLIT"); v = e ? std::max(1, atoi(e)) : 32; }
  return v;
}

static double early_stop() {
  static double v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_EARLY"); v = e ? atof(e) : 1e-9; }
  return v;
}
static bool small_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_SMALL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
static int small_max_cols() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_SMALL_MAXCOLS"); v = e ? std::max(2, std::min(JP, atoi(e))) : 32; }
  return v;
}
static bool wonly_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_WONLY"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// TN_SVD_JGEMM=0 routes the Gram blocks and the pair rotations through the general strided GEMM (tn_zgemm.cu) instead of the
// dedicated kernels of tn_jacobi.cu (A/B runs).
static bool jgemm_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_JGEMM"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static bool cross_enabled() {       // TN_SVD_CROSS=0: projection coefficients of the QR through the general GEMM (A/B runs)
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_CROSS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static void launch_1d(long long total, int& blocks) { blocks = (int)std::max<long long>(1, std::min<long long>(148 * 8, (total + 255) / 256)); }

// Optional phase timing (TN_SVD_PROFILE=1): CUDA events at the phase boundaries, summed per factorisation and printed
// to stderr as one JSON line.  Diagnostics only; never enabled in tests or benches.
struct SvdProf {
  bool on = false;
  std::vector<cudaEvent_t> ev; size_t used = 0;
  std::vector<int> tag;                 // phase tag of the interval ENDING at event i
  double ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void mark(int t, cudaStream_t s) {
    if (!on) return;
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); tag.push_back(0); }
    tag[used] = t; cudaEventRecord(ev[used++], s);
  }
  void flush(cudaStream_t s) {
    if (!on || used == 0) return;
    cudaStreamSynchronize(s);
    for (size_t i = 1; i < used; ++i) { float m = 0; cudaEventElapsedTime(&m, ev[i - 1], ev[i]); ms[tag[i]] += m; }
    // keep the last event as the start of the next interval
    std::swap(ev[0], ev[used - 1]); used = 1;
  }
};
static SvdProf& prof() {
  static thread_local SvdProf p;
  static int on = -1;
  if (on < 0) { const char* e = getenv("TN_SVD_PROFILE"); on = (e && e[0] == '1') ? 1 : 0; }
  p.on = on == 1;
  return p;
}
enum { PH_START = 0, PH_GRAM = 1, PH_EVD = 2, PH_ROT = 3, PH_QR = 4, PH_FIN = 5, PH_MISC = 6 };

// ---- Split schedule ---------------------------------------------------------------------------------------------------------
// In the circle method every pair of step t+1 takes its two blocks from the two NEIGHBOURING pairs of step t, so groups of pairs put
// on separate streams advance in lockstep and the pair EVD (a latency chain on ~np SMs, 108 us per step whatever np is) never overlaps
// the Gram / rotation kernels of another group (measured, < 3 %).  A sweep only has to visit every pair of column blocks once; the
// recursive order
//     RR(S)     = RR(S1) || RR(S2),  then BIP(S1, S2)                         (round robin inside a set of blocks)
//     BIP(A, B) = BIP(A1, B1) || BIP(A2, B2),  then BIP(A1, B2) || BIP(A2, B1)  (every pair between two sets)
// has the same number of steps when the block count divides evenly, but its concurrent tasks work on DISJOINT blocks for a whole
// phase (3 phases per sweep for 2 groups, 7 for 4), so their streams drift apart freely: while one group sits in its EVD the others'
// GEMMs have the machine.  tools/jacobi_sched_emul.py checks the schedule (every pair once, disjoint tasks) and its sweep counts.
struct SplitSched {
  struct Task { std::vector<int> step_off, step_np; int slot0 = 0; };   // offsets (in pairs) into the table; first G / J / skip slot
  struct Phase { std::vector<Task> tasks; };
  int nb = 0, groups = 1, depth = 0;
  std::vector<Phase> phases;
  std::vector<int> host;     // (p, q) of every pair, task after task, step after step
  int* dev = nullptr;
};
namespace {
typedef std::vector<std::vector<std::pair<int, int>>> Steps;          // one leaf task: steps of disjoint pairs
typedef std::vector<std::vector<Steps>> Phases;                        // phases of concurrent leaf tasks
Steps rr_steps(std::vector<int> S) {
  Steps out;
  if (S.size() < 2) return out;
  if (S.size() % 2) S.push_back(-1);                                   // bye
  const int n = (int)S.size();
  for (int st = 0; st < n - 1; ++st) {
    std::vector<std::pair<int, int>> prs;
    for (int k = 0; k < n / 2; ++k) {
      int a, b;
      if (n == 2) { a = 0; b = 1; }
      else if (k == 0) { a = n - 1; b = st; }
      else { a = (st + k) % (n - 1); b = (st - k + (n - 1)) % (n - 1); }
      if (S[a] >= 0 && S[b] >= 0) prs.push_back({std::min(S[a], S[b]), std::max(S[a], S[b])});
    }
    out.push_back(prs);
  }
  return out;
}
Steps bip_steps(std::vector<int> A, std::vector<int> B) {
  Steps out;
  if (A.size() > B.size()) std::swap(A, B);
  if (A.empty()) return out;
  const int nB = (int)B.size();
  for (int sft = 0; sft < nB; ++sft) {
    std::vector<std::pair<int, int>> prs;
    for (int i = 0; i < (int)A.size(); ++i) { const int q = B[(i + sft) % nB]; prs.push_back({std::min(A[i], q), std::max(A[i], q)}); }
    out.push_back(prs);
  }
  return out;
}
Phases par(const Phases& x, const Phases& y) {
  Phases out;
  for (size_t i = 0; i < std::max(x.size(), y.size()); ++i) {
    std::vector<Steps> ph;
    if (i < x.size()) ph.insert(ph.end(), x[i].begin(), x[i].end());
    if (i < y.size()) ph.insert(ph.end(), y[i].begin(), y[i].end());
    out.push_back(ph);
  }
  return out;
}
Phases cat(Phases x, const Phases& y) { x.insert(x.end(), y.begin(), y.end()); return x; }
Phases expand_bip(const std::vector<int>& A, const std::vector<int>& B, int c);
Phases expand_rr(const std::vector<int>& S, int c) {
  if (c <= 1 || S.size() < 4) { Steps st = rr_steps(S); return st.empty() ? Phases{} : Phases{{st}}; }
  const size_t h = (S.size() + 1) / 2;
  const std::vector<int> S1(S.begin(), S.begin() + h), S2(S.begin() + h, S.end());
  return cat(par(expand_rr(S1, c / 2), expand_rr(S2, c / 2)), expand_bip(S1, S2, c));
}
Phases expand_bip(const std::vector<int>& A, const std::vector<int>& B, int c) {
  if (c <= 1 || std::min(A.size(), B.size()) < 2) { Steps st = bip_steps(A, B); return st.empty() ? Phases{} : Phases{{st}}; }
  const size_t ha = (A.size() + 1) / 2, hb = (B.size() + 1) / 2;
  const std::vector<int> A1(A.begin(), A.begin() + ha), A2(A.begin() + ha, A.end()), B1(B.begin(), B.begin() + hb), B2(B.begin() + hb, B.end());
  return cat(par(expand_bip(A1, B1, c / 2), expand_bip(A2, B2, c / 2)), par(expand_bip(A1, B2, c / 2), expand_bip(A2, B1, c / 2)));
}
}  // namespace

// The schedule for nb blocks in `groups` concurrent groups (cached in the workspace); nullptr when it would need more steps than the
// circle method (block counts that do not halve evenly) or offers no concurrency.
static SplitSched* split_schedule(SvdWork& w, int nb, int groups, cudaStream_t s) {
  if (w.split && w.split->nb == nb && w.split->groups == groups) return w.split->phases.empty() ? nullptr : w.split;
  if (w.split) { if (w.split->dev) cudaFree(w.split->dev); delete w.split; w.split = nullptr; }
  SplitSched* sc = new SplitSched();
  sc->nb = nb; sc->groups = groups;
  w.split = sc;
  std::vector<int> all(nb);
  for (int i = 0; i < nb; ++i) all[i] = i;
  Phases ph;
  bool ok = false;
  for (int g = groups; g >= 2 && !ok; g /= 2) {       // e.g. 36 blocks: quarters of 9 need byes (more steps), halves of 18 do not
    ph = expand_rr(all, g);
    int depth = 0; size_t maxtasks = 0;
    for (auto& p : ph) { size_t d = 0; for (auto& t : p) d = std::max(d, t.size()); depth += (int)d; maxtasks = std::max(maxtasks, p.size()); }
    ok = depth == nb - 1 && maxtasks >= 2 && maxtasks <= 4;
  }
  if (!ok) return nullptr;       // phases stays empty: "no split schedule for this nb"
  for (auto& p : ph) {
    SplitSched::Phase P;
    int slot = 0;
    for (auto& t : p) {
      SplitSched::Task T;
      T.slot0 = slot;
      int mx = 0;
      for (auto& st : t) {
        T.step_off.push_back((int)(sc->host.size() / 2)); T.step_np.push_back((int)st.size());
        for (auto& pq : st) { sc->host.push_back(pq.first); sc->host.push_back(pq.second); }
        mx = std::max(mx, (int)st.size());
      }
      slot += mx;
      P.tasks.push_back(T);
    }
    if (slot > nb / 2) { sc->phases.clear(); sc->host.clear(); return nullptr; }   // would not fit the G / J / skip buffers of a circle-method step
    sc->phases.push_back(P);
  }
  sc->depth = nb - 1;
  TN_CUDA(cudaMalloc((void**)&sc->dev, sc->host.size() * sizeof(int)));
  TN_CUDA(cudaMemcpyAsync(sc->dev, sc->host.data(), sc->host.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TN_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < 3; ++i) if (!w.aux[i]) {
    TN_CUDA(cudaStreamCreateWithFlags(&w.aux[i], cudaStreamNonBlocking));
    TN_CUDA(cudaEventCreateWithFlags(&w.ev_join[i], cudaEventDisableTiming));
  }
  if (!w.ev_fork) TN_CUDA(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming));
  return sc;
}
// Host-side view of the split schedule for tests: writes (phase, task, step, p, q) per pair into out5 (capacity cap5 pairs) and returns the
// number of pairs, 0 when this block count has no split schedule (the sweeps then use the circle method), -1 when out5 is too small.
long long svd_split_schedule_dump(int nb, int groups, int* out5, long long cap5) {
  std::vector<int> all(nb);
  for (int i = 0; i < nb; ++i) all[i] = i;
  Phases ph = expand_rr(all, groups);
  int depth = 0; size_t maxtasks = 0;
  for (auto& p : ph) { size_t d = 0; for (auto& t : p) d = std::max(d, t.size()); depth += (int)d; maxtasks = std::max(maxtasks, p.size()); }
  if (depth != nb - 1 || maxtasks < 2 || maxtasks > 4) return 0;
  long long n = 0;
  for (size_t a = 0; a < ph.size(); ++a) {
    int slot = 0;
    for (size_t b = 0; b < ph[a].size(); ++b) {
      int mx = 0;
      for (size_t c = 0; c < ph[a][b].size(); ++c) {
        mx = std::max(mx, (int)ph[a][b][c].size());
        for (auto& pq : ph[a][b][c]) {
          if (n >= cap5) return -1;
          int* o = out5 + 5 * n++;
          o[0] = (int)a; o[1] = (int)b; o[2] = (int)c; o[3] = pq.first; o[4] = pq.second;
        }
      }
      slot += mx;
    }
    if (slot > nb / 2) return 0;
  }
  return n;
}

// number of concurrent pair groups (TN_SVD_SPLIT; 1 = circle method on one stream) and the smallest block count that is split
static int split_groups() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_SPLIT"); v = e ? std::max(1, std::min(4, atoi(e))) : 4; }
  return v;
}
static int split_min_blocks() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_SVD_SPLIT_MINNB"); v = e ? std::max(4, atoi(e)) : 32; }
  return v;
}

// Jacobi sweeps on Z = [W ; V] (W: jrows x ncols_pad, V: ncols_pad x ncols_pad), leading dimension ldz.
static void jacobi_sweeps(SvdWork& w, int jrows, cudaStream_t s) {
  const int nb = w.ncols_pad / JB, np = nb / 2, steps = nb - 1;
  // Gram GEMM grid = np pairs x ksplit: fill the 2 x 148 CTA slots of the 64 x 64 kernel once (no partial second wave)
  int ksplit = std::max(1, std::min(std::min(std::min(16, max_split()), jrows / 64), (2 * 148) / np));
  int kchunk = ((jrows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (jrows + kchunk - 1) / kchunk;
  ensure(w.Gpart, w.G_cap, (size_t)std::max(ksplit, 32) * np * JP * JP, s);
  ensure(w.J, w.J_cap, (size_t)np * JP * JP, s);
  ensure(w.skip, w.skip_cap, (size_t)np, s);                // per-pair "already converged" flags
  const int* tab = pair_table(w, nb, s);
  // Convergence threshold on |g_pq| / sqrt(g_pp g_qq).  The DMMA Gram blocks carry a rounding error of about sqrt(K) eps
  // (K = jrows accumulations, split-K partials summed in arbitrary order), so a threshold of exactly sqrt(K) eps makes
  // the last sweeps chase noise (9-11 sweeps run to run at n = 2048); 3 sqrt(K) eps sits just above that floor and is
  // still 100x below what the spectra / orthogonality bounds need.
  static double tol_factor = -1;
  if (tol_factor < 0) { const char* e = getenv("TN_SVD_TOL_FACTOR"); tol_factor = e ? atof(e) : 3.0; }
  const double tol = tol_factor * std::sqrt((double)jrows) * 2.220446049250313e-16;
  const long long colblk = (long long)JB * w.ldz;
  const int max_sweeps = 60;
  // One inner sweep per visit gives the same number of outer sweeps as a full inner diagonalisation
  // (tools/jacobi_emul.py) at 1/8 of the cost and with less rounding accumulated in V; a single pair
  // (n <= 64) has no outer parallelism to trade, so it is diagonalised fully.
  const int inner_sweeps = (np == 1) ? 12 : 1;
  const int nact = (np == 1) ? std::max(2, std::min(JP, (w.ncols + 1) / 2 * 2)) : JP;
  w.sweeps = 0;
  SplitSched* sched = (jgemm_enabled() && !prof().on && split_groups() > 1 && nb >= split_min_blocks() && np > 1)
                          ? split_schedule(w, nb, split_groups(), s) : nullptr;
  const int rot_rows = jrows + (w.wonly ? 0 : w.ncols_pad);
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    TN_CUDA(cudaMemsetAsync(w.offmax, 0, 8, s));
    if (sched) {
      for (const auto& ph : sched->phases) {
        const int nt = (int)ph.tasks.size();
        size_t depth = 0;
        for (const auto& t : ph.tasks) depth = std::max(depth, t.step_off.size());
        if (nt > 1) {
          TN_CUDA(cudaEventRecord(w.ev_fork, s));
          for (int t = 1; t < nt; ++t) TN_CUDA(cudaStreamWaitEvent(w.aux[t - 1], w.ev_fork, 0));
        }
        for (size_t i = 0; i < depth; ++i)
          for (int t = 0; t < nt; ++t) {
            const auto& T = ph.tasks[t];
            if (i >= T.step_off.size() || T.step_np[i] == 0) continue;
            cudaStream_t gs = t == 0 ? s : w.aux[t - 1];
            const int* tb = sched->dev + 2 * (size_t)T.step_off[i];
            const int npg = T.step_np[i];
            cplx* Gg = w.Gpart + (size_t)T.slot0 * JP * JP;
            cplx* Jg = w.J + (size_t)T.slot0 * JP * JP;
            jacobi_gram64(w.Z, w.ldz, jrows, tb, npg, Gg, std::min(max_split(), 8), gs);
            launch_evd(npg, Gg, Jg, tol, w.offmax, inner_sweeps, nact, w.skip + T.slot0, gs);
            jacobi_rot64(w.Z, w.ldz, rot_rows, tb, npg, Jg, w.skip + T.slot0, gs);
          }
        for (int t = 1; t < nt; ++t) {
          TN_CUDA(cudaEventRecord(w.ev_join[t - 1], w.aux[t - 1]));
          TN_CUDA(cudaStreamWaitEvent(s, w.ev_join[t - 1], 0));
        }
      }
    }
    for (int st = 0; st < (sched ? 0 : steps); ++st) {
      const int* tb = tab + (size_t)st * np * 2;
      // the pairs [p0, p1) of this step on stream gs: Gram -> EVD -> rotation.  (Splitting a step into pair groups on
      // separate streams, so that the EVD of one group overlaps the GEMMs of another, was measured: < 3 % -- the groups run
      // in lockstep -- and removed.)
      auto run_group = [&](int p0, int p1, int gsplit, int gchunk, cudaStream_t gs, bool marks) {
        const int npg = p1 - p0;
        Idx2 cols{JB, (long long)w.ldz, colblk, tb + 2 * p0, 2};
        cplx* Gg = w.Gpart + (size_t)p0 * JP * JP;
        cplx* Jg = w.J + (size_t)p0 * JP * JP;
        if (jgemm_enabled()) {
          jacobi_gram64(w.Z, w.ldz, jrows, tb + 2 * p0, npg, Gg, max_split(), gs);
          if (marks) prof().mark(PH_GRAM, gs);
          launch_evd(npg, Gg, Jg, tol, w.offmax, inner_sweeps, nact, w.skip + p0, gs);
          if (marks) prof().mark(PH_EVD, gs);
          jacobi_rot64(w.Z, w.ldz, jrows + (w.wonly ? 0 : w.ncols_pad), tb + 2 * p0, npg, Jg, w.skip + p0, gs);
          return;
        }
        GemmDesc g{};
        g.M = JP; g.N = JP; g.K = jrows;
        g.A = w.Z; g.am = cols; g.ak = idx1(1); g.conjA = 1;
        g.B = w.Z; g.bk = idx1(1); g.bn = cols; g.conjB = 0;
        g.C = Gg; g.cm = idx1(1); g.cn = idx1(JP);
        g.alpha = make_double2(1, 0); g.beta = make_double2(0, 0);
        g.batch = npg; g.bsA = 0; g.bsB = 0; g.bsC = (long long)JP * JP;
        // split-K partial Gram blocks are accumulated with red.global.add.f64 into one zeroed buffer (the single-CTA
        // consumers would otherwise spend tens of microseconds summing ksplit x 64 KiB partials through one SM)
        g.ksplit = gsplit; g.kchunk = gchunk; g.ssC = 0; g.atomic_c = gsplit > 1 ? 1 : 0;
        if (g.atomic_c) TN_CUDA(cudaMemsetAsync(Gg, 0, (size_t)npg * JP * JP * sizeof(cplx), gs));
        zgemm_auto(g, gs);
        if (marks) prof().mark(PH_GRAM, gs);
        launch_evd(npg, Gg, Jg, tol, w.offmax, inner_sweeps, nact, w.skip + p0, gs);
        if (marks) prof().mark(PH_EVD, gs);
        GemmDesc a{};
        a.M = jrows + (w.wonly ? 0 : w.ncols_pad); a.N = JP; a.K = JP;   // W rows (+ V rows unless W-only; ldz may carry padding rows)
        a.A = w.Z; a.am = idx1(1); a.ak = cols; a.conjA = 0;
        a.B = Jg; a.bk = idx1(1); a.bn = idx1(JP); a.conjB = 0;
        // in place (each CTA owns 128 rows x all 64 columns of its pair); pairs whose Gram block was already diagonal
        // (EVD wrote skip = 1, J = identity) are skipped, which makes the late / verification sweeps cheap
        a.C = w.Z; a.cm = idx1(1); a.cn = cols;
        a.alpha = make_double2(1, 0); a.beta = make_double2(0, 0);
        a.batch = npg; a.bsA = 0; a.bsB = (long long)JP * JP; a.bsC = 0;
        a.ksplit = 1; a.kchunk = JP; a.ssC = 0;
        a.skip = w.skip + p0;
        zgemm_auto(a, gs);
      };
      run_group(0, np, ksplit, kchunk, s, true);
      prof().mark(PH_ROT, s);
    }
    prof().flush(s);
    unsigned long long bits = 0;
    TN_CUDA(cudaMemcpyAsync(&bits, w.offmax, 8, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    double off; std::memcpy(&off, &bits, 8);
    w.last_off = off;
    w.sweeps = sweep + 1;
    // `off` was measured BEFORE this sweep's rotations; the sweeps converge quadratically, so a sweep that started below
    // ~1e-9 ends below the threshold and the Gram-only verification sweep is not needed
    if (off <= tol || off <= early_stop()) break;
  }
}

// Column-block "pairs" of the QR panels for the kernels of tn_jacobi.cu: panel pk of problem b = blocks (b nbk + 2 pk, b nbk + 2 pk + 1);
// entry ((pk * B + b) * 2) of the table.
static int* build_panel_table(int B, int nbk, cudaStream_t s) {
  const int npanels = nbk / 2;
  std::vector<int> h((size_t)npanels * B * 2);
  for (int pk = 0; pk < npanels; ++pk)
    for (int b = 0; b < B; ++b) { h[((size_t)pk * B + b) * 2] = b * nbk + 2 * pk; h[((size_t)pk * B + b) * 2 + 1] = b * nbk + 2 * pk + 1; }
  int* d = nullptr;
  TN_CUDA(cudaMalloc((void**)&d, h.size() * sizeof(int)));
  TN_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TN_CUDA(cudaStreamSynchronize(s));
  return d;
}

// Trailing tiles of the QR panels: for panel pk the tiles (b, i), i < npanels - pk - 1, = panel pk + 1 + i of problem b;
// the entries of panel pk start at offset 2 * B * (pk * npanels - pk * (pk + 1) / 2) ... computed by trail_offset().
static size_t trail_offset(int B, int npanels, int pk) { return (size_t)2 * B * ((size_t)pk * npanels - (size_t)pk * (pk + 1) / 2); }
static int* build_trail_table(int B, int nbk, cudaStream_t s) {
  const int npanels = nbk / 2;
  std::vector<int> h(trail_offset(B, npanels, npanels > 0 ? npanels - 1 : 0) + 2);
  for (int pk = 0; pk + 1 < npanels; ++pk) {
    const int ntl = npanels - pk - 1;
    int* o = h.data() + trail_offset(B, npanels, pk);
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < ntl; ++i) { o[((size_t)b * ntl + i) * 2] = b * nbk + 2 * (pk + 1 + i); o[((size_t)b * ntl + i) * 2 + 1] = b * nbk + 2 * (pk + 1 + i) + 1; }
  }
  int* d = nullptr;
  TN_CUDA(cudaMalloc((void**)&d, h.size() * sizeof(int)));
  TN_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TN_CUDA(cudaStreamSynchronize(s));
  return d;
}

static GemmDesc gd(int M, int N, int K, const cplx* A, Idx2 am, Idx2 ak, int conjA, const cplx* B, Idx2 bk, Idx2 bn, int conjB,
                   cplx* C, Idx2 cm, Idx2 cn, double alpha = 1.0, double beta = 0.0) {
  GemmDesc g{};
  g.M = M; g.N = N; g.K = K; g.A = A; g.am = am; g.ak = ak; g.conjA = conjA; g.B = B; g.bk = bk; g.bn = bn; g.conjB = conjB;
  g.C = C; g.cm = cm; g.cn = cn; g.alpha = make_double2(alpha, 0); g.beta = make_double2(beta, 0);
  g.batch = 1; g.ksplit = 1; g.kchunk = K;
  return g;
}

// One pass of right-looking block Gram-Schmidt QR with shifted-CholeskyQR3 panels (64 columns):
//   Q (rows x npad, ld = ldq) is overwritten by the orthonormal factor, R (npad x npad, ld = npad) receives the
//   upper-triangular factor.  Everything is GEMM-shaped (tn_zgemm.cu) plus the 64x64 Cholesky kernel.
static void bgs_pass(SvdWork& w, cplx* Q, long long ldq, int rows, int npad, cplx* R, int chol_passes, cudaStream_t s) {
  TN_CUDA(cudaMemsetAsync(R, 0, (size_t)npad * npad * sizeof(cplx), s));
  ensure(w.Gpart, w.G_cap, (size_t)32 * JP * JP, s);
  int ksplit = std::max(1, std::min(std::min(32, max_split()), rows / 64));   // one 64 x 64 tile per panel: spread its K range over many SMs
  int kchunk = ((rows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (rows + kchunk - 1) / kchunk;
  cplx* Rinv = w.small; cplx* Rtot = w.small + JP * JP;
  const double shift_factor = 11.0 * ((double)rows * JP + (double)JP * (JP + 1)) * 2.220446049250313e-16;
  const int npanels = npad / JP;
  const int* ptab = nullptr; const int* ttab = nullptr;
  if (jgemm_enabled()) {
    const int key = 1000000 + npad / JB;
    auto itb = w.tables.find(key);
    if (itb == w.tables.end()) itb = w.tables.emplace(key, build_panel_table(1, npad / JB, s)).first;
    ptab = itb->second;
    const int key2 = 2000000 + npad / JB;
    auto itt = w.tables.find(key2);
    if (itt == w.tables.end()) itt = w.tables.emplace(key2, build_trail_table(1, npad / JB, s)).first;
    ttab = itt->second;
  }
  for (int pk = 0; pk < npanels; ++pk) {
    cplx* P = Q + (long long)pk * JP * ldq;
    for (int it = 0; it < chol_passes; ++it) {
      // third pass of CholeskyQR3: skipped on the device when the second pass found the panel already orthonormal to 1e-3
      const int* skip3 = (chol_passes == 3 && it == 2) ? w.cflag : nullptr;
      if (ptab) {
        jacobi_gram64(Q, ldq, rows, ptab + 2 * pk, 1, w.Gpart, std::min(32, max_split()), s, skip3);
        launch_chol(1, w.Gpart, Rinv, Rtot, it, it == chol_passes - 1 ? 1 : 0, chol_passes == 3 ? shift_factor : 0.0, chol_passes == 3 ? w.cflag : nullptr, s);
        jacobi_rot64(Q, ldq, rows, ptab + 2 * pk, 1, Rinv, skip3, s);
        continue;
      }
      GemmDesc g = gd(JP, JP, rows, P, idx1(ldq), idx1(1), 1, P, idx1(1), idx1(ldq), 0, w.Gpart, idx1(1), idx1(JP));
      g.ksplit = ksplit; g.kchunk = kchunk; g.ssC = 0; g.atomic_c = ksplit > 1 ? 1 : 0;
      g.skip = skip3;
      if (g.atomic_c) TN_CUDA(cudaMemsetAsync(w.Gpart, 0, (size_t)JP * JP * sizeof(cplx), s));
      zgemm_auto(g, s);
      launch_chol(1, w.Gpart, Rinv, Rtot, it, it == chol_passes - 1 ? 1 : 0, chol_passes == 3 ? shift_factor : 0.0, chol_passes == 3 ? w.cflag : nullptr, s);
      // P <- P * Rinv (in place: each CTA owns 128 rows x all 64 columns)
      GemmDesc ap = gd(rows, JP, JP, P, idx1(1), idx1(ldq), 0, Rinv, idx1(1), idx1(JP), 0, P, idx1(1), idx1(ldq));
      ap.skip = skip3;
      zgemm_auto(ap, s);
    }
    // R(pk,pk) = Rtot
    TN_CUDA(cudaMemcpy2DAsync(R + (long long)pk * JP + (long long)pk * JP * npad, (size_t)npad * sizeof(cplx), Rtot, (size_t)JP * sizeof(cplx),
                              (size_t)JP * sizeof(cplx), JP, cudaMemcpyDeviceToDevice, s));
    int nt = npad - (pk + 1) * JP;
    if (nt > 0) {
      cplx* T = Q + (long long)(pk + 1) * JP * ldq;
      // C = P^H T  (64 x nt), split-K partials -> R(pk, trail)
      // 64 x 128 tiles, one CTA per SM: pick the split that fills one wave of 148 CTAs
      int csplit = std::max(1, std::min(std::min(std::min(16, max_split()), rows / 64), 148 / ((nt + 127) / 128)));
      int cchunk = ((rows + csplit - 1) / csplit + 7) / 8 * 8;
      csplit = (rows + cchunk - 1) / cchunk;
      // split-K contributions go straight into the block row of R (zeroed by the memset above) with red.global.add.f64
      cplx* Rrow = R + (long long)pk * JP + (long long)(pk + 1) * JP * npad;
      if (ttab && cross_enabled()) jacobi_cross64(Q, ldq, ptab + 2 * pk, Q, ldq, rows, ttab + trail_offset(1, npanels, pk), 1, nt / JP, Rrow, npad, 0, std::min(16, max_split()), s);
      else {
        GemmDesc c = gd(JP, nt, rows, P, idx1(ldq), idx1(1), 1, T, idx1(1), idx1(ldq), 0, Rrow, idx1(1), idx1(npad));
        c.ksplit = csplit; c.kchunk = cchunk; c.ssC = 0; c.atomic_c = 1;
        zgemm_auto(c, s);
      }
      // T <- T - P C
      if (ttab) jacobi_update64(Q, ldq, ptab + 2 * pk, Q, ldq, rows, ttab + trail_offset(1, npanels, pk), 1, nt / JP, Rrow, npad, 0, s);
      else zgemm_auto(gd(rows, nt, JP, P, idx1(1), idx1(ldq), 0, Rrow, idx1(1), idx1(npad), 0, T, idx1(1), idx1(ldq), -1.0, 1.0), s);
    }
  }
}

// Q <- orthonormal factor, Ra <- R with A = Q R; two Gram-Schmidt passes ("twice is enough"), R = R'' R'.
static void bgs_qr(SvdWork& w, cplx* Q, long long ldq, int rows, int npad, cudaStream_t s) {
  ensure(w.Ra, w.Ra_cap, (size_t)npad * npad, s);
  ensure(w.Rb, w.Rb_cap, (size_t)npad * npad, s);
  ensure(w.Rc, w.Rc_cap, (size_t)npad * npad, s);
  bgs_pass(w, Q, ldq, rows, npad, w.Rc, 3, s);    // R'  : shifted CholeskyQR3 panels (any conditioning)
  // R'' : the panels are orthonormal already and only lose O(delta) to the re-projection -> one (TN_SVD_REORTH_CHOL=2: two)
  // Cholesky pass per panel restores orthonormality to O(eps)
  static int reorth = -1;
  if (reorth < 0) { const char* e = getenv("TN_SVD_REORTH_CHOL"); reorth = (e && e[0] == '2') ? 2 : 1; }
  bgs_pass(w, Q, ldq, rows, npad, w.Rb, reorth, s);
  zgemm_auto(gd(npad, npad, npad, w.Rb, idx1(1), idx1(npad), 0, w.Rc, idx1(1), idx1(npad), 0, w.Ra, idx1(1), idx1(npad)), s);
}

// Leading dimension of Z: a column stride that is a large power of two (2*npad*16 B = 64 KiB at n = 2048) maps the
// k-walk of the Gram / rotation GEMMs onto few DRAM channels; 8 padding rows (128 B) break the pattern.
static int pad_ld(int rows) {
  static int pad = -1;
  if (pad < 0) { const char* e = getenv("TN_SVD_LDPAD"); pad = e ? atoi(e) : 8; }
  return (rows % 64 == 0) ? rows + pad : rows;
}

static int batcher_submit(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso, bool qr);
static thread_local SvdBatcher* tl_batcher = nullptr;

static bool gauge_qr_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_GAUGE_QR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

int svd_factor(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso, bool need_values) {
  TN_CHECK(m >= 1 && n >= 1, "svd: empty matrix");
  w.use_view = false;
  w.qr_mode = 0;
  {
    // the reference's rule (tensors.jl:201-215) keeps all min(m, n) values when cutoff == 0 and maxdim is 0 or >= min(m, n)
    const long long nsv = std::min(m, n);
    if (!need_values && iso != 0 && gauge_qr_enabled() && tr.cutoff == 0.0 && (tr.maxdim == 0 || tr.maxdim >= nsv)) {
      w.m = m; w.n = n; w.k = (int)nsv; w.nsv = (int)nsv; w.sweeps = 0; w.wonly = false; w.precond = false;
      const bool iso_long = (iso == 1 && m >= n) || (iso == 2 && n >= m);
      if (!iso_long) { w.qr_mode = 2; w.M0 = M; w.ld0 = ld; w.transposed = false; return w.k; }
      if (tl_batcher != nullptr) return batcher_submit(w, M, m, n, ld, tr, s, iso, true);    // the QR steps of all trajectories in one batched launch sequence
      w.qr_mode = 1;
      w.transposed = iso == 2;
      w.rows = w.transposed ? n : m;
      w.ncols = w.transposed ? m : n;
      w.ncols_pad = ((w.ncols + JP - 1) / JP) * JP;
      TN_CHECK(w.ncols_pad <= 8192, "svd: more than 8192 columns is not supported yet");
      const int npad = w.ncols_pad;
      if (!w.offmax) { TN_CUDA(cudaMalloc((void**)&w.offmax, 8)); TN_CUDA(cudaMalloc((void**)&w.kout, 4)); TN_CUDA(cudaMalloc((void**)&w.cflag, 4)); TN_CUDA(cudaMalloc((void**)&w.small, 3 * JP * JP * sizeof(cplx))); }
      ensure(w.Q1, w.Q1_cap, (size_t)w.rows * npad, s);
      int blocks;
      launch_1d((long long)w.rows * npad, blocks);
      svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Q1, w.rows, w.ncols, npad, w.rows);
      TN_CUDA(cudaGetLastError());
      count_launch(1);
      bgs_qr(w, w.Q1, w.rows, w.rows, npad, s);                  // T = Q1 Ra
      return w.k;
    }
  }
  if (tl_batcher != nullptr) return batcher_submit(w, M, m, n, ld, tr, s, iso, false);
  w.m = m; w.n = n;
  w.transposed = m < n || (m == n && iso == 2);       // a square matrix is factorised in the orientation whose long-side factor is the isometry
  w.rows = w.transposed ? n : m;
  w.ncols = w.transposed ? m : n;
  w.nsv = w.ncols;
  w.ncols_pad = ((w.ncols + JP - 1) / JP) * JP;
  TN_CHECK(w.ncols_pad <= 8192, "svd: more than 8192 columns is not supported yet");
  const int npad = w.ncols_pad;
  w.precond = precond_enabled() && npad > JP;
  w.wonly = w.precond && wonly_enabled() && ((iso == 1 && !w.transposed) || (iso == 2 && w.transposed));
  if (w.s_cap < (size_t)npad) {
    if (w.sig) { TN_CUDA(cudaFreeAsync(w.sig, s)); TN_CUDA(cudaFreeAsync(w.perm, s)); TN_CUDA(cudaFreeAsync(w.sig2, s)); }
    TN_CUDA(cudaMallocAsync((void**)&w.sig, npad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.sig2, npad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.perm, npad * sizeof(int), s));
    w.s_cap = npad;
  }
  if (!w.offmax) { TN_CUDA(cudaMalloc((void**)&w.offmax, 8)); TN_CUDA(cudaMalloc((void**)&w.kout, 4)); TN_CUDA(cudaMalloc((void**)&w.cflag, 4)); TN_CUDA(cudaMalloc((void**)&w.small, 3 * JP * JP * sizeof(cplx))); }
  int blocks;
  SvdProf& pf = prof();
  if (pf.on) { for (double& m : pf.ms) m = 0; pf.used = 0; pf.mark(PH_START, s); }
  if (!w.precond && w.ncols <= small_max_cols() && w.rows <= SMALL_MAX_ROWS && small_enabled()) {
    // small matrices (default: at most 32 columns, 128 rows): the whole factorisation (init, sweeps, norms, sort, truncation rank) in one
    // single-CTA launch.  Measured on a B200: 22 x 22 inside the C1 sweep 146 us against ~410 us for the general path; plain one-sided
    // Jacobi needs 17-28 sweeps on a graded spectrum without the QR preconditioner (22: 0.45 ms, 64: 4.2 ms), so 64 columns stay
    // on the general path (a de Rijk column swap made it worse in the parallel ordering: 23-43 sweeps)
    w.jrows = w.rows;
    w.ldz = pad_ld(w.rows + npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * npad, s);
    const int nact = (w.ncols + 1) & ~1;
    const size_t smem = ((size_t)w.rows * nact + (size_t)nact * nact) * sizeof(cplx);
    static DeviceOnce small_cfg;
    small_cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(small_svd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)SMALL_MAX_ROWS * JP + (size_t)JP * JP) * sizeof(cplx)))); });
    const double tol = 3.0 * std::sqrt((double)w.rows) * 2.220446049250313e-16;
    small_svd_kernel<<<1, 32 * std::max(2, nact / 2), smem, s>>>(M, ld, w.transposed ? 1 : 0, w.rows, w.ncols, w.Z, w.ldz, w.sig, w.perm, w.nsv,
                                                              tr.cutoff, tr.maxdim, tr.mindim, w.kout, w.cflag, tol);
    TN_CUDA(cudaGetLastError());
    count_launch(1);
    int ks[2] = {0, 0};
    TN_CUDA(cudaMemcpyAsync(&ks[0], w.kout, 4, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(&ks[1], w.cflag, 4, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    w.k = ks[0]; w.sweeps = ks[1];
    if (pf.on) { pf.mark(PH_FIN, s); pf.flush(s); }
    return w.k;
  }
  if (!w.precond) {
    w.jrows = w.rows;
    w.ldz = pad_ld(w.rows + npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * npad, s);
    launch_1d((long long)w.ldz * npad, blocks);
    svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Z, w.rows, w.ncols, npad, w.ldz);
    TN_CUDA(cudaGetLastError());
    count_launch(1);
    jacobi_sweeps(w, w.rows, s);
  } else {
    // A = Q1 R1 ; R1^H = Q2 R2 ; X = R2^H ;  X V' = U' Sigma  =>  A = (Q1 U') Sigma (Q2 V')^H
    ensure(w.Q1, w.Q1_cap, (size_t)w.rows * npad, s);
    launch_1d((long long)w.rows * npad, blocks);
    svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Q1, w.rows, w.ncols, npad, w.rows);   // ldz = rows: no identity part
    count_launch(1);
    bgs_qr(w, w.Q1, w.rows, w.rows, npad, s);                    // Ra = R1
    if (w.wonly) {                                               // keep R1: S V_T^H = U'^H R1 (see tn_svd.cuh)
      ensure(w.R1, w.R1_cap, (size_t)npad * npad, s);
      TN_CUDA(cudaMemcpyAsync(w.R1, w.Ra, (size_t)npad * npad * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
    }
    ensure(w.Q2, w.Q2_cap, (size_t)npad * npad, s);
    launch_1d((long long)npad * npad, blocks);
    conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra, npad, npad, npad, w.Q2, npad);   // Q2 <- R1^H
    count_launch(1);
    bgs_qr(w, w.Q2, npad, npad, npad, s);                        // Ra = R2
    w.jrows = npad;
    w.ldz = pad_ld(w.wonly ? npad : 2 * npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * npad, s);
    conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra, npad, npad, npad, w.Z, w.ldz);   // W <- R2^H
    if (!w.wonly) set_identity_kernel<<<blocks, 256, 0, s>>>(w.Z + npad, w.ldz, npad);
    TN_CUDA(cudaGetLastError());
    count_launch(2);
    pf.mark(PH_QR, s);
    jacobi_sweeps(w, npad, s);
  }
  colnorm2_kernel<<<npad, 128, 0, s>>>(w.Z, w.jrows, w.ldz, w.sig2);
  int npow2 = 64; while (npow2 < npad) npow2 <<= 1;
  static DeviceOnce sort_cfg;
  sort_cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(sort_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12)); });
  sort_trunc_kernel<<<1, 1024, npow2 * 12, s>>>(w.sig2, npad, npow2, w.sig, w.perm, w.nsv, tr.cutoff, tr.maxdim, tr.mindim, w.kout);
  TN_CUDA(cudaGetLastError());
  count_launch(2);
  int k = 0;
  TN_CUDA(cudaMemcpyAsync(&k, w.kout, 4, cudaMemcpyDeviceToHost, s));
  TN_CUDA(cudaStreamSynchronize(s));
  w.k = k;
  if (pf.on) {
    pf.mark(PH_FIN, s); pf.flush(s);
    fprintf(stderr, "{\"svd_profile\": {\"m\": %d, \"n\": %d, \"npad\": %d, \"precond\": %d, \"sweeps\": %d, \"qr_ms\": %.3f, \"gram_ms\": %.3f, "
            "\"evd_ms\": %.3f, \"rot_ms\": %.3f, \"finish_ms\": %.3f}}\n", m, n, npad, (int)w.precond, w.sweeps, pf.ms[PH_QR], pf.ms[PH_GRAM],
            pf.ms[PH_EVD], pf.ms[PH_ROT], pf.ms[PH_FIN]);
  }
  return k;
}

static void gather(const cplx* src, int lds, int rows, int k, const int* perm, const double* sig, int mode, int tr, cplx* out, long long ldo, cudaStream_t s) {
  long long total = (long long)rows * k;
  int blocks = std::max(1, (int)std::min<long long>(148 * 8, (total + 255) / 256));
  gather_kernel<<<blocks, 256, 0, s>>>(src, lds, rows, k, perm, sig, mode, tr, out, ldo);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

// With the preconditioner the Jacobi factors live in the small n x n problem:
//   left vectors of A:  Q1 * (W/sigma)(:,perm)   right vectors of A:  Q2 * V(:,perm)
static void precond_left(SvdWork& w, int scale_mode, cplx* out, long long ldo, bool conj_transposed, cudaStream_t s) {
  const int npad = w.ncols_pad, k = w.k;
  ensure(w.Tg, w.Tg_cap, (size_t)npad * std::max(k, 1), s);
  gather(w.Z, w.ldz, npad, k, w.perm, w.sig, scale_mode, 0, w.Tg, npad, s);                 // (W * scale)(:, perm)
  if (!conj_transposed) zgemm_auto(gd(w.rows, k, npad, w.Q1, idx1(1), idx1(w.rows), 0, w.Tg, idx1(1), idx1(npad), 0, out, idx1(1), idx1(ldo)), s);
  else zgemm_auto(gd(k, w.rows, npad, w.Tg, idx1(npad), idx1(1), 1, w.Q1, idx1(w.rows), idx1(1), 1, out, idx1(1), idx1(ldo)), s);
}
static void precond_right(SvdWork& w, int scale_mode, cplx* out, long long ldo, bool conj_transposed, cudaStream_t s) {
  const int npad = w.ncols_pad, k = w.k;
  ensure(w.Tg, w.Tg_cap, (size_t)npad * std::max(k, 1), s);
  gather(w.Z + npad, w.ldz, npad, k, w.perm, w.sig, scale_mode, 0, w.Tg, npad, s);          // (V * scale)(:, perm)
  if (!conj_transposed) zgemm_auto(gd(w.ncols, k, npad, w.Q2, idx1(1), idx1(npad), 0, w.Tg, idx1(1), idx1(npad), 0, out, idx1(1), idx1(ldo)), s);
  else zgemm_auto(gd(k, w.ncols, npad, w.Tg, idx1(npad), idx1(1), 1, w.Q2, idx1(npad), idx1(1), 1, out, idx1(1), idx1(ldo)), s);
}

// W-only mode: the factor on the short side of the tall orientation T, [S] V_T^H = (W [/sigma^2 | /sigma])(:,perm)^H R1 (k x ncols),
// or its conjugate transpose (ncols x k).  Without S the rows are divided by sigma_j (callers of the W-only mode ask for S).
static void wonly_other(SvdWork& w, bool times_S, cplx* out, long long ldo, bool conj_transposed, cudaStream_t s) {
  const int npad = w.ncols_pad, k = w.k;
  ensure(w.Tg, w.Tg_cap, (size_t)npad * std::max(k, 1), s);
  gather(w.Z, w.ldz, npad, k, w.perm, w.sig, times_S ? 2 : 3, 0, w.Tg, npad, s);
  if (!conj_transposed) zgemm_auto(gd(k, w.ncols, npad, w.Tg, idx1(npad), idx1(1), 1, w.R1, idx1(1), idx1(npad), 0, out, idx1(1), idx1(ldo)), s);
  else zgemm_auto(gd(w.ncols, k, npad, w.R1, idx1(npad), idx1(1), 1, w.Tg, idx1(1), idx1(npad), 0, out, idx1(1), idx1(ldo)), s);
}

// Factors of a gauge move that was not an SVD (SvdWork::qr_mode).  want_U: the m x k factor, else the k x n factor; which of
// the two is the isometry was fixed by the `iso` argument of svd_factor.
static void qr_gather(SvdWork& w, bool want_U, cplx* out, long long ldo, cudaStream_t s) {
  const int k = w.k, m = w.m, n = w.n;
  int blocks;
  if (w.qr_mode == 2) {
    // isometry on the short side: identity there, the matrix itself on the other side (k = min(m, n))
    const bool u_is_identity = m <= n;
    if (want_U == u_is_identity) {
      TN_CHECK(ldo == k, "gauge move: unexpected leading dimension");
      launch_1d((long long)k * k, blocks);
      set_identity_kernel<<<blocks, 256, 0, s>>>(out, ldo, k);
      TN_CUDA(cudaGetLastError());
      count_launch(1);
    } else {
      TN_CUDA(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(cplx), w.M0, (size_t)w.ld0 * sizeof(cplx), (size_t)m * sizeof(cplx), n, cudaMemcpyDeviceToDevice, s));
    }
    return;
  }
  const int npad = w.ncols_pad;
  if (!w.transposed) {          // A = Q1 R: U = Q1(:, :k), [S V^H] = R(:k, :n)
    if (want_U) TN_CUDA(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(cplx), w.Q1, (size_t)w.rows * sizeof(cplx), (size_t)m * sizeof(cplx), k, cudaMemcpyDeviceToDevice, s));
    else TN_CUDA(cudaMemcpy2DAsync(out, (size_t)ldo * sizeof(cplx), w.Ra, (size_t)npad * sizeof(cplx), (size_t)k * sizeof(cplx), n, cudaMemcpyDeviceToDevice, s));
  } else {                      // A^H = Q1 R: [U S] = R^H (m x k), V^H = Q1^H (k x n)
    if (want_U) {
      launch_1d((long long)k * m, blocks);
      conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra, npad, k, m, out, ldo);
    } else {
      launch_1d((long long)n * k, blocks);
      conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Q1, w.rows, n, k, out, ldo);
    }
    TN_CUDA(cudaGetLastError());
    count_launch(1);
  }
}

void svd_gather_U(SvdWork& w, cplx* U, long long ldu, bool times_S, cudaStream_t s) {
  if (w.use_view) {      // factors of a batching round: gather from this thread's slot, with this thread's scratch
    SvdWork& v = *w.bview; v.Tg = w.Tg; v.Tg_cap = w.Tg_cap;
    svd_gather_U(v, U, ldu, times_S, s);
    w.Tg = v.Tg; w.Tg_cap = v.Tg_cap;
    return;
  }
  if (w.qr_mode != 0) { qr_gather(w, true, U, ldu, s); return; }
  if (w.precond) {
    if (!w.transposed) precond_left(w, times_S ? 0 : 2, U, ldu, false, s);       // Q1 (W[/sigma])
    else if (w.wonly) wonly_other(w, times_S, U, ldu, true, s);                  // ([S] V_T^H)^H
    else precond_right(w, times_S ? 1 : 0, U, ldu, false, s);                    // Q2 (V[*sigma])
    return;
  }
  if (!w.transposed) gather(w.Z, w.ldz, w.m, w.k, w.perm, w.sig, times_S ? 0 : 2, 0, U, ldu, s);          // W / sigma
  else gather(w.Z + w.rows, w.ldz, w.m, w.k, w.perm, w.sig, times_S ? 1 : 0, 0, U, ldu, s);                // V' (* sigma)
}
void svd_gather_Vh(SvdWork& w, cplx* Vh, long long ldv, bool times_S, cudaStream_t s) {
  if (w.use_view) {
    SvdWork& v = *w.bview; v.Tg = w.Tg; v.Tg_cap = w.Tg_cap;
    svd_gather_Vh(v, Vh, ldv, times_S, s);
    w.Tg = v.Tg; w.Tg_cap = v.Tg_cap;
    return;
  }
  if (w.qr_mode != 0) { qr_gather(w, false, Vh, ldv, s); return; }
  if (w.precond) {
    if (!w.transposed && w.wonly) wonly_other(w, times_S, Vh, ldv, false, s);    // [S] V_T^H = U'^H R1
    else if (!w.transposed) precond_right(w, times_S ? 1 : 0, Vh, ldv, true, s); // (Q2 V[*sigma])^H
    else precond_left(w, times_S ? 0 : 2, Vh, ldv, true, s);                     // (Q1 W[/sigma])^H
    return;
  }
  if (!w.transposed) gather(w.Z + w.rows, w.ldz, w.n, w.k, w.perm, w.sig, times_S ? 1 : 0, 1, Vh, ldv, s); // conj(V)^T (* sigma)
  else gather(w.Z, w.ldz, w.n, w.k, w.perm, w.sig, times_S ? 0 : 2, 1, Vh, ldv, s);                        // conj(W'/sigma)^T
}
void svd_copy_S(SvdWork& w, double* S, cudaStream_t s) {
  const double* sig = w.use_view ? w.bview->sig : w.sig;
  TN_CUDA(cudaMemcpyAsync(S, sig, (size_t)w.k * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void svd_free(SvdWork& w) {
  if (w.Z) cudaFree(w.Z);
  if (w.skip) cudaFree(w.skip);
  if (w.dtab) cudaFree(w.dtab);
  if (w.Gpart) cudaFree(w.Gpart);
  if (w.J) cudaFree(w.J);
  if (w.sig) { cudaFree(w.sig); cudaFree(w.sig2); cudaFree(w.perm); }
  if (w.offmax) { cudaFree(w.offmax); cudaFree(w.kout); cudaFree(w.small); cudaFree(w.cflag); }
  for (cplx* p : {w.Q1, w.Q2, w.Ra, w.Rb, w.Rc, w.Tg, w.R1}) if (p) cudaFree(p);
  for (auto& kv : w.tables) cudaFree(kv.second);
  if (w.split) { if (w.split->dev) cudaFree(w.split->dev); delete w.split; }
  for (int i = 0; i < 3; ++i) { if (w.aux[i]) cudaStreamDestroy(w.aux[i]); if (w.ev_join[i]) cudaEventDestroy(w.ev_join[i]); }
  if (w.ev_fork) cudaEventDestroy(w.ev_fork);
  delete w.bview;
  w = SvdWork{};
}

// ================================================================================================
// Factorisation in three calls for a caller that owns the Jacobi pair schedule (the distributed sweeps of tnb200/sharded.py:
// every rank prepares the same Z, rotates the column-block pairs its schedule assigns to it, exchanges column blocks with its
// peers between the steps, and finishes on the complete Z):
//   svd_dist_begin  = svd_factor up to, not including, the Jacobi sweeps (init, the two QR steps, Z = [R2^H ; I])
//   svd_dist_step   = one Gram -> EVD -> rotation launch sequence over a caller-supplied list of column-block pairs
//   svd_dist_finish = column norms, sort, truncation rule (the tail of svd_factor); the gathers then work as usual
// (A copy of the corresponding parts of svd_factor / jacobi_sweeps, kept separate until it has been validated on the GPU.)
// ================================================================================================
int svd_dist_begin(SvdWork& w, const cplx* M, int m, int n, long long ld, cudaStream_t s) {
  TN_CHECK(m >= 1 && n >= 1, "svd: empty matrix");
  w.use_view = false;
  w.m = m; w.n = n;
  w.transposed = m < n;
  w.rows = w.transposed ? n : m;
  w.ncols = w.transposed ? m : n;
  w.nsv = w.ncols;
  w.ncols_pad = ((w.ncols + JP - 1) / JP) * JP;
  TN_CHECK(w.ncols_pad <= 8192, "svd: more than 8192 columns is not supported yet");
  const int npad = w.ncols_pad;
  w.precond = precond_enabled() && npad > JP;
  w.wonly = false;
  if (w.s_cap < (size_t)npad) {
    if (w.sig) { TN_CUDA(cudaFreeAsync(w.sig, s)); TN_CUDA(cudaFreeAsync(w.perm, s)); TN_CUDA(cudaFreeAsync(w.sig2, s)); }
    TN_CUDA(cudaMallocAsync((void**)&w.sig, npad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.sig2, npad * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.perm, npad * sizeof(int), s));
    w.s_cap = npad;
  }
  if (!w.offmax) { TN_CUDA(cudaMalloc((void**)&w.offmax, 8)); TN_CUDA(cudaMalloc((void**)&w.kout, 4)); TN_CUDA(cudaMalloc((void**)&w.cflag, 4)); TN_CUDA(cudaMalloc((void**)&w.small, 3 * JP * JP * sizeof(cplx))); }
  int blocks;
  if (!w.precond) {
    w.jrows = w.rows;
    w.ldz = pad_ld(w.rows + npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * npad, s);
    launch_1d((long long)w.ldz * npad, blocks);
    svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Z, w.rows, w.ncols, npad, w.ldz);
    TN_CUDA(cudaGetLastError());
    count_launch(1);
  } else {
    ensure(w.Q1, w.Q1_cap, (size_t)w.rows * npad, s);
    launch_1d((long long)w.rows * npad, blocks);
    svd_init_kernel<<<blocks, 256, 0, s>>>(M, ld, m, n, w.transposed ? 1 : 0, w.Q1, w.rows, w.ncols, npad, w.rows);
    count_launch(1);
    bgs_qr(w, w.Q1, w.rows, w.rows, npad, s);
    ensure(w.Q2, w.Q2_cap, (size_t)npad * npad, s);
    launch_1d((long long)npad * npad, blocks);
    conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra, npad, npad, npad, w.Q2, npad);
    count_launch(1);
    bgs_qr(w, w.Q2, npad, npad, npad, s);
    w.jrows = npad;
    w.ldz = pad_ld(2 * npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * npad, s);
    conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra, npad, npad, npad, w.Z, w.ldz);
    set_identity_kernel<<<blocks, 256, 0, s>>>(w.Z + npad, w.ldz, npad);
    TN_CUDA(cudaGetLastError());
    count_launch(2);
  }
  w.sweeps = 0;
  TN_CUDA(cudaStreamSynchronize(s));
  return npad / JB;
}

double svd_dist_tol(const SvdWork& w) { return 3.0 * std::sqrt((double)w.jrows) * 2.220446049250313e-16; }

// pairs_host: npairs x 2 column-block indices (disjoint blocks).  Returns max |g_pq| / sqrt(g_pp g_qq) over these pairs before the rotation.
double svd_dist_step(SvdWork& w, const int* pairs_host, int npairs, cudaStream_t s) {
  TN_CHECK(npairs >= 1 && w.Z != nullptr, "svd step: nothing to do / factorisation not begun");
  const int jrows = w.jrows, nb = w.ncols_pad / JB;
  for (int i = 0; i < 2 * npairs; ++i) TN_CHECK(pairs_host[i] >= 0 && pairs_host[i] < nb, "svd step: column block out of range");
  int ksplit = std::max(1, std::min(std::min(std::min(16, max_split()), jrows / 64), (2 * 148) / npairs));
  int kchunk = ((jrows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (jrows + kchunk - 1) / kchunk;
  ensure(w.Gpart, w.G_cap, (size_t)std::max(npairs, 32) * JP * JP, s);
  ensure(w.J, w.J_cap, (size_t)npairs * JP * JP, s);
  ensure(w.skip, w.skip_cap, (size_t)npairs, s);
  ensure(w.dtab, w.dtab_cap, (size_t)2 * npairs, s);
  TN_CUDA(cudaMemcpyAsync(w.dtab, pairs_host, (size_t)2 * npairs * sizeof(int), cudaMemcpyHostToDevice, s));
  const double tol = svd_dist_tol(w);
  const long long colblk = (long long)JB * w.ldz;
  TN_CUDA(cudaMemsetAsync(w.offmax, 0, 8, s));
  Idx2 cols{JB, (long long)w.ldz, colblk, w.dtab, 2};
  if (jgemm_enabled()) {
    jacobi_gram64(w.Z, w.ldz, jrows, w.dtab, npairs, w.Gpart, max_split(), s);
    launch_evd(npairs, w.Gpart, w.J, tol, w.offmax, 1, JP, w.skip, s);
    jacobi_rot64(w.Z, w.ldz, jrows + w.ncols_pad, w.dtab, npairs, w.J, w.skip, s);
    unsigned long long bits0 = 0;
    TN_CUDA(cudaMemcpyAsync(&bits0, w.offmax, 8, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    double off0; std::memcpy(&off0, &bits0, 8);
    return off0;
  }
  GemmDesc g{};
  g.M = JP; g.N = JP; g.K = jrows;
  g.A = w.Z; g.am = cols; g.ak = idx1(1); g.conjA = 1;
  g.B = w.Z; g.bk = idx1(1); g.bn = cols; g.conjB = 0;
  g.C = w.Gpart; g.cm = idx1(1); g.cn = idx1(JP);
  g.alpha = make_double2(1, 0); g.beta = make_double2(0, 0);
  g.batch = npairs; g.bsA = 0; g.bsB = 0; g.bsC = (long long)JP * JP;
  g.ksplit = ksplit; g.kchunk = kchunk; g.ssC = 0; g.atomic_c = ksplit > 1 ? 1 : 0;
  if (g.atomic_c) TN_CUDA(cudaMemsetAsync(w.Gpart, 0, (size_t)npairs * JP * JP * sizeof(cplx), s));
  zgemm_auto(g, s);
  launch_evd(npairs, w.Gpart, w.J, tol, w.offmax, 1, JP, w.skip, s);
  GemmDesc a{};
  a.M = jrows + w.ncols_pad; a.N = JP; a.K = JP;
  a.A = w.Z; a.am = idx1(1); a.ak = cols; a.conjA = 0;
  a.B = w.J; a.bk = idx1(1); a.bn = idx1(JP); a.conjB = 0;
  a.C = w.Z; a.cm = idx1(1); a.cn = cols;
  a.alpha = make_double2(1, 0); a.beta = make_double2(0, 0);
  a.batch = npairs; a.bsA = 0; a.bsB = (long long)JP * JP; a.bsC = 0;
  a.ksplit = 1; a.kchunk = JP; a.ssC = 0;
  a.skip = w.skip;
  zgemm_auto(a, s);
  unsigned long long bits = 0;
  TN_CUDA(cudaMemcpyAsync(&bits, w.offmax, 8, cudaMemcpyDeviceToHost, s));
  TN_CUDA(cudaStreamSynchronize(s));
  double off; std::memcpy(&off, &bits, 8);
  return off;
}

int svd_dist_finish(SvdWork& w, Trunc tr, int sweeps, cudaStream_t s) {
  const int npad = w.ncols_pad;
  colnorm2_kernel<<<npad, 128, 0, s>>>(w.Z, w.jrows, w.ldz, w.sig2);
  int npow2 = 64; while (npow2 < npad) npow2 <<= 1;
  TN_CUDA(cudaFuncSetAttribute(sort_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
  sort_trunc_kernel<<<1, 1024, npow2 * 12, s>>>(w.sig2, npad, npow2, w.sig, w.perm, w.nsv, tr.cutoff, tr.maxdim, tr.mindim, w.kout);
  TN_CUDA(cudaGetLastError());
  count_launch(2);
  int k = 0;
  TN_CUDA(cudaMemcpyAsync(&k, w.kout, 4, cudaMemcpyDeviceToHost, s));
  TN_CUDA(cudaStreamSynchronize(s));
  w.k = k; w.sweeps = sweeps;
  return k;
}

// ================================================================================================
// Batched factorisation: B matrices of the same shape (the gate / gauge-move SVDs of B QJMC trajectories that advance in
// lockstep; SURVEY 2b K5 "batched variant").  The problems are stacked side by side -- Z = [Z_0 | Z_1 | ...], Q1, Q2, R alike --
// so every launch of the single-problem pipeline becomes ONE launch over B problems: the GEMMs get a batch dimension (problem
// stride), the Jacobi pair table lists the column-block pairs of all problems (B * np CTAs per EVD launch instead of np: at
// n = 512 one problem occupies 8 of 148 SMs), the Cholesky / sort kernels run one CTA per problem.  All problems sweep until the
// slowest has converged (pairs that are already diagonal are skipped by the rotation GEMM).  The per-problem factors are read
// through the ordinary gathers on a view of the stacked workspace.
// ================================================================================================
static const int* pair_table_b(SvdBatch& w, int B, int nb, cudaStream_t s) {
  auto key = std::make_pair(B, nb);
  auto it = w.tables.find(key);
  if (it != w.tables.end()) return it->second;
  const int steps = nb - 1, np = nb / 2;
  std::vector<int> h((size_t)steps * B * np * 2);
  for (int st = 0; st < steps; ++st)
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < np; ++k) {
        int a, c;
        if (nb == 2) { a = 0; c = 1; }
        else if (k == 0) { a = nb - 1; c = st; }
        else { a = (st + k) % (nb - 1); c = (st - k + (nb - 1)) % (nb - 1); }
        const size_t e = (((size_t)st * B + b) * np + k) * 2;
        h[e + 0] = b * nb + std::min(a, c);
        h[e + 1] = b * nb + std::max(a, c);
      }
  int* d = nullptr;
  TN_CUDA(cudaMalloc((void**)&d, h.size() * sizeof(int)));
  TN_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TN_CUDA(cudaStreamSynchronize(s));
  w.tables[key] = d;
  return d;
}

static void jacobi_sweeps_b(SvdBatch& w, int jrows, cudaStream_t s) {
  const int B = w.B, nb = w.npad / JB, np = nb / 2, steps = nb - 1, npg = B * np;
  // Gram GEMM grid = B * np pairs x ksplit: about one wave of the 2 x 148 CTA slots
  int ksplit = std::max(1, std::min(std::min(std::min(16, max_split()), jrows / 64), (2 * 148) / npg));
  int kchunk = ((jrows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (jrows + kchunk - 1) / kchunk;
  ensure(w.Gpart, w.G_cap, (size_t)npg * JP * JP, s);
  ensure(w.J, w.J_cap, (size_t)npg * JP * JP, s);
  ensure(w.skip, w.skip_cap, (size_t)npg, s);
  const int* tab = pair_table_b(w, B, nb, s);
  const double tol = 3.0 * std::sqrt((double)jrows) * 2.220446049250313e-16;
  const long long colblk = (long long)JB * w.ldz;
  const int inner_sweeps = (np == 1) ? 12 : 1;
  const int nact = (np == 1) ? std::max(2, std::min(JP, (w.ncols + 1) / 2 * 2)) : JP;
  w.sweeps = 0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    TN_CUDA(cudaMemsetAsync(w.offmax, 0, 8, s));
    for (int st = 0; st < steps; ++st) {
      Idx2 cols{JB, (long long)w.ldz, colblk, tab + (size_t)st * npg * 2, 2};
      if (jgemm_enabled() && npg <= 65535) {
        const int* tb = tab + (size_t)st * npg * 2;
        jacobi_gram64(w.Z, w.ldz, jrows, tb, npg, w.Gpart, max_split(), s);
        launch_evd(npg, w.Gpart, w.J, tol, w.offmax, inner_sweeps, nact, w.skip, s);
        jacobi_rot64(w.Z, w.ldz, jrows + (w.wonly ? 0 : w.npad), tb, npg, w.J, w.skip, s);
        continue;
      }
      GemmDesc g{};
      g.M = JP; g.N = JP; g.K = jrows;
      g.A = w.Z; g.am = cols; g.ak = idx1(1); g.conjA = 1;
      g.B = w.Z; g.bk = idx1(1); g.bn = cols; g.conjB = 0;
      g.C = w.Gpart; g.cm = idx1(1); g.cn = idx1(JP);
      g.alpha = make_double2(1, 0); g.beta = make_double2(0, 0);
      g.batch = npg; g.bsA = 0; g.bsB = 0; g.bsC = (long long)JP * JP;
      g.ksplit = ksplit; g.kchunk = kchunk; g.ssC = 0; g.atomic_c = ksplit > 1 ? 1 : 0;
      if (g.atomic_c) TN_CUDA(cudaMemsetAsync(w.Gpart, 0, (size_t)npg * JP * JP * sizeof(cplx), s));
      zgemm_auto(g, s);
      launch_evd(npg, w.Gpart, w.J, tol, w.offmax, inner_sweeps, nact, w.skip, s);
      GemmDesc a{};
      a.M = jrows + (w.wonly ? 0 : w.npad); a.N = JP; a.K = JP;
      a.A = w.Z; a.am = idx1(1); a.ak = cols; a.conjA = 0;
      a.B = w.J; a.bk = idx1(1); a.bn = idx1(JP); a.conjB = 0;
      a.C = w.Z; a.cm = idx1(1); a.cn = cols;
      a.alpha = make_double2(1, 0); a.beta = make_double2(0, 0);
      a.batch = npg; a.bsA = 0; a.bsB = (long long)JP * JP; a.bsC = 0;
      a.ksplit = 1; a.kchunk = JP; a.ssC = 0;
      a.skip = w.skip;
      zgemm_auto(a, s);
    }
    unsigned long long bits = 0;
    TN_CUDA(cudaMemcpyAsync(&bits, w.offmax, 8, cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaStreamSynchronize(s));
    double off; std::memcpy(&off, &bits, 8);
    w.sweeps = sweep + 1;
    if (off <= tol || off <= early_stop()) break;
  }
}

// Right-looking block Gram-Schmidt pass (bgs_pass) over B stacked problems: Q = [Q_0 | Q_1 | ...] (rows x B*npad, ld = ldq),
// R = [R_0 | R_1 | ...] (npad x B*npad, ld = npad).
static void bgs_pass_b(SvdBatch& w, cplx* Q, long long ldq, int rows, int npad, cplx* R, int chol_passes, cudaStream_t s) {
  const int B = w.B;
  const long long qs = (long long)npad * ldq, rs = (long long)npad * npad, gs = (long long)JP * JP;
  TN_CUDA(cudaMemsetAsync(R, 0, (size_t)B * rs * sizeof(cplx), s));
  ensure(w.Gpart, w.G_cap, (size_t)B * gs, s);
  // B tiles of 64 x 64 per panel: split their K range until about one wave of CTAs is in flight
  int ksplit = std::max(1, std::min(std::min(std::min(32, max_split()), rows / 64), (2 * 148) / B));
  int kchunk = ((rows + ksplit - 1) / ksplit + 7) / 8 * 8;
  ksplit = (rows + kchunk - 1) / kchunk;
  cplx* Rinv = w.small; cplx* Rtot = w.small + (size_t)B * gs;
  const double shift_factor = 11.0 * ((double)rows * JP + (double)JP * (JP + 1)) * 2.220446049250313e-16;
  const int npanels = npad / JP;
  const int* ptab = nullptr; const int* ttab = nullptr;
  if (jgemm_enabled() && B <= 65535) {
    const auto key = std::make_pair(-B, npad / JB);
    auto itb = w.tables.find(key);
    if (itb == w.tables.end()) itb = w.tables.emplace(key, build_panel_table(B, npad / JB, s)).first;
    ptab = itb->second;
    const auto key2 = std::make_pair(-B - 100000, npad / JB);
    auto itt = w.tables.find(key2);
    if (itt == w.tables.end()) itt = w.tables.emplace(key2, build_trail_table(B, npad / JB, s)).first;
    ttab = itt->second;
  }
  for (int pk = 0; pk < npanels; ++pk) {
    cplx* P = Q + (long long)pk * JP * ldq;
    for (int it = 0; it < chol_passes; ++it) {
      const int* skip3 = (chol_passes == 3 && it == 2) ? w.cflag : nullptr;     // per problem: third CholeskyQR pass not needed
      if (ptab) {
        const int* tb = ptab + (size_t)pk * B * 2;
        jacobi_gram64(Q, ldq, rows, tb, B, w.Gpart, std::min(32, max_split()), s, skip3);
        launch_chol(B, w.Gpart, Rinv, Rtot, it, it == chol_passes - 1 ? 1 : 0, chol_passes == 3 ? shift_factor : 0.0, chol_passes == 3 ? w.cflag : nullptr, s);
        jacobi_rot64(Q, ldq, rows, tb, B, Rinv, skip3, s);
        continue;
      }
      GemmDesc g = gd(JP, JP, rows, P, idx1(ldq), idx1(1), 1, P, idx1(1), idx1(ldq), 0, w.Gpart, idx1(1), idx1(JP));
      g.batch = B; g.bsA = qs; g.bsB = qs; g.bsC = gs;
      g.ksplit = ksplit; g.kchunk = kchunk; g.ssC = 0; g.atomic_c = ksplit > 1 ? 1 : 0;
      g.skip = skip3;
      if (g.atomic_c) TN_CUDA(cudaMemsetAsync(w.Gpart, 0, (size_t)B * gs * sizeof(cplx), s));
      zgemm_auto(g, s);
      launch_chol(B, w.Gpart, Rinv, Rtot, it, it == chol_passes - 1 ? 1 : 0, chol_passes == 3 ? shift_factor : 0.0, chol_passes == 3 ? w.cflag : nullptr, s);
      GemmDesc ap = gd(rows, JP, JP, P, idx1(1), idx1(ldq), 0, Rinv, idx1(1), idx1(JP), 0, P, idx1(1), idx1(ldq));
      ap.batch = B; ap.bsA = qs; ap.bsB = gs; ap.bsC = qs;
      ap.skip = skip3;
      zgemm_auto(ap, s);
    }
    for (int b = 0; b < B; ++b)     // R_b(pk,pk) = Rtot_b
      TN_CUDA(cudaMemcpy2DAsync(R + b * rs + (long long)pk * JP + (long long)pk * JP * npad, (size_t)npad * sizeof(cplx), Rtot + b * gs,
                                (size_t)JP * sizeof(cplx), (size_t)JP * sizeof(cplx), JP, cudaMemcpyDeviceToDevice, s));
    const int nt = npad - (pk + 1) * JP;
    if (nt > 0) {
      cplx* T = Q + (long long)(pk + 1) * JP * ldq;
      int csplit = std::max(1, std::min(std::min(std::min(16, max_split()), rows / 64), 148 / (B * ((nt + 127) / 128))));
      int cchunk = ((rows + csplit - 1) / csplit + 7) / 8 * 8;
      csplit = (rows + cchunk - 1) / cchunk;
      cplx* Rrow = R + (long long)pk * JP + (long long)(pk + 1) * JP * npad;
      if (ttab && cross_enabled() && (long long)B * (nt / JP) <= 65535) {
        jacobi_cross64(Q, ldq, ptab + (size_t)pk * B * 2, Q, ldq, rows, ttab + trail_offset(B, npanels, pk), B, nt / JP, Rrow, npad, rs, std::min(16, max_split()), s);
      } else {
        GemmDesc c = gd(JP, nt, rows, P, idx1(ldq), idx1(1), 1, T, idx1(1), idx1(ldq), 0, Rrow, idx1(1), idx1(npad));
        c.batch = B; c.bsA = qs; c.bsB = qs; c.bsC = rs;
        c.ksplit = csplit; c.kchunk = cchunk; c.ssC = 0; c.atomic_c = 1;
        zgemm_auto(c, s);
      }
      if (ttab && (long long)B * (nt / JP) <= 65535) {
        jacobi_update64(Q, ldq, ptab + (size_t)pk * B * 2, Q, ldq, rows, ttab + trail_offset(B, npanels, pk), B, nt / JP, Rrow, npad, rs, s);
      } else {
        GemmDesc u = gd(rows, nt, JP, P, idx1(1), idx1(ldq), 0, Rrow, idx1(1), idx1(npad), 0, T, idx1(1), idx1(ldq), -1.0, 1.0);
        u.batch = B; u.bsA = qs; u.bsB = rs; u.bsC = qs;
        zgemm_auto(u, s);
      }
    }
  }
}

static void bgs_qr_b(SvdBatch& w, cplx* Q, long long ldq, int rows, int npad, cudaStream_t s) {
  const size_t rs = (size_t)npad * npad;
  ensure(w.Ra, w.Ra_cap, w.B * rs, s);
  ensure(w.Rb, w.Rb_cap, w.B * rs, s);
  ensure(w.Rc, w.Rc_cap, w.B * rs, s);
  bgs_pass_b(w, Q, ldq, rows, npad, w.Rc, 3, s);
  bgs_pass_b(w, Q, ldq, rows, npad, w.Rb, 1, s);
  GemmDesc g = gd(npad, npad, npad, w.Rb, idx1(1), idx1(npad), 0, w.Rc, idx1(1), idx1(npad), 0, w.Ra, idx1(1), idx1(npad));   // R = R'' R'
  g.batch = w.B; g.bsA = (long long)rs; g.bsB = (long long)rs; g.bsC = (long long)rs;
  zgemm_auto(g, s);
}

void svd_batched_factor(SvdBatch& w, int B, const cplx* const* Ms, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso, bool qr_only) {
  TN_CHECK(B >= 1 && B <= 1024 && m >= 1 && n >= 1, "batched svd: bad batch / shape");
  w.B = B; w.m = m; w.n = n;
  w.transposed = m < n || (m == n && iso == 2);
  w.rows = w.transposed ? n : m;
  w.ncols = w.transposed ? m : n;
  w.npad = ((w.ncols + JP - 1) / JP) * JP;
  const int npad = w.npad, rows = w.rows;
  TN_CHECK(npad <= 8192 && (long long)B * (npad / JB) < (1 << 24), "batched svd: problem too large");
  w.qr_mode = 0;
  if (qr_only) {
    // gauge moves that cannot truncate: T_b = Q1_b R_b for every problem, nothing else (see SvdWork::qr_mode)
    TN_CHECK((iso == 1 && !w.transposed) || (iso == 2 && w.transposed), "batched QR: the isometry must sit on the long side");
    w.qr_mode = 1; w.precond = false; w.wonly = false; w.sweeps = 0;
    ensure(w.cflag, w.cflag_cap, (size_t)B, s);
    ensure(w.small, w.small_cap, (size_t)2 * B * JP * JP, s);
    ensure(w.Q1, w.Q1_cap, (size_t)rows * B * npad, s);
    int blocks;
    launch_1d((long long)rows * npad, blocks);
    for (int b = 0; b < B; ++b)
      svd_init_kernel<<<blocks, 256, 0, s>>>(Ms[b], ld, m, n, w.transposed ? 1 : 0, w.Q1 + (size_t)b * npad * rows, rows, w.ncols, npad, rows);
    TN_CUDA(cudaGetLastError());
    count_launch(B);
    bgs_qr_b(w, w.Q1, rows, rows, npad, s);
    w.k.assign(B, std::min(m, n));
    TN_CUDA(cudaStreamSynchronize(s));
    return;
  }
  w.precond = precond_enabled() && npad > JP;
  w.wonly = w.precond && wonly_enabled() && ((iso == 1 && !w.transposed) || (iso == 2 && w.transposed));
  const size_t tot = (size_t)B * npad;
  if (w.s_cap < tot) {
    if (w.sig) { TN_CUDA(cudaFreeAsync(w.sig, s)); TN_CUDA(cudaFreeAsync(w.perm, s)); TN_CUDA(cudaFreeAsync(w.sig2, s)); }
    TN_CUDA(cudaMallocAsync((void**)&w.sig, tot * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.sig2, tot * sizeof(double), s));
    TN_CUDA(cudaMallocAsync((void**)&w.perm, tot * sizeof(int), s));
    w.s_cap = tot;
  }
  if (!w.offmax) TN_CUDA(cudaMalloc((void**)&w.offmax, 8));
  ensure(w.kout, w.kout_cap, (size_t)B, s);
  ensure(w.cflag, w.cflag_cap, (size_t)B, s);
  ensure(w.small, w.small_cap, (size_t)2 * B * JP * JP, s);
  int blocks;
  if (!w.precond) {
    w.jrows = rows;
    w.ldz = pad_ld(rows + npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * tot, s);
    launch_1d((long long)w.ldz * npad, blocks);
    for (int b = 0; b < B; ++b)
      svd_init_kernel<<<blocks, 256, 0, s>>>(Ms[b], ld, m, n, w.transposed ? 1 : 0, w.Z + (size_t)b * npad * w.ldz, rows, w.ncols, npad, w.ldz);
    TN_CUDA(cudaGetLastError());
    count_launch(B);
    jacobi_sweeps_b(w, rows, s);
  } else {
    ensure(w.Q1, w.Q1_cap, (size_t)rows * tot, s);
    launch_1d((long long)rows * npad, blocks);
    for (int b = 0; b < B; ++b)
      svd_init_kernel<<<blocks, 256, 0, s>>>(Ms[b], ld, m, n, w.transposed ? 1 : 0, w.Q1 + (size_t)b * npad * rows, rows, w.ncols, npad, rows);
    count_launch(B);
    bgs_qr_b(w, w.Q1, rows, rows, npad, s);                       // Ra_b = R1 of problem b
    if (w.wonly) {
      ensure(w.R1, w.R1_cap, (size_t)npad * tot, s);
      TN_CUDA(cudaMemcpyAsync(w.R1, w.Ra, (size_t)npad * tot * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
    }
    ensure(w.Q2, w.Q2_cap, (size_t)npad * tot, s);
    launch_1d((long long)npad * npad, blocks);
    for (int b = 0; b < B; ++b)
      conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra + (size_t)b * npad * npad, npad, npad, npad, w.Q2 + (size_t)b * npad * npad, npad);
    count_launch(B);
    bgs_qr_b(w, w.Q2, npad, npad, npad, s);                       // Ra_b = R2
    w.jrows = npad;
    w.ldz = pad_ld(w.wonly ? npad : 2 * npad);
    ensure(w.Z, w.Z_cap, (size_t)w.ldz * tot, s);
    for (int b = 0; b < B; ++b) {
      cplx* Zb = w.Z + (size_t)b * npad * w.ldz;
      conj_transpose_kernel<<<blocks, 256, 0, s>>>(w.Ra + (size_t)b * npad * npad, npad, npad, npad, Zb, w.ldz);   // W <- R2^H
      if (!w.wonly) set_identity_kernel<<<blocks, 256, 0, s>>>(Zb + npad, w.ldz, npad);
    }
    TN_CUDA(cudaGetLastError());
    count_launch(2 * B);
    jacobi_sweeps_b(w, npad, s);
  }
  colnorm2_kernel<<<(unsigned)tot, 128, 0, s>>>(w.Z, w.jrows, w.ldz, w.sig2);
  int npow2 = 64; while (npow2 < npad) npow2 <<= 1;
  TN_CUDA(cudaFuncSetAttribute(sort_trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
  sort_trunc_kernel<<<B, 1024, npow2 * 12, s>>>(w.sig2, npad, npow2, w.sig, w.perm, w.ncols, tr.cutoff, tr.maxdim, tr.mindim, w.kout);
  TN_CUDA(cudaGetLastError());
  count_launch(2);
  w.k.assign(B, 0);
  TN_CUDA(cudaMemcpyAsync(w.k.data(), w.kout, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, s));
  TN_CUDA(cudaStreamSynchronize(s));
}

// single-problem view of problem b: the ordinary gathers work on it unchanged
static SvdWork batch_view(SvdBatch& w, int b) {
  TN_CHECK(b >= 0 && b < w.B, "batched svd: problem index out of range");
  SvdWork v;
  const size_t npad = (size_t)w.npad;
  v.Z = w.Z ? w.Z + (size_t)b * npad * w.ldz : nullptr;
  v.sig = w.sig ? w.sig + b * npad : nullptr; v.perm = w.perm ? w.perm + b * npad : nullptr;
  v.Q1 = w.Q1 ? w.Q1 + (size_t)b * npad * w.rows : nullptr;
  v.Q2 = w.Q2 ? w.Q2 + (size_t)b * npad * npad : nullptr;
  v.Tg = w.Tg; v.Tg_cap = w.Tg_cap;
  v.precond = w.precond; v.jrows = w.jrows;
  v.m = w.m; v.n = w.n; v.rows = w.rows; v.ncols = w.ncols; v.ncols_pad = w.npad; v.ldz = w.ldz; v.nsv = w.ncols; v.k = w.k[b];
  v.transposed = w.transposed;
  v.wonly = w.wonly; v.qr_mode = w.qr_mode;
  v.R1 = (w.wonly && w.R1) ? w.R1 + (size_t)b * npad * npad : nullptr;
  v.Ra = w.Ra ? w.Ra + (size_t)b * npad * npad : nullptr;
  return v;
}
void svd_batched_gather_U(SvdBatch& w, int b, cplx* U, long long ldu, bool times_S, cudaStream_t s) {
  SvdWork v = batch_view(w, b);
  svd_gather_U(v, U, ldu, times_S, s);
  w.Tg = v.Tg; w.Tg_cap = v.Tg_cap;
}
void svd_batched_gather_Vh(SvdBatch& w, int b, cplx* Vh, long long ldv, bool times_S, cudaStream_t s) {
  SvdWork v = batch_view(w, b);
  svd_gather_Vh(v, Vh, ldv, times_S, s);
  w.Tg = v.Tg; w.Tg_cap = v.Tg_cap;
}
void svd_batched_copy_S(SvdBatch& w, int b, double* S, cudaStream_t s) {
  TN_CUDA(cudaMemcpyAsync(S, w.sig + (size_t)b * w.npad, (size_t)w.k[b] * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void svd_batched_free(SvdBatch& w) {
  for (cplx* p : {w.Z, w.Q1, w.Q2, w.Ra, w.Rb, w.Rc, w.Gpart, w.J, w.small, w.Tg, w.R1}) if (p) cudaFree(p);
  if (w.skip) cudaFree(w.skip);
  if (w.cflag) cudaFree(w.cflag);
  if (w.kout) cudaFree(w.kout);
  if (w.sig) { cudaFree(w.sig); cudaFree(w.sig2); cudaFree(w.perm); }
  if (w.offmax) cudaFree(w.offmax);
  for (auto& kv : w.tables) cudaFree(kv.second);
  w = SvdBatch{};
}

// ================================================================================================
// Batching rounds across worker threads (see tn_svd.cuh)
// ================================================================================================
struct BatchReq { SvdWork* w; const cplx* M; int m, n; long long ld; Trunc tr; int k; int iso; bool qr; };
struct SvdBatcher {
  Rounds<BatchReq> rounds;
  typedef std::tuple<int, int, long long, double, long long, long long, int> Key;   // m, n, ld, cutoff, maxdim, mindim, iso + 4 * qr
  std::map<Key, SvdBatch> work;                                               // one stacked workspace per shape group
  explicit SvdBatcher(int n) : rounds(n) {}
};

SvdBatcher* svd_batcher_create(int nworkers) { return new SvdBatcher(nworkers); }
void svd_batcher_destroy(SvdBatcher* b) {
  if (!b) return;
  for (auto& kv : b->work) svd_batched_free(kv.second);
  delete b;
}
void svd_batcher_attach(SvdBatcher* b) { tl_batcher = b; }
long long svd_batcher_rounds(SvdBatcher* b, long long* problems) { if (problems) *problems = b->rounds.requests(); return (long long)b->rounds.rounds(); }

// One round (called by Rounds with its lock held, by the thread that completes the round): factorises every parked request on stream s.
static void batcher_run_round(SvdBatcher& bt, std::vector<BatchReq*>& pending, cudaStream_t s) {
  // data-dependent bond dimensions produce many shapes over a long run: drop the cached workspaces now and then (safe here:
  // every worker synchronised its stream before parking, so nobody still reads the previous round's factors)
  if (bt.work.size() > 48) { for (auto& kv : bt.work) svd_batched_free(kv.second); bt.work.clear(); }
  std::map<SvdBatcher::Key, std::vector<BatchReq*>> groups;
  for (BatchReq* r : pending) groups[SvdBatcher::Key(r->m, r->n, r->ld, r->tr.cutoff, r->tr.maxdim, r->tr.mindim, r->iso + (r->qr ? 4 : 0))].push_back(r);
  for (auto& kv : groups) {
    std::vector<BatchReq*>& g = kv.second;
    SvdBatch& wb = bt.work[kv.first];
    std::vector<const cplx*> ptrs;
    for (BatchReq* r : g) ptrs.push_back(r->M);
    svd_batched_factor(wb, (int)g.size(), ptrs.data(), g[0]->m, g[0]->n, g[0]->ld, g[0]->tr, s, g[0]->iso, g[0]->qr);   // ends with a stream synchronise
    for (size_t b = 0; b < g.size(); ++b) {
      SvdWork& w = *g[b]->w;
      if (!w.bview) w.bview = new SvdWork();
      *w.bview = batch_view(wb, (int)b);
      w.use_view = true;
      w.k = wb.k[b]; w.m = wb.m; w.n = wb.n; w.sweeps = wb.sweeps; w.qr_mode = 0;      // (the view carries the mode of the batch)
      g[b]->k = wb.k[b];
    }
  }
}

static int batcher_submit(SvdWork& w, const cplx* M, int m, int n, long long ld, Trunc tr, cudaStream_t s, int iso, bool qr) {
  SvdBatcher& bt = *tl_batcher;
  // the input matrix is complete, and every gather this thread enqueued from the previous round's workspace has finished,
  // before the workspace can be overwritten by the next round
  TN_CUDA(cudaStreamSynchronize(s));
  BatchReq rq{&w, M, m, n, ld, tr, 0, iso, qr};
  try { bt.rounds.submit(&rq, [&](std::vector<BatchReq*>& pending) { batcher_run_round(bt, pending, s); }); }
  catch (const std::runtime_error& e) { throw tn::Error(-3, std::string("svd batching round: ") + e.what()); }
  return rq.k;
}

void svd_batcher_detach(cudaStream_t s) {
  SvdBatcher* b = tl_batcher;
  if (!b) return;
  tl_batcher = nullptr;
  b->rounds.leave([&](std::vector<BatchReq*>& pending) { batcher_run_round(*b, pending, s); });
}

}  // namespace tn
