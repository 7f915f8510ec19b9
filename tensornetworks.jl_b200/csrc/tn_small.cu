// Small bond dimensions: the whole eigensolver of one DMRG bond in ONE single-CTA launch.
//
// At the reference's own example (examples/dmrg.jl: TFIM N = 100, maxdim 32, bonds <= 11) the two-site tensor has 484 elements and a
// bond of the general path is ~60 launches of near-empty GEMM / dot / axpy kernels plus two host round trips for the Lanczos
// coefficients (profiles/r01_c1_sweep_launches_summary.csv: the GPU lost to 16 host threads there).  Here everything the bond's
// eigsolve needs lives in shared memory -- L, R, W = M1.M2, the three Lanczos vectors, the residual, the two intermediates of the
// H_eff application -- and one CTA runs
//     Theta0 = A1.A2                                       (dmrg.jl:44-47)
//     KrylovKit's schedule eigsolve(Heff, Theta0, 1, :SR; krylovdim = 3, maxiter = 2, tol)   (dmrg.jl:51-53: 3 + 2 applications,
//       thick restart keeping one Ritz vector, full re-orthogonalisation -- the same steps as lanczos_core in tn_mps.cu)
//     H_eff.Theta in the flop-optimal order (L.Theta).W.R  (projmps.jl:107-134)
// with plain FP64 FMAs (a 484-dimensional problem has no use for the tensor pipe).  The host reads nothing back: the energy and the
// number of H_eff applications stay in device memory until the end of the half sweep.
#include "tn_mps.cuh"
#include <cmath>

namespace tn {
void count_launch(int n);

constexpr int SL_THREADS = 512;

struct SmallBond {
  int ca, cm, ca2, d, w, w1, w2;       // Theta (ca, d, d, ca2); A1 (ca, d, cm); A2 (cm, d, ca2); L (ca, w, ca); R (ca2, w2, ca2)
  const cplx *A1, *A2, *L, *R, *M1, *M2;
  cplx coeff;
  cplx* theta_out;                     // normalised Ritz vector (n)
  double* energy;                      // lowest Ritz value
  int* numops;                         // += H_eff applications of this bond
  int krylovdim, maxiter; double tol;
};

__device__ __forceinline__ cplx cmul_(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// cyclic Jacobi on a real symmetric K x K (K <= 3) matrix; eigenvalues ascending (same routine as eigh_sym3 in tn_mps.cu)
__device__ void eigh_sym3_dev(int K, const double (*T)[3], double* D, double (*U)[3]) {
  double A[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = (i < K && j < K) ? T[i][j] : 0.0; U[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int i = 0; i < K; ++i) for (int j = i + 1; j < K; ++j) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < K; ++p)
      for (int q = p + 1; q < K; ++q) {
        if (A[p][q] == 0.0) continue;
        double tau = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
        double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
        for (int k = 0; k < K; ++k) { double x = A[k][p], y = A[k][q]; A[k][p] = cs * x - sn * y; A[k][q] = sn * x + cs * y; }
        for (int k = 0; k < K; ++k) { double x = A[p][k], y = A[q][k]; A[p][k] = cs * x - sn * y; A[q][k] = sn * x + cs * y; }
        for (int k = 0; k < K; ++k) { double x = U[k][p], y = U[k][q]; U[k][p] = cs * x - sn * y; U[k][q] = sn * x + cs * y; }
      }
  }
  int ord[3] = {0, 1, 2};
  for (int i = 1; i < K; ++i)
    for (int j = i; j > 0 && A[ord[j]][ord[j]] < A[ord[j - 1]][ord[j - 1]]; --j) { int t = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = t; }
  double Us[3][3];
  for (int j = 0; j < K; ++j) { D[j] = A[ord[j]][ord[j]]; for (int i = 0; i < K; ++i) Us[i][j] = U[i][ord[j]]; }
  for (int i = 0; i < K; ++i) for (int j = 0; j < K; ++j) U[i][j] = Us[i][j];
}

struct SmallShared {
  double T[3][3], D[3], U[3][3];
  cplx dots[4];
  double alpha[3], beta2[3];
  double red[SL_THREADS / 32][8];
  int K, numiter, numops, logn, stop, keep, converged;
  double bet;
};

// <x_j, y> for j < nx (conjugating x), result in sh.dots[j]; every thread returns after the values are visible
__device__ void block_dots(SmallShared& sh, int n, int nx, const cplx* x0, const cplx* x1, const cplx* x2, const cplx* y) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = tid; i < n; i += SL_THREADS) {
    const cplx v = y[i];
    const cplx a = x0[i];
    acc[0] = fma(a.x, v.x, fma(a.y, v.y, acc[0])); acc[1] = fma(a.x, v.y, fma(-a.y, v.x, acc[1]));
    if (nx > 1) { const cplx b = x1[i]; acc[2] = fma(b.x, v.x, fma(b.y, v.y, acc[2])); acc[3] = fma(b.x, v.y, fma(-b.y, v.x, acc[3])); }
    if (nx > 2) { const cplx c = x2[i]; acc[4] = fma(c.x, v.x, fma(c.y, v.y, acc[4])); acc[5] = fma(c.x, v.y, fma(-c.y, v.x, acc[5])); }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  if (lane == 0) { for (int q = 0; q < 6; ++q) sh.red[warp][q] = acc[q]; }
  __syncthreads();
  if (tid < 6) { double s = 0; for (int wv = 0; wv < SL_THREADS / 32; ++wv) s += sh.red[wv][tid]; if (tid & 1) sh.dots[tid >> 1].y = s; else sh.dots[tid >> 1].x = s; }
  __syncthreads();
}

__global__ void __launch_bounds__(SL_THREADS, 1) lanczos_small_kernel(SmallBond p) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ SmallShared sh;
  const int tid = threadIdx.x;
  const int ca = p.ca, ca2 = p.ca2, d = p.d, d2 = d * d, w = p.w, w1 = p.w1, w2 = p.w2, cm = p.cm;
  const int n = ca * d2 * ca2, nL = ca * w * ca, nR = ca2 * w2 * ca2, KW = w * d2, NW = d2 * w2, nT1 = ca * w * d2 * ca2, nT2 = ca * d2 * w2 * ca2;
  cplx* Ls = reinterpret_cast<cplx*>(sm_raw);
  cplx* Rs = Ls + nL;
  cplx* Ws = Rs + nR;
  cplx* V[3]; V[0] = Ws + KW * NW; V[1] = V[0] + n; V[2] = V[1] + n;
  cplx* wv = V[2] + n;
  cplx* tmp0 = wv + n;
  cplx* tmp1 = tmp0 + n;
  cplx* T1 = tmp1 + n;
  cplx* T2 = T1 + nT1;
  for (int e = tid; e < nL; e += SL_THREADS) Ls[e] = p.L[e];
  for (int e = tid; e < nR; e += SL_THREADS) Rs[e] = p.R[e];
  // W[(w,s1',s2'),(s1,s2,w2)] = sum_{w1} M1(w,s1,s1',w1) M2(w1,s2,s2',w2)
  for (int e = tid; e < KW * NW; e += SL_THREADS) {
    const int kk = e % KW, nn = e / KW;
    const int iw = kk % w, s1p = (kk / w) % d, s2p = kk / (w * d);
    const int s1 = nn % d, s2 = (nn / d) % d, iw2 = nn / (d * d);
    double xr = 0, xi = 0;
    for (int j = 0; j < w1; ++j) {
      const cplx a = p.M1[iw + w * (s1 + d * (s1p + d * j))];
      const cplx b = p.M2[j + w1 * (s2 + d * (s2p + d * iw2))];
      xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x;
    }
    Ws[e] = make_double2(xr, xi);
  }
  // Theta0[(a,s1),(s2,a')] = sum_m A1[(a,s1),m] A2[m,(s2,a')]  -> tmp0
  for (int e = tid; e < n; e += SL_THREADS) {
    const int r = e % (ca * d), c = e / (ca * d);
    double xr = 0, xi = 0;
    for (int m = 0; m < cm; ++m) { const cplx a = p.A1[r + ca * d * m], b = p.A2[m + cm * c]; xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x; }
    tmp0[e] = make_double2(xr, xi);
  }
  if (tid == 0) { sh.numops = 0; sh.logn = 0; sh.numiter = 1; sh.stop = 0; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sh.T[i][j] = 0; }
  __syncthreads();

  // out = coeff * H_eff * in   (in, out: n-vectors in shared memory)
  auto apply = [&](const cplx* in, cplx* out) {
    for (int e = tid; e < nT1; e += SL_THREADS) {       // T1[(a,w),(s1',s2',b')] = sum_b L[(a,w),b] in[b,(s1',s2',b')]
      const int r = e % (ca * w), col = e / (ca * w);
      double xr = 0, xi = 0;
      for (int b = 0; b < ca; ++b) { const cplx l = Ls[r + ca * w * b], t = in[b + ca * col]; xr += l.x * t.x - l.y * t.y; xi += l.x * t.y + l.y * t.x; }
      T1[e] = make_double2(xr, xi);
    }
    __syncthreads();
    for (int e = tid; e < nT2; e += SL_THREADS) {       // T2(a,(s1,s2,w2),b') = sum_k T1(a,k,b') W[k,(s1,s2,w2)]
      const int a = e % ca, nn = (e / ca) % NW, bp = e / (ca * NW);
      double xr = 0, xi = 0;
      for (int k = 0; k < KW; ++k) { const cplx t = T1[a + ca * k + ca * KW * bp], x = Ws[k + KW * nn]; xr += t.x * x.x - t.y * x.y; xi += t.x * x.y + t.y * x.x; }
      T2[e] = make_double2(xr, xi);
    }
    __syncthreads();
    for (int e = tid; e < n; e += SL_THREADS) {         // out[(a,s1,s2),a'] = coeff sum_{(w2,b')} T2[(a,s1,s2),(w2,b')] R[a',(w2,b')]
      const int m = e % (ca * d2), ap = e / (ca * d2);
      const cplx* t2 = T2 + (m % ca) + ca * (m / ca);   // + ca d2 iw2 + ca NW b'
      const cplx* rr = Rs + ap;                         // + ca2 (iw2 + w2 b')
      double xr = 0, xi = 0;
      for (int bp = 0; bp < ca2; ++bp)
        for (int iw2 = 0; iw2 < w2; ++iw2) {
          const cplx t = t2[ca * d2 * iw2 + ca * NW * bp], r = rr[ca2 * (iw2 + w2 * bp)];
          xr += t.x * r.x - t.y * r.y; xi += t.x * r.y + t.y * r.x;
        }
      out[e] = cmul_(p.coeff, make_double2(xr, xi));
    }
    __syncthreads();
    if (tid == 0) sh.numops++;
  };
  // w = H V[Kc-1]; alpha = Re<v,w>; w -= sum_j <v_j,w> v_j (twice); beta^2 = <w,w>   (recorded at slot sh.logn)
  auto expand = [&](int Kc) {
    apply(V[Kc - 1], wv);
    block_dots(sh, n, Kc, V[0], V[1], V[2], wv);
    if (tid == 0) sh.alpha[sh.logn] = sh.dots[Kc - 1].x;
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1) block_dots(sh, n, Kc, V[0], V[1], V[2], wv);
      const cplx h0 = sh.dots[0], h1 = sh.dots[1], h2 = sh.dots[2];
      for (int i = tid; i < n; i += SL_THREADS) {
        cplx v = wv[i];
        { const cplx t = cmul_(h0, V[0][i]); v.x -= t.x; v.y -= t.y; }
        if (Kc > 1) { const cplx t = cmul_(h1, V[1][i]); v.x -= t.x; v.y -= t.y; }
        if (Kc > 2) { const cplx t = cmul_(h2, V[2][i]); v.x -= t.x; v.y -= t.y; }
        wv[i] = v;
      }
      __syncthreads();
    }
    block_dots(sh, n, 1, wv, wv, wv, wv);
    if (tid == 0) { sh.beta2[sh.logn] = sh.dots[0].x; sh.logn++; }
    __syncthreads();
  };
  auto scale_into = [&](const cplx* x, double nrm2, cplx* out) {      // out = x / sqrt(nrm2)
    const double inv = nrm2 > 0.0 ? 1.0 / sqrt(nrm2) : 0.0;      // invariant subspace (beta == 0): zeros, never NaN (as zscale_invnorm)
    for (int i = tid; i < n; i += SL_THREADS) out[i] = make_double2(x[i].x * inv, x[i].y * inv);
    __syncthreads();
  };

  const int KD = p.krylovdim;
  // v1 = theta0 / ||theta0||
  block_dots(sh, n, 1, tmp0, tmp0, tmp0, tmp0);
  scale_into(tmp0, sh.dots[0].x, V[0]);
  // ---- round 1
  expand(1);
  for (int k = 2; k <= KD; ++k) { scale_into(wv, sh.beta2[sh.logn - 1], V[k - 1]); expand(k); }
  if (tid == 0) {
    double beta[3];
    for (int i = 0; i < KD; ++i) beta[i] = sqrt(fmax(0.0, sh.beta2[i]));
    sh.K = KD;
    for (int i = 0; i < KD; ++i) { sh.T[i][i] = sh.alpha[i]; if (i + 1 < KD) sh.T[i][i + 1] = sh.T[i + 1][i] = beta[i]; }
    sh.bet = beta[KD - 1];
    for (int i = 0; i < KD - 1; ++i) if (beta[i] <= p.tol) { sh.K = i + 1; sh.bet = beta[i]; break; }   // invariant subspace inside round 1
  }
  __syncthreads();
  while (true) {
    if (tid == 0) {
      eigh_sym3_dev(sh.K, sh.T, sh.D, sh.U);
      int conv = 0;
      while (conv < sh.K && fabs(sh.U[sh.K - 1][conv] * sh.bet) <= p.tol) conv++;
      sh.converged = conv;
      int keep = (3 * KD + 2 * conv) / 5;
      if (keep >= KD) keep = KD - 1;
      sh.keep = keep;
      sh.stop = (conv >= 1 || sh.bet <= p.tol || sh.K < KD || sh.numiter == p.maxiter || keep < 1) ? 1 : 0;
    }
    __syncthreads();
    if (sh.stop) break;
    // ---- thick restart: keep Ritz vectors + the residual direction
    const int keep = sh.keep, K = sh.K;
    for (int j = 0; j < keep; ++j) {
      cplx* dst = j == 0 ? tmp0 : tmp1;
      const double c0 = sh.U[0][j], c1 = sh.U[1][j], c2 = sh.U[2][j];
      for (int i = tid; i < n; i += SL_THREADS) {
        double xr = c0 * V[0][i].x, xi = c0 * V[0][i].y;
        if (K > 1) { xr += c1 * V[1][i].x; xi += c1 * V[1][i].y; }
        if (K > 2) { xr += c2 * V[2][i].x; xi += c2 * V[2][i].y; }
        dst[i] = make_double2(xr, xi);
      }
    }
    __syncthreads();
    for (int j = 0; j < keep; ++j) { const cplx* src = j == 0 ? tmp0 : tmp1; for (int i = tid; i < n; i += SL_THREADS) V[j][i] = src[i]; }
    __syncthreads();
    scale_into(wv, sh.beta2[sh.logn - 1], V[keep]);
    if (tid == 0) {
      double f[3];
      for (int j = 0; j < keep; ++j) f[j] = sh.U[K - 1][j] * sh.bet;
      double D0[3] = {sh.D[0], sh.D[1], sh.D[2]};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sh.T[i][j] = 0;
      for (int j = 0; j < keep; ++j) { sh.T[j][j] = D0[j]; sh.T[j][keep] = sh.T[keep][j] = f[j]; }
      sh.logn = 0;
    }
    __syncthreads();
    const int first = keep + 1;
    expand(first);
    for (int k = first + 1; k <= KD; ++k) { scale_into(wv, sh.beta2[sh.logn - 1], V[k - 1]); expand(k); }
    if (tid == 0) {
      const int cnt = KD - keep;
      double beta[3];
      for (int i = 0; i < cnt; ++i) beta[i] = sqrt(fmax(0.0, sh.beta2[i]));
      for (int i = 0; i < cnt; ++i) {
        const int r = keep + i;
        sh.T[r][r] = sh.alpha[i];
        if (r + 1 < KD) sh.T[r][r + 1] = sh.T[r + 1][r] = beta[i];
      }
      sh.bet = beta[cnt - 1];
      sh.K = KD;
      sh.numiter++;
    }
    __syncthreads();
  }
  // Ritz vector of the lowest Ritz value, renormalised
  {
    const int K = sh.K;
    const double c0 = sh.U[0][0], c1 = sh.U[1][0], c2 = sh.U[2][0];
    for (int i = tid; i < n; i += SL_THREADS) {
      double xr = c0 * V[0][i].x, xi = c0 * V[0][i].y;
      if (K > 1) { xr += c1 * V[1][i].x; xi += c1 * V[1][i].y; }
      if (K > 2) { xr += c2 * V[2][i].x; xi += c2 * V[2][i].y; }
      wv[i] = make_double2(xr, xi);
    }
    __syncthreads();
    block_dots(sh, n, 1, wv, wv, wv, wv);
    const double inv = 1.0 / sqrt(sh.dots[0].x);
    for (int i = tid; i < n; i += SL_THREADS) p.theta_out[i] = make_double2(wv[i].x * inv, wv[i].y * inv);
    if (tid == 0) { *p.energy = sh.D[0]; atomicAdd(p.numops, sh.numops); }
  }
}

size_t small_bond_smem(int ca, int ca2, int d, int w, int w2) {
  const size_t d2 = (size_t)d * d, n = (size_t)ca * d2 * ca2;
  return ((size_t)ca * w * ca + (size_t)ca2 * w2 * ca2 + (size_t)w * d2 * d2 * w2 + 6 * n + n * w + n * w2) * sizeof(cplx);
}

// Returns false when the bond does not fit (the caller then runs the general path).
bool lanczos_small(Env* e, int site, double* energy_dev, int* numops_dev, cplx* theta_out, Lanczos lz) {
  static int enabled = -1;
  if (enabled < 0) { const char* ev = getenv("TN_SMALL_BOND"); enabled = (ev && ev[0] == '0') ? 0 : 1; }
  if (!enabled || e->mpo == nullptr || e->squared || lz.krylovdim < 1 || lz.krylovdim > 3) return false;
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  const Tensor& L = env_block(e, site - 1);
  const Tensor& R = env_block(e, site + 2);
  const Tensor& M1 = e->mpo->sites[site - 1];
  const Tensor& M2 = e->mpo->sites[site];
  const Tensor& A1 = e->ket->sites[site - 1];
  const Tensor& A2 = e->ket->sites[site];
  SmallBond p;
  p.d = e->ket->d;
  p.ca = (int)L.dims[0]; p.w = (int)L.dims[1]; p.ca2 = (int)R.dims[0]; p.w2 = (int)R.dims[1]; p.w1 = (int)M1.dims[3]; p.cm = (int)A1.dims[2];
  if (L.dims[2] != L.dims[0] || R.dims[2] != R.dims[0] || A1.dims[0] != p.ca || A2.dims[2] != p.ca2) return false;   // bra and ket bonds must agree (the eigenproblem is square)
  TN_CHECK(M1.dims[0] == p.w && M2.dims[0] == p.w1 && M2.dims[3] == p.w2, "product: MPO / block bond mismatch");
  const size_t smem = small_bond_smem(p.ca, p.ca2, p.d, p.w, p.w2);
  if (smem > 200 * 1024) return false;
  static DeviceOnce cfg;
  cfg.run([&] { TN_CUDA(cudaFuncSetAttribute(lanczos_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); });
  p.A1 = A1.p; p.A2 = A2.p; p.L = L.p; p.R = R.p; p.M1 = M1.p; p.M2 = M2.p;
  p.coeff = e->coeff; p.theta_out = theta_out; p.energy = energy_dev; p.numops = numops_dev;
  p.krylovdim = lz.krylovdim; p.maxiter = lz.maxiter; p.tol = lz.tol;
  lanczos_small_kernel<<<1, SL_THREADS, smem, s>>>(p);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
  return true;
}

}  // namespace tn
