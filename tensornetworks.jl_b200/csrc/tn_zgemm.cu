// Complex-FP64 tensor contraction as a strided GEMM on the FP64 tensor pipe (DMMA.8x8x4) for sm_100a.
//
// Replaces every `contract` on the hot path (reference src/tensors.jl:9-18 -> TensorOperations TTGT:
// permute copy + zgemm + permute copy).  Here the permutes are folded into the tile loads: each operand
// index (m, n, k) may be a fused pair of tensor indices with independent strides (Idx2), addressed
// directly by the 16-byte cp.async that stages the tile, so no transpose pass ever touches HBM.
//
// Arithmetic: complex MAC = 4 real DMMA.8x8x4 (re += ar*br, re += ai*(-bi), im += ar*bi, im += ai*br).
// Complex elements stay interleaved in shared memory; one LDS.128 per fragment element yields the re and
// im fragment registers at once, so the "de-interleave" costs nothing.
//
// Tile: CTA = BM x BN complex outputs, 8 warps (WARPS_M x WARPS_N), warp tile (8*TM) x (8*TN),
// BK = 8 per stage, 4-stage cp.async pipeline.  Shared layout is [k][m] / [k][n] with the leading
// dimension padded to == 2 (mod 8) 16-byte units so the 8 lanes of a quarter-warp hit 8 distinct
// bank groups on fragment loads.
#include "tn_common.cuh"
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdlib>

namespace tn {

static std::atomic<long long> g_launches{0};
long long zgemm_launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Sign flips of fragment registers as integer operations on the sign bit: a DMUL / DADD would occupy the FP64 pipe the DMMAs run on
// (12 of them per 64 DMMA in the k loop).  flip_hi = 0 or 0x80000000 (applied to the high word).
__device__ __forceinline__ double sign_xor(double x, unsigned flip_hi) { return __hiloint2double(__double2hiint(x) ^ (int)flip_hi, __double2loint(x)); }

__device__ __forceinline__ long long idx_off(const Idx2& ix, int i, int batch) {
  if (i < ix.n0) return (long long)i * ix.s0;
  int i1 = i / ix.n0, i0 = i - i1 * ix.n0;
  if (ix.tab) i1 = ix.tab[(long long)ix.tab_bs * batch + i1];
  return (long long)i0 * ix.s0 + (long long)i1 * ix.s1;
}
// table lookups must also apply when i < n0 (i1 == 0)
__device__ __forceinline__ long long idx_off_t(const Idx2& ix, int i, int batch) {
  if (!ix.tab) return idx_off(ix, i, batch);
  int i1 = i / ix.n0, i0 = i - i1 * ix.n0;
  i1 = ix.tab[(long long)ix.tab_bs * batch + i1];
  return (long long)i0 * ix.s0 + (long long)i1 * ix.s1;
}

template <int WARPS_M, int WARPS_N, int TM, int TN>
struct TileCfg {
  static constexpr int BM = WARPS_M * TM * 8;
  static constexpr int BN = WARPS_N * TN * 8;
  static constexpr int BK = 8;
  static constexpr int STAGES = 4;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int LDA = BM + 2;
  static constexpr int LDB = BN + 2;
  static constexpr int A_STAGE = BK * LDA;   // complex elements
  static constexpr int B_STAGE = BK * LDB;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * 16;
  static constexpr int A_PER_THR = BM * BK / THREADS;
  static constexpr int B_PER_THR = BN * BK / THREADS;
};

// resident CTAs per SM the register allocation must allow: 4 / 2 for 2- / 4-warp CTAs, 2 for 8 warps with small warp tiles
template <int WARPS_M, int WARPS_N, int TM, int TN>
struct MinBlocks { static constexpr int value = (WARPS_M * WARPS_N <= 2) ? 4 : ((WARPS_M * WARPS_N <= 4) ? 2 : ((TM * TN <= 8) ? 2 : 1)); };

// KMODE selects how the load slots walk the k index:
//   1 (simple)  both operands' k index is single-level without a lookup table: each slot advances a pointer by
//               BK * stride per k-tile (no index arithmetic inside the pipeline);
//   2 (aligned) two-level / table k index whose level-1 blocks are multiples of BK (Jacobi column-block pairs): the same
//               pointer walk, plus one shared pointer correction per operand when a k-tile enters the next block;
//   0 (general) any two-level index (e.g. k = (w, s') with w = 5): per-slot (k0, k1) bookkeeping.
// One CTA tile: C[m_blk.., n_blk..] (+)= alpha * sum_{k in [k_begin, k_end)} op(A) op(B).
// ATOMIC = false: C = alpha*acc + beta*C (plain stores); ATOMIC = true: C += alpha*acc with red.global.add.f64
// (stream-K partial tiles; C was zeroed by the launcher).
template <int WARPS_M, int WARPS_N, int TM, int TN, int KMODE>
__device__ __forceinline__ void gemm_tile(const GemmDesc& d, cplx* As, cplx* Bs, const int m_blk, const int n_blk, const int batch,
                                          const int split, const int k_begin, const int k_end, const bool atomic) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, TM, TN>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int LDA = Cfg::LDA, LDB = Cfg::LDB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int g = lane >> 2, t = lane & 3;
  const cplx* __restrict__ Ab = d.A + (long long)batch * d.bsA;
  const cplx* __restrict__ Bb = d.B + (long long)batch * d.bsB;

  // ---- per-thread global->shared load assignment (fixed m/n, walking k) ------------------------
  // Each slot keeps its running k position as (k0, k1) = (k mod n0, k div n0) so that advancing by BK
  // needs no integer division inside the pipeline.
  // (base = offset of the current level-1 block, looked up / multiplied only when k1 changes: the table lives in global
  // memory and a lookup in front of every cp.async would put an L2 round trip into the load-issue path)
  struct KPos { int k, k0, k1; long long base; };
  auto kpos_base = [&](const Idx2& ix, int k1) -> long long {
    int l1 = k1;
    if (ix.tab) l1 = ix.tab[(long long)ix.tab_bs * batch + k1];
    return (long long)l1 * ix.s1;
  };
  auto kpos_init = [&](const Idx2& ix, int k, int kend) {
    KPos p; p.k = k; p.k1 = k / ix.n0; p.k0 = k - p.k1 * ix.n0; p.base = (KMODE == 0 && k < kend) ? kpos_base(ix, p.k1) : 0; return p;
  };
  auto kpos_off = [&](const Idx2& ix, const KPos& p) -> long long { return (long long)p.k0 * ix.s0 + p.base; };
  auto kpos_adv = [&](const Idx2& ix, KPos& p, int kend) {
    p.k += BK; p.k0 += BK;
    if (p.k0 >= ix.n0) {
      while (p.k0 >= ix.n0) { p.k0 -= ix.n0; p.k1++; }
      if (p.k < kend) p.base = kpos_base(ix, p.k1);
    }
  };
  int a_m[Cfg::A_PER_THR], a_k[Cfg::A_PER_THR];
  const cplx* a_base[Cfg::A_PER_THR]; bool a_ok[Cfg::A_PER_THR]; KPos a_pos[Cfg::A_PER_THR];
#pragma unroll
  for (int i = 0; i < Cfg::A_PER_THR; ++i) {
    int e = tid + i * THREADS;
    if (d.a_kfast) { a_k[i] = e % BK; a_m[i] = e / BK; } else { a_m[i] = e % BM; a_k[i] = e / BM; }
    int m = m_blk + a_m[i];
    a_ok[i] = m < d.M;
    a_base[i] = Ab + (a_ok[i] ? idx_off_t(d.am, m, batch) : 0);
    a_pos[i] = kpos_init(d.ak, k_begin + a_k[i], k_end);
    if (KMODE == 1) a_base[i] += (long long)(k_begin + a_k[i]) * d.ak.s0;
  }
  int b_n[Cfg::B_PER_THR], b_k[Cfg::B_PER_THR];
  const cplx* b_base[Cfg::B_PER_THR]; bool b_ok[Cfg::B_PER_THR]; KPos b_pos[Cfg::B_PER_THR];
#pragma unroll
  for (int i = 0; i < Cfg::B_PER_THR; ++i) {
    int e = tid + i * THREADS;
    if (d.b_kfast) { b_k[i] = e % BK; b_n[i] = e / BK; } else { b_n[i] = e % BN; b_k[i] = e / BN; }
    int n = n_blk + b_n[i];
    b_ok[i] = n < d.N;
    b_base[i] = Bb + (b_ok[i] ? idx_off_t(d.bn, n, batch) : 0);
    b_pos[i] = kpos_init(d.bk, k_begin + b_k[i], k_end);
    if (KMODE == 1) b_base[i] += (long long)(k_begin + b_k[i]) * d.bk.s0;
  }
  // KMODE 2: per-operand block state (k of the next tile, elements left in the current level-1 block, its index / offset)
  int a_kt = k_begin, b_kt = k_begin, a_left = 0, b_left = 0, a_k1 = 0, b_k1 = 0;
  long long a_cur = 0, b_cur = 0;
  if (KMODE == 2) {
    a_k1 = k_begin / d.ak.n0; b_k1 = k_begin / d.bk.n0;
    const int a_k0 = k_begin - a_k1 * d.ak.n0, b_k0 = k_begin - b_k1 * d.bk.n0;
    a_left = d.ak.n0 - a_k0; b_left = d.bk.n0 - b_k0;
    if (k_begin < k_end) { a_cur = kpos_base(d.ak, a_k1); b_cur = kpos_base(d.bk, b_k1); }
#pragma unroll
    for (int i = 0; i < Cfg::A_PER_THR; ++i) a_base[i] += (long long)(a_k0 + a_k[i]) * d.ak.s0 + a_cur;
#pragma unroll
    for (int i = 0; i < Cfg::B_PER_THR; ++i) b_base[i] += (long long)(b_k0 + b_k[i]) * d.bk.s0 + b_cur;
  }

  // loads the next k-tile (tiles are always requested in increasing order)
  const long long a_step = (long long)BK * d.ak.s0, b_step = (long long)BK * d.bk.s0;
  auto load_stage = [&](int stage) {
    cplx* as = As + stage * Cfg::A_STAGE;
    cplx* bs = Bs + stage * Cfg::B_STAGE;
    if (KMODE == 1) {
#pragma unroll
      for (int i = 0; i < Cfg::A_PER_THR; ++i) {
        bool p = a_ok[i] && (a_pos[i].k < k_end);
        cp_async16(as + a_k[i] * LDA + a_m[i], p ? a_base[i] : Ab, p);
        a_base[i] += a_step; a_pos[i].k += BK;
      }
#pragma unroll
      for (int i = 0; i < Cfg::B_PER_THR; ++i) {
        bool p = b_ok[i] && (b_pos[i].k < k_end);
        cp_async16(bs + b_k[i] * LDB + b_n[i], p ? b_base[i] : Bb, p);
        b_base[i] += b_step; b_pos[i].k += BK;
      }
    } else if (KMODE == 2) {
#pragma unroll
      for (int i = 0; i < Cfg::A_PER_THR; ++i) {
        bool p = a_ok[i] && (a_kt + a_k[i] < k_end);
        cp_async16(as + a_k[i] * LDA + a_m[i], p ? a_base[i] : Ab, p);
        a_base[i] += a_step;
      }
      a_kt += BK; a_left -= BK;
      if (a_left <= 0 && a_kt < k_end) {          // the next tile starts a new level-1 block: shift every slot's pointer
        const long long nb = kpos_base(d.ak, ++a_k1), delta = nb - a_cur - (long long)d.ak.n0 * d.ak.s0;
#pragma unroll
        for (int i = 0; i < Cfg::A_PER_THR; ++i) a_base[i] += delta;
        a_cur = nb; a_left = d.ak.n0;
      }
#pragma unroll
      for (int i = 0; i < Cfg::B_PER_THR; ++i) {
        bool p = b_ok[i] && (b_kt + b_k[i] < k_end);
        cp_async16(bs + b_k[i] * LDB + b_n[i], p ? b_base[i] : Bb, p);
        b_base[i] += b_step;
      }
      b_kt += BK; b_left -= BK;
      if (b_left <= 0 && b_kt < k_end) {
        const long long nb = kpos_base(d.bk, ++b_k1), delta = nb - b_cur - (long long)d.bk.n0 * d.bk.s0;
#pragma unroll
        for (int i = 0; i < Cfg::B_PER_THR; ++i) b_base[i] += delta;
        b_cur = nb; b_left = d.bk.n0;
      }
    } else {
#pragma unroll
      for (int i = 0; i < Cfg::A_PER_THR; ++i) {
        bool p = a_ok[i] && (a_pos[i].k < k_end);
        long long off = p ? kpos_off(d.ak, a_pos[i]) : 0;
        cp_async16(as + a_k[i] * LDA + a_m[i], a_base[i] + off, p);
        kpos_adv(d.ak, a_pos[i], k_end);
      }
#pragma unroll
      for (int i = 0; i < Cfg::B_PER_THR; ++i) {
        bool p = b_ok[i] && (b_pos[i].k < k_end);
        long long off = p ? kpos_off(d.bk, b_pos[i]) : 0;
        cp_async16(bs + b_k[i] * LDB + b_n[i], b_base[i] + off, p);
        kpos_adv(d.bk, b_pos[i], k_end);
      }
    }
  };

  double cre[TM][TN][2], cim[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }

  const int ktiles = (k_end - k_begin + BK - 1) / BK;
  __syncthreads();   // persistent CTAs: the previous tile's last stages may still be being read
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s);
    cp_async_commit();
  }

  const unsigned fa = d.conjA ? 0x80000000u : 0u;   // conj-on-load: flip the sign of the imaginary fragment
  const unsigned fb = d.conjB ? 0x80000000u : 0u, fnb = fb ^ 0x80000000u;
  const int a_frag = wm * TM * 8 + g;       // + mi*8, row (k) = t
  const int b_frag = wn * TN * 8 + g;

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt+STAGES-1 into the slot freed at iteration kt-1
      int nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage(nk % STAGES);
      cp_async_commit();
    }
    const cplx* as = As + (kt % STAGES) * Cfg::A_STAGE;
    const cplx* bs = Bs + (kt % STAGES) * Cfg::B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double ar[TM], ai[TM], br[TN], bi[TN], nbi[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        cplx v = as[(kk + t) * LDA + a_frag + i * 8];
        ar[i] = v.x; ai[i] = sign_xor(v.y, fa);
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        cplx v = bs[(kk + t) * LDB + b_frag + j * 8];
        br[j] = v.x; bi[j] = sign_xor(v.y, fb); nbi[j] = sign_xor(v.y, fnb);
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma(cim[i][j][0], cim[i][j][1], ar[i], bi[j]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma(cre[i][j][0], cre[i][j][1], ai[i], nbi[j]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma(cim[i][j][0], cim[i][j][1], ai[i], br[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: C = alpha*acc + beta*C, strided store -----------------------------------------
  cplx* __restrict__ Cb = d.C + (long long)batch * d.bsC + (long long)split * d.ssC;
  const bool has_beta = (d.beta.x != 0.0 || d.beta.y != 0.0);
  long long noff[TN][2]; bool nok[TN][2];
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int n = n_blk + wn * TN * 8 + j * 8 + 2 * t + q;
      nok[j][q] = n < d.N;
      noff[j][q] = nok[j][q] ? idx_off_t(d.cn, n, batch) : 0;
    }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m_blk + wm * TM * 8 + i * 8 + g;
    if (m >= d.M) continue;
    long long moff = idx_off_t(d.cm, m, batch);
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!nok[j][q]) continue;
        double xr = cre[i][j][q], xi = cim[i][j][q];
        cplx o;
        o.x = d.alpha.x * xr - d.alpha.y * xi;
        o.y = d.alpha.x * xi + d.alpha.y * xr;
        cplx* p = Cb + moff + noff[j][q];
        if (atomic) {
          atomicAdd(&p->x, o.x);
          atomicAdd(&p->y, o.y);
          continue;
        }
        if (has_beta) {
          cplx c = *p;
          o.x += d.beta.x * c.x - d.beta.y * c.y;
          o.y += d.beta.x * c.y + d.beta.y * c.x;
        }
        *p = o;
      }
  }
}

// Raster order of the tiles.  Tiles are visited fast-index first; with `panel` > 0 the fast direction is cut into panels of that many
// tiles and a panel is walked down the whole slow direction before the next one starts.  Why: a wave of ~296 concurrent tiles then
// covers (296 / panel) x panel tiles, so it touches 296 / panel blocks of the slow-side operand and `panel` blocks of the fast-side
// operand instead of ~1 and ALL of them -- at the bench shape (M = 40960, N = 8192, K = 2048) the single-panel order streamed the whole
// 268 MB Theta from DRAM in every one of its 277 waves: ncu measured 83.6 GB of DRAM reads against 1.6 GB of operands
// (profiles/r02_matvec_stage_metrics.csv).  The launcher picks panel = sqrt(slots * BM / BN) (or BN / BM), the width that minimises
// the bytes a wave touches, whenever the operands exceed what L2 can hold.
__device__ __forceinline__ void raster_tile(long long lin, int nfast, int nslow, int panel, int& fast, int& slow) {
  if (panel <= 0 || panel >= nfast) { fast = (int)(lin % nfast); slow = (int)(lin / nfast); return; }
  const int full = nfast / panel;
  const long long span = (long long)panel * nslow;
  if (lin < (long long)full * span) {
    const int pnl = (int)(lin / span);
    const long long rem = lin - (long long)pnl * span;
    slow = (int)(rem / panel); fast = pnl * panel + (int)(rem % panel);
  } else {
    const long long rem = lin - (long long)full * span;
    const int pl = nfast - full * panel;
    slow = (int)(rem / pl); fast = full * panel + (int)(rem % pl);
  }
}

template <int WARPS_M, int WARPS_N, int TM, int TN, int KMODE>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MinBlocks<WARPS_M, WARPS_N, TM, TN>::value) zgemm_kernel(const GemmDesc d) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, TM, TN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* As = reinterpret_cast<cplx*>(smem_raw);
  cplx* Bs = As + Cfg::STAGES * Cfg::A_STAGE;
  const int z = blockIdx.z;
  const int batch = z / d.ksplit, split = z - batch * d.ksplit;
  if (d.skip != nullptr && d.skip[batch] != 0) return;   // e.g. identity rotation of an already converged Jacobi pair
  // raster: consecutive CTAs walk m (default) or n (swap_raster) so that the LARGER operand is streamed from DRAM once
  int fast = blockIdx.x, slow = blockIdx.y;
  if (d.panel > 0) raster_tile((long long)blockIdx.x + (long long)gridDim.x * blockIdx.y, (int)gridDim.x, (int)gridDim.y, d.panel, fast, slow);
  const int m_blk = (d.swap_raster ? slow : fast) * Cfg::BM, n_blk = (d.swap_raster ? fast : slow) * Cfg::BN;
  const int k_begin = split * d.kchunk;
  const int k_end = min(d.K, k_begin + d.kchunk);
  gemm_tile<WARPS_M, WARPS_N, TM, TN, KMODE>(d, As, Bs, m_blk, n_blk, batch, split, k_begin, k_end, d.atomic_c != 0);
}

// Persistent stream-K variant (batch == 1, no split-K, beta == 0, dense C zeroed by the launcher).  The tile space is
// walked in raster order.  Whole waves of tiles (dp_tiles) are processed one tile per CTA per wave.  The rem_tiles tiles
// that would form a partial last wave (or a sub-wave problem) are each cut into nseg equal k-segments, chosen so that
// the rem_tiles * nseg work items fill the CTA slots in as few rounds as possible; items are dealt so that the CTAs of a
// round work on neighbouring tiles over the SAME k range (they share operand rows / columns through L2, exactly like
// the tiles of a full wave -- a free-running unit split loses that reuse and cost 4x the DRAM reads).  Tiles with
// nseg > 1 are accumulated with red.global.add.f64.
// Problems of less than two waves of tiles (operands fit in L2 anyway) instead use one contiguous run of k-tile
// units per CTA (unit_ctas > 0): fewer, longer runs amortise the pipeline ramp and the atomic epilogue better.
struct SkPlan { int tiles_fast, tiles_slow, dp_tiles, rem_tiles, nseg, kt, unit_ctas; };
template <int WARPS_M, int WARPS_N, int TM, int TN, int KMODE>
__global__ void __launch_bounds__(WARPS_M * WARPS_N * 32, MinBlocks<WARPS_M, WARPS_N, TM, TN>::value) zgemm_sk_kernel(const GemmDesc d, const SkPlan pl) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, TM, TN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* As = reinterpret_cast<cplx*>(smem_raw);
  cplx* Bs = As + Cfg::STAGES * Cfg::A_STAGE;
  auto origin = [&](int tile, int& m_blk, int& n_blk) {
    int fast, slow;
    raster_tile(tile, pl.tiles_fast, pl.tiles_slow, d.panel, fast, slow);
    m_blk = (d.swap_raster ? slow : fast) * Cfg::BM;
    n_blk = (d.swap_raster ? fast : slow) * Cfg::BN;
  };
  const int c = blockIdx.x;
  for (int tile = c; tile < pl.dp_tiles; tile += gridDim.x) {
    int m_blk, n_blk;
    origin(tile, m_blk, n_blk);
    gemm_tile<WARPS_M, WARPS_N, TM, TN, KMODE>(d, As, Bs, m_blk, n_blk, 0, 0, 0, d.K, false);
  }
  if (pl.unit_ctas > 0) {
    if (c >= pl.unit_ctas) return;
    const long long units = (long long)pl.rem_tiles * pl.kt;
    long long u = units * c / pl.unit_ctas;
    const long long u_end = units * (c + 1) / pl.unit_ctas;
    while (u < u_end) {
      const int tile = (int)(u / pl.kt), k0 = (int)(u - (long long)tile * pl.kt);
      const int k1 = (int)min((long long)pl.kt, k0 + (u_end - u));
      int m_blk, n_blk;
      origin(tile, m_blk, n_blk);
      gemm_tile<WARPS_M, WARPS_N, TM, TN, KMODE>(d, As, Bs, m_blk, n_blk, 0, 0, k0 * Cfg::BK, min(d.K, k1 * Cfg::BK), !(k0 == 0 && k1 == pl.kt));
      u += k1 - k0;
    }
    return;
  }
  const int items = pl.rem_tiles * pl.nseg;
  for (int item = c; item < items; item += gridDim.x) {
    const int seg = item / pl.rem_tiles, tile = pl.dp_tiles + (item - seg * pl.rem_tiles);
    const int k0 = (int)((long long)pl.kt * seg / pl.nseg), k1 = (int)((long long)pl.kt * (seg + 1) / pl.nseg);
    int m_blk, n_blk;
    origin(tile, m_blk, n_blk);
    gemm_tile<WARPS_M, WARPS_N, TM, TN, KMODE>(d, As, Bs, m_blk, n_blk, 0, 0, k0 * Cfg::BK, min(d.K, k1 * Cfg::BK), pl.nseg > 1);
  }
}

// ---- 128 x 32 tile fed by the bulk-copy (TMA) engine ------------------------------------------------------------------------------
// Same tile, fragments and arithmetic as zgemm_kernel<4,1,4,4,1>, for the shapes of the two chi^3 stages of the H_eff application: full
// tiles, m unit-stride (a k-row of the A tile is 2 KiB of contiguous memory), B either k-fast (Theta: 8 contiguous elements per column)
// or n-fast (the right block: 32 contiguous elements per k-row).  The four warps no longer issue their own 16-byte cp.async copies: per
// k-tile ONE warp (they take turns) arms the stage's mbarrier with the byte count and issues 16 / 40 `cp.async.bulk` copies
// (SASS: UBLKCP) straight into the padded stage rows; consumers wait on the stage's "full" mbarrier and release it through an "empty"
// mbarrier -- no __syncthreads, no cp.async.wait_group, no per-thread pointer walk in the k loop.
namespace bulk {
constexpr int BM = 128, BN = 32, LDA = BM + 2, LDBN = BN + 2;
template <int BK_, int ST_>
struct Cfg {        // k-tile depth and ring length: 8 x 4 stages, or 16 x 2 stages (half as many stage hand-overs per DMMA); both 2 CTAs / SM
  static constexpr int BK = BK_, ST = ST_, LDBK = BK_ + 1;
  static constexpr int A_STAGE = BK * LDA, B_STAGE = (BN * LDBK > BK * LDBN) ? BN * LDBK : BK * LDBN;   // complex elements, either B layout
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int SMEM_BYTES = ST * STAGE * 16 + 2 * ST * 8;
  static constexpr unsigned TX_BYTES = BK * BM * 16 + BK * BN * 16;
};

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
}  // namespace bulk

template <bool BKFAST, int BK_, int ST_>
__global__ void __launch_bounds__(128, 2) zgemm_bulk_kernel(const GemmDesc d) {
  using namespace bulk;
  using C_ = bulk::Cfg<BK_, ST_>;
  constexpr int BK = C_::BK, ST = C_::ST, LDBK = C_::LDBK, A_STAGE = C_::A_STAGE, STAGE = C_::STAGE;
  constexpr unsigned TX_BYTES = C_::TX_BYTES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* S = reinterpret_cast<cplx*>(smem_raw);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + ST * STAGE * 16);
  unsigned long long* empty = full + ST;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  int fast = blockIdx.x, slow = blockIdx.y;
  if (d.panel > 0) raster_tile((long long)blockIdx.x + (long long)gridDim.x * blockIdx.y, (int)gridDim.x, (int)gridDim.y, d.panel, fast, slow);
  const int m_blk = (d.swap_raster ? slow : fast) * BM, n_blk = (d.swap_raster ? fast : slow) * BN;
  const cplx* Ag = d.A + m_blk;                                                     // am.s0 == 1
  const cplx* Bg = BKFAST ? d.B + (long long)n_blk * d.bn.s0 : d.B + n_blk;         // k-fast: bk.s0 == 1; n-fast: bn.s0 == 1
  const int ktiles = d.K / BK;
  if (tid == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  // executed by all lanes of one warp: refill the stage of k-tile `tile`
  auto issue = [&](int tile) {
    const int s = tile % ST;
    cplx* as = S + s * STAGE;
    cplx* bs = as + A_STAGE;
    if (lane == 0) {
      if (tile >= ST) mbar_wait(empty + s, (unsigned)((tile / ST - 1) & 1));       // every warp has finished the previous use of the stage
      mbar_expect_tx(full + s, TX_BYTES);
    }
    __syncwarp();
    const long long k0 = (long long)tile * BK;
    if (lane < BK) bulk_g2s(as + lane * LDA, Ag + (k0 + lane) * d.ak.s0, BM * 16, full + s);
    if (BKFAST) bulk_g2s(bs + lane * LDBK, Bg + (long long)lane * d.bn.s0 + k0, BK * 16, full + s);
    else if (lane >= BK && lane < 2 * BK) bulk_g2s(bs + (lane - BK) * LDBN, Bg + (k0 + lane - BK) * d.bk.s0, BN * 16, full + s);
  };
  if (warp == 0)
    for (int tile = 0; tile < ST - 1 && tile < ktiles; ++tile) issue(tile);

  double cre[4][4][2], cim[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }
  const unsigned fa = d.conjA ? 0x80000000u : 0u, fb = d.conjB ? 0x80000000u : 0u, fnb = fb ^ 0x80000000u;
  const int a_frag = warp * 32 + g;

  for (int kt = 0; kt < ktiles; ++kt) {
    const int s = kt % ST;
    mbar_wait(full + s, (unsigned)((kt / ST) & 1));
    const cplx* as = S + s * STAGE;
    const cplx* bs = as + A_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double ar[4], ai[4], br[4], bi[4], nbi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { cplx v = as[(kk + t) * LDA + a_frag + i * 8]; ar[i] = v.x; ai[i] = sign_xor(v.y, fa); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cplx v = BKFAST ? bs[(g + j * 8) * LDBK + kk + t] : bs[(kk + t) * LDBN + g + j * 8];
        br[j] = v.x; bi[j] = sign_xor(v.y, fb); nbi[j] = sign_xor(v.y, fnb);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(cim[i][j][0], cim[i][j][1], ar[i], bi[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(cre[i][j][0], cre[i][j][1], ai[i], nbi[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(cim[i][j][0], cim[i][j][1], ai[i], br[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    const int nt = kt + ST - 1;
    if (nt < ktiles && warp == (kt & 3)) issue(nt);
    static_assert(2 * BK <= 32, "one warp issues the A rows and the n-fast B rows of a stage");
  }

  // epilogue: C = alpha*acc + beta*C, strided store (as gemm_tile)
  cplx* __restrict__ Cb = d.C;
  const bool has_beta = (d.beta.x != 0.0 || d.beta.y != 0.0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m_blk + warp * 32 + i * 8 + g;
    const long long moff = idx_off_t(d.cm, m, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n = n_blk + j * 8 + 2 * t + q;
        const double xr = cre[i][j][q], xi = cim[i][j][q];
        cplx o;
        o.x = d.alpha.x * xr - d.alpha.y * xi;
        o.y = d.alpha.x * xi + d.alpha.y * xr;
        cplx* p = Cb + moff + idx_off_t(d.cn, n, 0);
        if (has_beta) { const cplx c = *p; o.x += d.beta.x * c.x - d.beta.y * c.y; o.y += d.beta.x * c.y + d.beta.y * c.x; }
        *p = o;
      }
  }
}

// TN_GEMM_BULK=1 / 2 moves the chi^3 stages to the bulk-copy kernel (8-deep k-tiles x 4 stages / 16-deep x 2 stages); default: cp.async kernel
static int bulk_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_GEMM_BULK"); v = e ? atoi(e) : 0; }
  return v;
}
static bool single_level(const Idx2& ix, int extent) { return ix.tab == nullptr && (ix.n0 >= extent || ix.s1 == (long long)ix.n0 * ix.s0); }
// 0: not applicable, 1: B k-fast, 2: B n-fast
static int bulk_mode(const GemmDesc& d) {
  using namespace bulk;
  const int BK = bulk_variant() == 2 ? 16 : 8, ST = bulk_variant() == 2 ? 2 : 4;
  if (bulk_variant() <= 0 || d.batch != 1 || d.ksplit != 1 || d.atomic_c || d.skip != nullptr) return 0;
  if (d.M % BM || d.N % BN || d.K % BK || d.K < BK * ST) return 0;
  if (!single_level(d.am, d.M) || d.am.s0 != 1 || !single_level(d.ak, d.K) || !single_level(d.bk, d.K) || !single_level(d.bn, d.N)) return 0;
  if (d.cm.tab || d.cn.tab || d.C == d.A || d.C == d.B) return 0;
  if (d.bk.s0 == 1 && d.bn.s0 != 1) return 1;
  if (d.bn.s0 == 1) return 2;
  return 0;
}

template <int WARPS_M, int WARPS_N, int TM, int TN, int KMODE>
static void launch2(const GemmDesc& d, cudaStream_t stream) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, TM, TN>;
  static DeviceOnce configured;
  auto kern = zgemm_kernel<WARPS_M, WARPS_N, TM, TN, KMODE>;
  configured.run([&] { TN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); });
  if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return;
  unsigned tm = (d.M + Cfg::BM - 1) / Cfg::BM, tn_ = (d.N + Cfg::BN - 1) / Cfg::BN;
  GemmDesc dd = d;
  // Within a wave the CTAs share the operand indexed by the slow raster direction through L2; across waves the other
  // operand is re-read from DRAM.  Stream the operand with more bytes (A: M*K, B: K*N) only once.
  dd.swap_raster = (d.batch * d.ksplit == 1 && (long long)d.M > (long long)d.N && tm <= 65535 && tn_ > 1) ? 1 : 0;
  dd.panel = 0;
  if (d.batch * d.ksplit == 1 && 16.0 * ((double)d.M * d.K + (double)d.K * d.N) > 96.0 * 1024 * 1024) {
    // operands beyond L2: panels along the fast direction (see raster_tile); ~296 CTA slots on a B200
    const double ratio = dd.swap_raster ? (double)Cfg::BM / Cfg::BN : (double)Cfg::BN / Cfg::BM;
    const int nfast = dd.swap_raster ? (int)tn_ : (int)tm;
    static double scale = -1;
    if (scale < 0) { const char* e = getenv("TN_GEMM_PANEL_SCALE"); scale = e ? atof(e) : 1.0; }
    const int pw = std::max(1, (int)std::lround(scale * std::sqrt(296.0 * ratio)));
    if (pw < nfast) dd.panel = pw;
  }
  if (d.streamk) {
    // stream-K: only when the plain launch would leave a partial last wave (or less than one wave) of CTAs
    static int slots = 0;          // identical B200s: the value of the first device holds for all
    static DeviceOnce sk_configured;
    auto kern_sk = zgemm_sk_kernel<WARPS_M, WARPS_N, TM, TN, KMODE>;
    sk_configured.run([&] {
      TN_CUDA(cudaFuncSetAttribute(kern_sk, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      int per_sm = 0, dev = 0, sms = 0;
      TN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_sk, Cfg::THREADS, Cfg::SMEM_BYTES));
      TN_CUDA(cudaGetDevice(&dev));
      TN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      slots = std::max(1, per_sm * sms);
    });
    const long long T = (long long)tm * tn_;
    const int KT = (d.K + Cfg::BK - 1) / Cfg::BK;
    const long long waves = (T + slots - 1) / slots;
    const double eff = (double)T / (double)(waves * slots);
    if (eff < 0.94 && KT >= 16 && T < (1ll << 30)) {
      SkPlan pl;
      pl.tiles_fast = dd.swap_raster ? (int)tn_ : (int)tm;
      pl.tiles_slow = dd.swap_raster ? (int)tm : (int)tn_;
      pl.dp_tiles = (T < 2 * (long long)slots) ? 0 : (int)((T / slots) * slots);   // under two waves: one unit run per CTA
      pl.rem_tiles = (int)(T - pl.dp_tiles);
      pl.kt = KT;
      // segments per remainder tile: minimise the duration ceil(rem * nseg / slots) / nseg of the last phase
      // (in units of one full tile), keeping at least 8 k-tiles per segment
      pl.nseg = 1;
      double best = 1e30;
      for (int ns = 1; ns <= 16 && KT / ns >= 8; ++ns) {
        const double dur = (double)(((long long)pl.rem_tiles * ns + slots - 1) / slots) / ns;
        if (dur < best - 1e-9) { best = dur; pl.nseg = ns; }
      }
      pl.unit_ctas = 0;
      // Under two waves with operands that fit in L2: contiguous unit runs.  With operands beyond L2 (e.g. the a'-slices of the
      // pipelined host-buffer matvec: M = 8192, N = 256, K = 40960, A = 5.4 GB) the unit runs read A once PER N-TILE -- ncu measured
      // 44 GB of DRAM reads for such a launch -- whereas the lockstep segments above keep the CTAs of neighbouring tiles on the
      // same k range (profiles/r02_ncu_full_summary.json).
      const double operand_bytes = 16.0 * ((double)d.M * d.K + (double)d.K * d.N);
      if (pl.dp_tiles == 0 && operand_bytes <= 64.0 * 1024 * 1024) {   // small problem: contiguous unit runs, at least 16 k-tiles per CTA
        pl.unit_ctas = (int)std::max<long long>(1, std::min<long long>(slots, (long long)pl.rem_tiles * KT / 16));
        pl.nseg = 2;            // (marks the launch as split)
      }
      if (pl.nseg > 1) {     // otherwise no split beats the plain partial wave: fall through to the ordinary launch
        const int grid = pl.unit_ctas > 0 ? pl.unit_ctas
                                          : (int)std::min<long long>(slots, std::max<long long>(pl.dp_tiles, (long long)pl.rem_tiles * pl.nseg));
        // partial tiles are accumulated atomically: zero the (dense, possibly pitched) output first
        TN_CUDA(cudaMemset2DAsync(d.C, (size_t)d.cn.s0 * sizeof(cplx), 0, (size_t)d.M * sizeof(cplx), (size_t)d.N, stream));
        kern_sk<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(dd, pl);
        TN_CUDA(cudaGetLastError());
        count_launch(1);
        return;
      }
    }
  }
  dim3 grid(dd.swap_raster ? tn_ : tm, dd.swap_raster ? tm : tn_, d.batch * d.ksplit);
  TN_CHECK(grid.y <= 65535 && grid.z <= 65535, "zgemm: grid too large");
  if (WARPS_M == 4 && WARPS_N == 1 && TM == 4 && TN == 4 && KMODE == 1) {
    const int mode = bulk_mode(dd);
    if (mode != 0) {
      static DeviceOnce bcfg;
      bcfg.run([&] {
        TN_CUDA(cudaFuncSetAttribute(zgemm_bulk_kernel<true, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bulk::Cfg<8, 4>::SMEM_BYTES));
        TN_CUDA(cudaFuncSetAttribute(zgemm_bulk_kernel<false, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bulk::Cfg<8, 4>::SMEM_BYTES));
        TN_CUDA(cudaFuncSetAttribute(zgemm_bulk_kernel<true, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bulk::Cfg<16, 2>::SMEM_BYTES));
        TN_CUDA(cudaFuncSetAttribute(zgemm_bulk_kernel<false, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bulk::Cfg<16, 2>::SMEM_BYTES));
      });
      if (bulk_variant() == 2) {
        if (mode == 1) zgemm_bulk_kernel<true, 16, 2><<<grid, 128, bulk::Cfg<16, 2>::SMEM_BYTES, stream>>>(dd);
        else zgemm_bulk_kernel<false, 16, 2><<<grid, 128, bulk::Cfg<16, 2>::SMEM_BYTES, stream>>>(dd);
      } else {
        if (mode == 1) zgemm_bulk_kernel<true, 8, 4><<<grid, 128, bulk::Cfg<8, 4>::SMEM_BYTES, stream>>>(dd);
        else zgemm_bulk_kernel<false, 8, 4><<<grid, 128, bulk::Cfg<8, 4>::SMEM_BYTES, stream>>>(dd);
      }
      TN_CUDA(cudaGetLastError());
      count_launch(1);
      return;
    }
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(dd);
  TN_CUDA(cudaGetLastError());
  count_launch(1);
}

template <int WARPS_M, int WARPS_N, int TM, int TN>
static void launch(const GemmDesc& d, cudaStream_t stream) {
  using Cfg = TileCfg<WARPS_M, WARPS_N, TM, TN>;
  const bool simple = d.ak.tab == nullptr && d.bk.tab == nullptr && d.ak.n0 >= d.K && d.bk.n0 >= d.K;
  // level-1 blocks (or the whole index) are multiples of BK and every split starts on a tile boundary
  auto blk_ok = [&](const Idx2& ix) { return ix.n0 >= d.K || ix.n0 % Cfg::BK == 0; };
  const bool aligned = blk_ok(d.ak) && blk_ok(d.bk) && (d.ksplit <= 1 || d.kchunk % Cfg::BK == 0);
  if (simple) launch2<WARPS_M, WARPS_N, TM, TN, 1>(d, stream);
  else if (aligned) launch2<WARPS_M, WARPS_N, TM, TN, 2>(d, stream);
  else launch2<WARPS_M, WARPS_N, TM, TN, 0>(d, stream);
}

void zgemm(const GemmDesc& d, cudaStream_t stream) { launch<4, 2, 4, 4>(d, stream); }

// Chooses the CTA tile from the problem shape and the load mapping from which sub-index is unit-stride.
void zgemm_auto(GemmDesc d, cudaStream_t stream) {
  if (d.ksplit <= 0) { d.ksplit = 1; }
  if (d.ksplit == 1) d.kchunk = d.K;
  if (d.batch <= 0) d.batch = 1;
  d.a_kfast = (d.ak.s0 == 1 && d.am.s0 != 1) ? 1 : 0;
  d.b_kfast = (d.bk.s0 == 1 && d.bn.s0 != 1) ? 1 : 0;
  // In-place updates (C aliases an operand: Jacobi / CholeskyQR panel rotations Z(:,pair) <- Z(:,pair) * J) are only
  // safe when one CTA owns every column of its row block, i.e. a single n-tile of 64.
  const bool inplace = (d.C == d.A || d.C == d.B);
  if (inplace) {
    TN_CHECK(d.N <= 64, "zgemm: in-place update needs N <= 64");
    launch<4, 2, 4, 4>(d, stream);
    return;
  }
  static int small_variant = -1;
  if (small_variant < 0) { const char* e = getenv("TN_GEMM_SMALL"); small_variant = e ? atoi(e) : 0; }
  if (d.M <= 64 && d.N <= 64) {                                      // 64 x 64   (Gram blocks, tiny bonds)
    if (small_variant == 1) launch<2, 2, 4, 4>(d, stream);          // 4 warps with 32 x 32 warp tiles, 2 CTAs / SM
    else launch<2, 4, 4, 2>(d, stream);                             // 8 warps with 32 x 16 warp tiles, 2 CTAs / SM
  }
  else if (d.N <= 32)              launch<8, 1, 4, 4>(d, stream);   // 256 x 32  (skinny right operand)
  else if (d.M <= 64)              launch<2, 4, 4, 4>(d, stream);   // 64 x 128
  else {
    // Large problems.  Default: 128 x 32 tiles, 4 warps, 2 CTAs per SM -- the two CTAs' k-tile barriers
    // de-synchronise so the DMMA pipe does not drain at every barrier (measured 29.8 vs 27.4 TFLOP/s for the
    // chi=1024 matvec against the single-CTA 128 x 64 tile).  TN_GEMM_VARIANT selects alternatives for A/B runs.
    static int variant = -1, use_sk = -1;
    if (variant < 0) { const char* e = getenv("TN_GEMM_VARIANT"); variant = e ? atoi(e) : 2; }
    if (use_sk < 0) { const char* e = getenv("TN_GEMM_STREAMK"); use_sk = (e && e[0] == '0') ? 0 : 1; }
    // stream-K needs a dense single-level C it may zero (leading dimension cn.s0 >= M), beta == 0 and a single problem
    d.streamk = (use_sk && d.batch == 1 && d.ksplit == 1 && d.beta.x == 0.0 && d.beta.y == 0.0 && !d.cm.tab && !d.cn.tab &&
                 d.cm.s0 == 1 && d.cm.n0 >= d.M && d.cn.n0 >= d.N && d.cn.s0 >= d.M && d.C != d.A && d.C != d.B &&
                 (double)d.M * d.N * d.K >= 8.0e6) ? 1 : 0;
    if (variant == 0) launch<4, 2, 4, 4>(d, stream);                 // 128 x 64, 1 CTA / SM
    else if (variant == 1) launch<2, 2, 4, 4>(d, stream);            // 64 x 64, 2 CTAs / SM
    else if (variant == 3) launch<2, 1, 4, 4>(d, stream);            // 64 x 32, 4 CTAs / SM
    else if (variant == 4) launch<1, 2, 4, 4>(d, stream);            // 32 x 64, 4 CTAs / SM
    else if (variant == 5) launch<4, 2, 4, 2>(d, stream);            // 128 x 32 with 8 warps of 32 x 16: 16 warps / SM (4 per scheduler)
    else if (variant == 6) launch<2, 4, 4, 2>(d, stream);            // 64 x 64 with 8 warps of 32 x 16: 16 warps / SM
    else launch<4, 1, 4, 4>(d, stream);                              // 128 x 32, 2 CTAs / SM
  }
}

}  // namespace tn
