// Shared declarations for libtnb200 (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>
#include <vector>
#include <mutex>

namespace tn {

typedef double2 cplx;   // interleaved (re, im): binary-identical to Julia ComplexF64 / numpy complex128

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define TN_CUDA(x)                                                                                  \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess)                                                                          \
      throw tn::Error(-2, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + \
                              ":" + std::to_string(__LINE__));                                      \
  } while (0)

// Function attributes (opt-in shared memory sizes) are per device: run `fn` once for each device a kernel is launched on.
struct DeviceOnce {
  std::once_flag flags[32];
  template <class F> void run(F&& fn) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(flags[dev & 31], fn);
  }
};

#define TN_CHECK(cond, msg)                   \
  do {                                        \
    if (!(cond)) throw tn::Error(-1, (msg));  \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Strided two-level index: i = i0 + n0 * i1  ->  element offset i0*s0 + lvl1(i1)*s1
// where lvl1(i1) = tab ? tab[tab_batch_stride*batch + i1] : i1.
// This is how a fused ("combined") tensor index is addressed without any permute pass in HBM.
// ---------------------------------------------------------------------------------------------
struct Idx2 {
  int n0;            // extent of the fast sub-index (>= total extent  => single level)
  long long s0, s1;  // element strides of the two sub-indices
  const int* tab;    // optional level-1 lookup table (Jacobi column-block pairs)
  int tab_bs;        // table entries per batch
};

static inline Idx2 idx1(long long stride) { return Idx2{0x7fffffff, stride, 0, nullptr, 0}; }
static inline Idx2 idx2(int n0, long long s0, long long s1) { return Idx2{n0, s0, s1, nullptr, 0}; }

// C[m,n] = alpha * sum_k opA(A)[m,k] * opB(B)[k,n] + beta * C[m,n]      (per batch, per k-split)
struct GemmDesc {
  int M, N, K;
  const cplx* A; Idx2 am, ak; int conjA;
  const cplx* B; Idx2 bk, bn; int conjB;
  cplx* C;       Idx2 cm, cn;
  cplx alpha, beta;
  int batch;               // number of independent problems (grid.z = batch * ksplit)
  long long bsA, bsB, bsC; // batch strides (elements)
  int ksplit;              // split-K factor; split s handles k in [s*kchunk, (s+1)*kchunk)
  int kchunk;
  long long ssC;           // C stride between k-splits (partials are summed by the consumer)
  int swap_raster;         // set by the launcher
  int panel;               // set by the launcher: width (in tiles) of the raster panels along the fast tile direction, 0 = one panel
  int a_kfast, b_kfast;    // global-load thread mapping: 1 = consecutive threads walk k (k is the unit-stride index)
  const int* skip;         // optional per-batch flags (device): CTAs of a batch with skip[batch] != 0 exit immediately
  int atomic_c;            // split-K contributions are added to C with red.global.add.f64 (ssC = 0, caller zeroes C, beta ignored)
  int streamk;             // set by zgemm_auto: the launcher may use the persistent stream-K kernel (zeroes C first)
};

void zgemm(const GemmDesc& d, cudaStream_t stream);          // 128x64 tile (default)
void zgemm_auto(GemmDesc d, cudaStream_t stream);            // picks tile + load mapping from the strides
long long zgemm_launch_count();                              // kernels launched so far (bench: gpu_launches)

}  // namespace tn
