// MPO-bond-sharded H_eff application behind the C ABI (SURVEY 8(e), 8(b) "context owns the NCCL communicators"): ONE process,
// G devices, one stream per device, NCCL reduce_scatter + all_reduce inside the library -- so a ccall caller (the Julia shim)
// reaches the multi-GPU matvec without any Python.  The reference is single-process (projmps.jl:103-145); the result equals
// tn_env_product up to rounding.
//
// Split (even for any MPO bond dimension, as tnb200.sharded.BalancedShardedHeff):
//   stage 1  device g owns rows [m0, m1) of the fused (a, w) index of L:  T1_g = L_g . Theta          (chi^3 d^2 w / G)
//   stage 2  T2p(a,s1,s2,[b',w2]) = sum_{w in g,s1',s2'} T1_g W_g   -- partial sums for ALL (b', w2)     (small)
//   exchange ncclReduceScatter over the fused contraction index k = (b', w2) in equal chunks -- Theta's right bond is cut into slices
//            and the exchange of slice j runs on a communication stream under stage 1 + 2 of slice j + 1 (events, no host sync)
//   stage 3  out_p = T2_g . R_g[k in chunk g, a']                                                        (chi^3 d^2 w2 / G)
//   exchange ncclAllReduce(out_p)
// NCCL is resolved at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it and reports a clear error
// when it is absent.
#include "../../include/tn_c_api.h"
#include "tn_mps.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <memory>

namespace tn {
void count_launch(int n);

namespace {
typedef void* nccl_comm_t;
struct Nccl {
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;      // ncclFloat64, ncclSum (nccl.h)

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    auto sym = [&](const char* s) { return dlsym(h, s); };
    n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    n.ReduceScatter = (decltype(n.ReduceScatter))sym("ncclReduceScatter");
    n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    n.ok = n.CommInitAll && n.CommDestroy && n.GroupStart && n.GroupEnd && n.ReduceScatter && n.AllReduce && n.GetErrorString;
  });
  return n;
}
#define TN_NCCL(x)                                                                                        \
  do {                                                                                                    \
    int r_ = (x);                                                                                         \
    if (r_ != 0) throw tn::Error(-2, std::string("NCCL error: ") + nccl().GetErrorString(r_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

struct SlicePart {                       // one slice [b0, b0 + nb) of Theta's right bond on one device
  cplx *Rg = nullptr, *T2p = nullptr, *T2g = nullptr;
  cudaEvent_t ev12 = nullptr, evx = nullptr;
};
struct DevPart {
  int device = 0;
  cudaStream_t stream = nullptr, comm = nullptr;   // contractions / collectives
  int mloc = 0, nw = 0, r0 = 0;
  cplx *Lg = nullptr, *Wg = nullptr, *theta = nullptr, *T1 = nullptr, *out = nullptr;
  std::vector<SlicePart> sl;
};
}  // namespace

struct Shard {
  int G = 0;
  int ca, cb, ca2, cb2, d, w, w1, w2;
  cplx coeff;
  std::vector<long long> cut, c;       // slices of Theta's right bond: [cut[j], cut[j+1]); c[j] = rows of the fused (b', w2) index of slice j per device
  std::vector<DevPart> parts;
  std::vector<nccl_comm_t> comms;
  ~Shard() {
    for (auto cm : comms) if (cm) nccl().CommDestroy(cm);
    for (auto& p : parts) {
      cudaSetDevice(p.device);
      if (p.stream) cudaStreamSynchronize(p.stream);
      if (p.comm) cudaStreamSynchronize(p.comm);
      for (cplx* q : {p.Lg, p.Wg, p.theta, p.T1, p.out}) if (q) cudaFree(q);
      for (auto& q : p.sl) {
        for (cplx* b : {q.Rg, q.T2p, q.T2g}) if (b) cudaFree(b);
        if (q.ev12) cudaEventDestroy(q.ev12);
        if (q.evx) cudaEventDestroy(q.evx);
      }
      if (p.comm) cudaStreamDestroy(p.comm);
      if (p.stream) cudaStreamDestroy(p.stream);
    }
  }
};

static cplx* dev_zeros(size_t n) {
  cplx* p = nullptr;
  TN_CUDA(cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(cplx)));
  TN_CUDA(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(cplx)));
  return p;
}
static cplx* dev_copy(const std::vector<cplx>& h) {
  cplx* p = dev_zeros(h.size());
  if (!h.empty()) TN_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  return p;
}

Shard* shard_create(int G, const int* devices, long long ca, long long ca2, int d, long long w, long long w1, long long w2,
                    const cplx* L, const cplx* R, const cplx* M1, const cplx* M2, cplx coeff) {
  TN_CHECK(G >= 1 && G <= 64 && devices, "sharded matvec: bad device list");
  TN_CHECK(ca >= 1 && ca2 >= 1 && d >= 1 && w >= 1 && w1 >= 1 && w2 >= 1 && L && R && M1 && M2, "sharded matvec: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw tn::Error(TN_ERR_CUDA, "no CUDA device available: libtnb200 has no CPU fallback");
  for (int g = 0; g < G; ++g) {
    TN_CHECK(devices[g] >= 0 && devices[g] < ndev, "sharded matvec: device index out of range");
    for (int h = 0; h < g; ++h) TN_CHECK(devices[h] != devices[g], "sharded matvec: a device is listed twice");
  }
  if (G > 1 && !nccl().ok) throw tn::Error(TN_ERR_CUDA, "sharded matvec: libnccl.so.2 could not be loaded");
  auto sh = std::make_unique<Shard>();
  sh->G = G; sh->ca = (int)ca; sh->cb = (int)ca; sh->ca2 = (int)ca2; sh->cb2 = (int)ca2; sh->d = d; sh->w = (int)w; sh->w1 = (int)w1; sh->w2 = (int)w2;
  sh->coeff = coeff;
  const long long d2 = (long long)d * d, mtot = ca * w, cb = ca, cb2 = ca2;
  // slices of Theta's right bond: the reduce_scatter of slice j runs on the communication stream under stage 1 + 2 of slice j + 1
  int S = 1;
  if (G > 1) { const char* e = getenv("TN_SHARD_SLICES"); S = e ? std::max(1, atoi(e)) : std::max(4, G); }
  S = (int)std::min<long long>(S, cb2);
  sh->cut.resize(S + 1); sh->c.resize(S);
  for (int j = 0; j <= S; ++j) sh->cut[j] = cb2 * j / S;
  for (int j = 0; j < S; ++j) sh->c[j] = ((sh->cut[j + 1] - sh->cut[j]) * w2 + G - 1) / G;
  // W[(w,s1',s2'),(s1,s2,w2)] = sum_{w1} M1(w,s1,s1',w1) M2(w1,s2,s2',w2)  (host, tiny)
  const long long KW = w * d2, NW = d2 * w2;
  std::vector<cplx> Wfull((size_t)(KW * NW));
  for (long long e = 0; e < KW * NW; ++e) {
    const long long kk = e % KW, nn = e / KW;
    const long long iw = kk % w, s1p = (kk / w) % d, s2p = kk / (w * d);
    const long long s1 = nn % d, s2 = (nn / d) % d, iw2 = nn / d2;
    double xr = 0, xi = 0;
    for (long long j = 0; j < w1; ++j) {
      const cplx a = M1[iw + w * (s1 + d * (s1p + d * j))];
      const cplx b = M2[j + w1 * (s2 + d * (s2p + d * iw2))];
      xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x;
    }
    Wfull[(size_t)e] = cplx{xr, xi};
  }
  sh->parts.resize(G);
  for (int g = 0; g < G; ++g) {
    DevPart& p = sh->parts[g];
    p.device = devices[g];
    TN_CUDA(cudaSetDevice(p.device));
    cudaDeviceProp prop; TN_CUDA(cudaGetDeviceProperties(&prop, p.device));
    if (prop.major != 10) throw tn::Error(TN_ERR_CUDA, std::string("libtnb200 is built for sm_100a (B200) only; found ") + prop.name);
    TN_CUDA(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
    TN_CUDA(cudaStreamCreateWithFlags(&p.comm, cudaStreamNonBlocking));
    const long long m0 = mtot * g / G, m1 = mtot * (g + 1) / G;
    p.mloc = (int)(m1 - m0);
    const long long w_first = p.mloc > 0 ? m0 / ca : 0, w_last = p.mloc > 0 ? (m1 - 1) / ca : 0;
    p.nw = (int)(w_last - w_first + 1);
    p.r0 = (int)(m0 - ca * w_first);
    std::vector<cplx> Lg((size_t)p.mloc * cb), Wg((size_t)p.nw * d2 * NW);
    for (long long b = 0; b < cb; ++b) for (long long m = 0; m < p.mloc; ++m) Lg[(size_t)(m + p.mloc * b)] = L[(m0 + m) + mtot * b];
    for (long long nn = 0; nn < NW; ++nn)
      for (long long sp = 0; sp < d2; ++sp)
        for (long long wl = 0; wl < p.nw; ++wl) Wg[(size_t)(wl + p.nw * sp + p.nw * d2 * nn)] = Wfull[(size_t)((w_first + wl) + w * sp + KW * nn)];
    p.Lg = dev_copy(Lg); p.Wg = dev_copy(Wg);
    p.theta = dev_zeros((size_t)(cb * d2 * cb2));
    p.T1 = dev_zeros((size_t)(ca * p.nw * d2 * cb2));                     // rows outside [r0, r0 + mloc) stay zero
    p.out = dev_zeros((size_t)(ca * d2 * ca2));
    p.sl.resize(S);
    for (int j = 0; j < S; ++j) {
      SlicePart& q = p.sl[j];
      const long long b0 = sh->cut[j], nb = sh->cut[j + 1] - b0, ktot = nb * w2, c = sh->c[j];
      const long long k0 = std::min<long long>(g * c, ktot), k1 = std::min<long long>((g + 1) * c, ktot);
      std::vector<cplx> Rg((size_t)(c * ca2), cplx{0, 0});
      for (long long ap = 0; ap < ca2; ++ap)
        for (long long k = k0; k < k1; ++k) {
          const long long bl = k % nb, iw2 = k / nb;                        // k = b'_local + nb * w2
          Rg[(size_t)((k - k0) + c * ap)] = R[ap + ca2 * (iw2 + w2 * (b0 + bl))];
        }
      q.Rg = dev_copy(Rg);
      q.T2p = dev_zeros((size_t)(ca * d2 * c * G));                       // the padding beyond k = ktot stays zero
      q.T2g = dev_zeros((size_t)(ca * d2 * c));
      TN_CUDA(cudaEventCreateWithFlags(&q.ev12, cudaEventDisableTiming));
      TN_CUDA(cudaEventCreateWithFlags(&q.evx, cudaEventDisableTiming));
    }
  }
  if (G > 1) {
    sh->comms.assign(G, nullptr);
    TN_NCCL(nccl().CommInitAll(sh->comms.data(), G, devices));
  }
  return sh.release();
}

void shard_free(Shard* s) { delete s; }

// theta_host (cb, d, d, cb2) -> out_host (ca, d, d, ca2); both borrowed for the call.  Everything is enqueued by this one host thread:
// per device a contraction stream and a communication stream, ordered by events (no host synchronisation before the end).
void shard_apply(Shard* sh, const cplx* theta_host, cplx* out_host) {
  const int G = sh->G, S = (int)sh->c.size();
  const long long ca = sh->ca, cb = sh->cb, ca2 = sh->ca2, cb2 = sh->cb2, d2 = (long long)sh->d * sh->d, w2 = sh->w2;
  const long long n_in = cb * d2 * cb2, n_out = ca * d2 * ca2;
  auto stage3 = [&](int j) {            // out (+)= coeff sum_k T2g_j[(a,s1,s2), k] Rg_j[k, a']
    for (int g = 0; g < G; ++g) {
      DevPart& p = sh->parts[g];
      TN_CUDA(cudaSetDevice(p.device));
      if (G > 1) TN_CUDA(cudaStreamWaitEvent(p.stream, p.sl[j].evx, 0));
      const cplx* T2 = G > 1 ? p.sl[j].T2g : p.sl[j].T2p;
      zgemm_auto(mk((int)(ca * d2), (int)ca2, (int)sh->c[j], T2, idx1(1), idx1(ca * d2), 0, p.sl[j].Rg, idx1(1), idx1(sh->c[j]), 0,
                    p.out, idx1(1), idx1(ca * d2), sh->coeff, j == 0 ? cplx{0.0, 0.0} : cplx{1.0, 0.0}), p.stream);
    }
  };
  for (int g = 0; g < G; ++g) {
    DevPart& p = sh->parts[g];
    TN_CUDA(cudaSetDevice(p.device));
    TN_CUDA(cudaMemcpyAsync(p.theta, theta_host, (size_t)n_in * sizeof(cplx), cudaMemcpyHostToDevice, p.stream));
  }
  for (int j = 0; j < S; ++j) {
    const long long b0 = sh->cut[j], nb = sh->cut[j + 1] - b0;
    for (int g = 0; g < G; ++g) {
      DevPart& p = sh->parts[g];
      SlicePart& q = p.sl[j];
      TN_CUDA(cudaSetDevice(p.device));
      if (p.mloc > 0) {
        const long long ld1 = ca * p.nw;
        cplx* T1j = p.T1 + ld1 * d2 * b0;
        // T1[r0 + m, (s1',s2',b')] = L_g[m, b] Theta[b, (s1',s2',b')]   for b' in the slice
        zgemm_auto(mk(p.mloc, (int)(d2 * nb), (int)cb, p.Lg, idx1(1), idx1(p.mloc), 0, p.theta + cb * d2 * b0, idx1(1), idx1(cb), 0, T1j + p.r0, idx1(1), idx1(ld1)), p.stream);
        // T2p_j(a,s1,s2,[b',w2]) = sum_{(w,s1',s2')} T1(a,(w,s1',s2'),b') W_g[(w,s1',s2'),(s1,s2,w2)]
        zgemm_auto(mk((int)(ca * nb), (int)(d2 * w2), (int)(p.nw * d2), T1j, idx2((int)ca, 1, ld1 * d2), idx1(ca), 0, p.Wg, idx1(1), idx1(p.nw * d2), 0,
                      q.T2p, idx2((int)ca, 1, ca * d2), idx2((int)d2, ca, ca * d2 * nb)), p.stream);
      } else {
        TN_CUDA(cudaMemsetAsync(q.T2p, 0, (size_t)(ca * d2 * sh->c[j] * G) * sizeof(cplx), p.stream));
      }
      if (G > 1) {
        TN_CUDA(cudaEventRecord(q.ev12, p.stream));
        TN_CUDA(cudaStreamWaitEvent(p.comm, q.ev12, 0));
      }
    }
    if (G > 1) {
      TN_NCCL(nccl().GroupStart());
      for (int g = 0; g < G; ++g) {
        DevPart& p = sh->parts[g];
        TN_NCCL(nccl().ReduceScatter(p.sl[j].T2p, p.sl[j].T2g, (size_t)(2 * ca * d2 * sh->c[j]), NCCL_DOUBLE, NCCL_SUM, sh->comms[g], p.comm));
      }
      TN_NCCL(nccl().GroupEnd());
      for (int g = 0; g < G; ++g) { DevPart& p = sh->parts[g]; TN_CUDA(cudaSetDevice(p.device)); TN_CUDA(cudaEventRecord(p.sl[j].evx, p.comm)); }
    }
    if (j >= 1) stage3(j - 1);           // its exchange overlapped stage 1 + 2 of slice j
  }
  stage3(S - 1);
  if (G > 1) {
    TN_NCCL(nccl().GroupStart());
    for (int g = 0; g < G; ++g) {
      DevPart& p = sh->parts[g];
      TN_NCCL(nccl().AllReduce(p.out, p.out, (size_t)(2 * n_out), NCCL_DOUBLE, NCCL_SUM, sh->comms[g], p.stream));
    }
    TN_NCCL(nccl().GroupEnd());
  }
  // every device holds the result; device 0's copy goes back to the host, the others are only waited for
  DevPart& p0 = sh->parts[0];
  TN_CUDA(cudaSetDevice(p0.device));
  TN_CUDA(cudaMemcpyAsync(out_host, p0.out, (size_t)n_out * sizeof(cplx), cudaMemcpyDeviceToHost, p0.stream));
  for (int g = 0; g < G; ++g) { TN_CUDA(cudaSetDevice(sh->parts[g].device)); TN_CUDA(cudaStreamSynchronize(sh->parts[g].stream)); TN_CUDA(cudaStreamSynchronize(sh->parts[g].comm)); }
}

}  // namespace tn
