// extern "C" boundary of libtnb200 (declared in include/tn_c_api.h).
#include "../../include/tn_c_api.h"
#include "tn_mps.cuh"
#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

namespace tn {
int qjmc_run(Mps* psi, Gates* gates, int njump, const int* jump_sites, const cplx* jump_ops, const double* jump_coeffs,
             int steps, double dt, Trunc tr, const double* uniforms, uint64_t seed, uint64_t traj,
             const cplx* obs_op, int save_every, cplx* obs_out, int* jumps_out, double* jumptimes_out, int jump_cap, bool classical);
}

namespace tn {
struct Shard;
Shard* shard_create(int G, const int* devices, long long ca, long long ca2, int d, long long w, long long w1, long long w2,
                    const cplx* L, const cplx* R, const cplx* M1, const cplx* M2, cplx coeff);
void shard_apply(Shard* sh, const cplx* theta_host, cplx* out_host);
void shard_free(Shard* s);
}

using namespace tn;

// One context = one GPU + one stream + its workspaces (also used by the worker threads of tn_qjmc_ensemble).
static void ctx_init(Ctx& c, int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) throw tn::Error(TN_ERR_CUDA, "no CUDA device available: libtnb200 has no CPU fallback");
  TN_CHECK(device >= 0 && device < ndev, "device index out of range");
  TN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop; TN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) throw tn::Error(TN_ERR_CUDA, std::string("libtnb200 is built for sm_100a (B200) only; found ") + prop.name);
  c.device = device;
  TN_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaMemPool_t pool; TN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX; TN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  TN_CUDA(cudaMalloc((void**)&c.dscal, 64 * sizeof(cplx)));
  TN_CUDA(cudaMemset(c.dscal, 0, 64 * sizeof(cplx)));
  TN_CUDA(cudaMallocHost((void**)&c.hscal, 64 * sizeof(cplx)));
  TN_CUDA(cudaMalloc((void**)&c.partials, 4 * (DOT_BLOCKS + 8) * sizeof(cplx)));
}
static void ctx_release(Ctx& c) {
  cudaSetDevice(c.device);
  if (c.stream) cudaStreamSynchronize(c.stream);
  svd_free(c.svd);
  svd_batched_free(c.svdb);
  for (auto& b : c.scratch) b.release();
  cudaFree(c.dscal); cudaFreeHost(c.hscal); cudaFree(c.partials);
  if (c.copy_stream) { cudaStreamDestroy(c.copy_stream); for (auto& ev : c.copy_ev) cudaEventDestroy(ev); }
  if (c.stream) cudaStreamDestroy(c.stream);
  c.stream = nullptr;
}

struct tn_ctx { Ctx c; };
struct tn_mps { Mps* m; };
struct tn_env { Env* e; };
struct tn_gates { Gates* g; };
struct tn_envsum { EnvSum* s; };
struct tn_imps { IMps* m; };
struct tn_shard { Shard* s; };


// Every handle remembers the device of its context; an entry point makes that device current for the calling thread before it
// touches the handle, so that two contexts on different GPUs can be used from one process / thread.
static inline void use_device_idx(int dev) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != dev) TN_CUDA(cudaSetDevice(dev));
}
static inline void use_device(tn_ctx* h) { use_device_idx(h->c.device); }
static inline void use_device(tn_mps* h) { use_device_idx(h->m->ctx->device); }
static inline void use_device(tn_env* h) { use_device_idx(h->e->ctx->device); }
[[maybe_unused]] static inline void use_device(tn_gates* h) { use_device_idx(h->g->ctx->device); }
static inline void use_device(tn_envsum* h) { use_device_idx(h->s->ctx->device); }
static inline void use_device(tn_imps* h) { use_device_idx(h->m->ctx->device); }

static thread_local std::string g_err;

template <class F>
static int32_t guard(F&& f) {
  try { f(); return TN_OK; }
  catch (const tn::Error& e) { g_err = e.what(); return e.code; }
  catch (const std::exception& e) { g_err = e.what(); return TN_ERR_INTERNAL; }
  catch (...) { g_err = "unknown error"; return TN_ERR_INTERNAL; }
}
static inline const cplx* C(const tn_cplx* p) { return reinterpret_cast<const cplx*>(p); }
static inline cplx* C(tn_cplx* p) { return reinterpret_cast<cplx*>(p); }
static inline Trunc T(tn_trunc_t t) { return Trunc{t.cutoff, (long long)t.maxdim, (long long)t.mindim}; }

extern "C" {

const char* tn_last_error(void) { return g_err.c_str(); }
int32_t tn_version(void) { return 100; }

int32_t tn_ctx_create(int32_t device, tn_ctx** out) {
  return guard([&] {
    TN_CHECK(out != nullptr, "null output pointer");
    auto* c = new tn_ctx();
    try { ctx_init(c->c, device); } catch (...) { delete c; throw; }
    *out = c;
  });
}
int32_t tn_ctx_destroy(tn_ctx* ctx) {
  return guard([&] {
    if (!ctx) return;
    ctx_release(ctx->c);
    delete ctx;
  });
}
int32_t tn_sync(tn_ctx* ctx) { return guard([&] { TN_CHECK(ctx, "tn_sync: null handle"); use_device(ctx); ctx->c.sync(); }); }
int32_t tn_ctx_stream(tn_ctx* ctx, void** s) { return guard([&] { TN_CHECK(ctx, "tn_ctx_stream: null handle"); use_device(ctx); *s = (void*)ctx->c.stream; }); }
int32_t tn_counters(tn_ctx* ctx, int64_t* launches, int64_t* matvecs, int64_t* svds) {
  return guard([&] { TN_CHECK(ctx, "tn_counters: null handle"); use_device(ctx);
    if (launches) *launches = zgemm_launch_count();
    if (matvecs) *matvecs = ctx->c.matvecs;
    if (svds) *svds = ctx->c.svds;
  });
}

// ---- MPS ---------------------------------------------------------------------------------------
int32_t tn_mps_upload(tn_ctx* ctx, int32_t rank, int32_t d, int32_t N, const int64_t* dims, const tn_cplx* const* site_ptrs,
                      int32_t center, tn_mps** out) {
  return guard([&] { TN_CHECK(ctx, "tn_mps_upload: null handle"); use_device(ctx);
    TN_CHECK(ctx && dims && site_ptrs && out, "null pointer");
    TN_CHECK(center >= 0 && center <= N, "center out of range");
    std::vector<long long> dd((size_t)N * (rank + 2));
    for (size_t i = 0; i < dd.size(); ++i) dd[i] = dims[i];
    auto* h = new tn_mps();
    try { h->m = mps_create(&ctx->c, rank, d, N, dd.data(), reinterpret_cast<const cplx* const*>(site_ptrs), center); } catch (...) { delete h; throw; }
    *out = h;
  });
}
int32_t tn_mps_free(tn_mps* m) { return guard([&] { if (m) { mps_free(m->m); delete m; } }); }
int32_t tn_mps_info(tn_mps* m, int32_t* rank, int32_t* d, int32_t* N, int32_t* center) {
  return guard([&] { TN_CHECK(m, "tn_mps_info: null handle"); use_device(m); if (rank) *rank = m->m->rank; if (d) *d = m->m->d; if (N) *N = m->m->N; if (center) *center = m->m->center; });
}
int32_t tn_mps_dims(tn_mps* m, int64_t* dims) {
  return guard([&] { TN_CHECK(m, "tn_mps_dims: null handle"); use_device(m);
    int r = m->m->rank + 2;
    for (int i = 0; i < m->m->N; ++i) for (int k = 0; k < r; ++k) dims[(size_t)i * r + k] = m->m->sites[i].dims[k];
  });
}
int32_t tn_mps_download_site(tn_mps* m, int32_t site, tn_cplx* out) { return guard([&] { TN_CHECK(m, "tn_mps_download_site: null handle"); use_device(m); mps_download_site(m->m, site, C(out)); }); }
int32_t tn_mps_upload_site(tn_mps* m, int32_t site, const int64_t* dims, const tn_cplx* data) {
  return guard([&] { TN_CHECK(m, "tn_mps_upload_site: null handle"); use_device(m); std::vector<long long> dd(dims, dims + m->m->rank + 2); mps_upload_site(m->m, site, dd.data(), C(data)); });
}
int32_t tn_mps_set_center(tn_mps* m, int32_t center) {
  return guard([&] { TN_CHECK(m, "tn_mps_set_center: null handle"); use_device(m); TN_CHECK(center >= 0 && center <= m->m->N, "center out of range"); m->m->center = center; });
}
int32_t tn_mps_maxbonddim(tn_mps* m, int64_t* out) { return guard([&] { TN_CHECK(m, "tn_mps_maxbonddim: null handle"); use_device(m); *out = m->m->maxbonddim(); }); }
int32_t tn_mps_norm(tn_mps* m, tn_cplx* out) { return guard([&] { TN_CHECK(m, "tn_mps_norm: null handle"); use_device(m); cplx v = mps_norm(m->m); out->re = v.x; out->im = v.y; }); }
int32_t tn_mps_normalize(tn_mps* m) { return guard([&] { TN_CHECK(m, "tn_mps_normalize: null handle"); use_device(m); mps_normalize(m->m); }); }
int32_t tn_mps_movecenter(tn_mps* m, int32_t idx, tn_trunc_t tr) { return guard([&] { TN_CHECK(m, "tn_mps_movecenter: null handle"); use_device(m); mps_movecenter(m->m, idx, T(tr)); }); }
int32_t tn_mps_replacesites(tn_mps* m, const tn_cplx* theta, int32_t site, int32_t direction, int32_t normalize, tn_trunc_t tr) {
  return guard([&] { TN_CHECK(m, "tn_mps_replacesites: null handle"); use_device(m);
    Mps* p = m->m; Ctx* c = p->ctx;
    TN_CHECK(site >= 1 && site + 1 <= p->N, "replacesites: site out of range");
    long long n = p->chiL(site) * p->phys() * p->phys() * p->chiR(site + 1);
    cplx* d = c->scratch[14].get((size_t)n, c->stream);
    TN_CUDA(cudaMemcpyAsync(d, theta, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
    mps_replacesites2(p, d, site, direction != 0, normalize != 0, T(tr));
    c->sync();
  });
}
int32_t tn_mps_applyop(tn_mps* m, int32_t site, const tn_cplx* op) {
  return guard([&] { TN_CHECK(m, "tn_mps_applyop: null handle"); use_device(m);
    Mps* p = m->m; Ctx* c = p->ctx;
    TN_CHECK(site >= 1 && site <= p->N, "site out of range");
    cplx* d = c->scratch[15].get((size_t)p->d * p->d, c->stream);
    TN_CUDA(cudaMemcpyAsync(d, op, (size_t)p->d * p->d * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
    c->sync();
    mps_applyop1(p, site, d);
  });
}
int32_t tn_mps_bond_spectrum(tn_mps* m, int32_t site, double* out, int64_t cap, int64_t* k_out) {
  return guard([&] { TN_CHECK(m, "tn_mps_bond_spectrum: null handle"); use_device(m);
    std::vector<double> s; mps_bond_spectrum(m->m, site, s);
    TN_CHECK((int64_t)s.size() <= cap, "spectrum buffer too small");
    std::memcpy(out, s.data(), s.size() * sizeof(double));
    *k_out = (int64_t)s.size();
  });
}
int32_t tn_mps_copy(tn_mps* m, tn_mps** out) {      // deepcopy(psi): abstractmps.jl:95-96
  return guard([&] {
    TN_CHECK(m && out, "tn_mps_copy: null pointer");
    Mps* src = m->m; Ctx* c = src->ctx;
    auto cp = std::make_unique<Mps>();
    cp->ctx = c; cp->rank = src->rank; cp->d = src->d; cp->N = src->N; cp->center = src->center;
    cp->sites.resize(src->N);
    for (int i = 0; i < src->N; ++i) {
      c->alloc(cp->sites[i], src->sites[i].dims);
      TN_CUDA(cudaMemcpyAsync(cp->sites[i].p, src->sites[i].p, (size_t)src->sites[i].size() * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->sync();
    auto* h = new tn_mps();
    h->m = cp.release();
    *out = h;
  });
}
int32_t tn_mps_scale(tn_mps* m, tn_cplx a) {        // psi * a in place: the centre tensor (site 1 if unset) is multiplied, abstractmps.jl:99-109
  return guard([&] {
    TN_CHECK(m, "tn_mps_scale: null handle"); use_device(m);
    Mps* p = m->m;
    Tensor& t = p->sites[(p->center != 0 ? p->center : 1) - 1];
    zscal(t.size(), cplx{a.re, a.im}, t.p, p->ctx->stream);
    p->ctx->sync();
  });
}
int32_t tn_mpo_apply(tn_mps* O, tn_mps* psi, tn_trunc_t tr, tn_mps** out) {
  return guard([&] {
    TN_CHECK(O && psi && out, "tn_mpo_apply: null pointer");
    auto* h = new tn_mps();
    try { h->m = mpo_apply(O->m, psi->m, T(tr)); } catch (...) { delete h; throw; }
    psi->m->ctx->sync();
    *out = h;
  });
}
int32_t tn_mpo_compress(tn_mps* m, tn_trunc_t tr) { return guard([&] { TN_CHECK(m, "tn_mpo_compress: null handle"); use_device(m); TN_CHECK(m != nullptr, "null pointer"); mpo_compress(m->m, T(tr)); m->m->ctx->sync(); }); }
int32_t tn_expect_local(tn_mps* m, int32_t nops, const int32_t* sites, const tn_cplx* ops, tn_cplx* out) {
  return guard([&] { TN_CHECK(m, "tn_expect_local: null handle"); use_device(m); expect_local(m->m, nops, sites, C(ops), C(out)); });
}

// ---- SVD ---------------------------------------------------------------------------------------
int32_t tn_svd_trunc(tn_ctx* ctx, const tn_cplx* mat, int64_t m, int64_t n, tn_trunc_t tr, tn_cplx* U, double* S, tn_cplx* Vh,
                     int64_t* k_out, int32_t* sweeps_out) {
  return guard([&] { TN_CHECK(ctx, "tn_svd_trunc: null handle"); use_device(ctx);
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    TN_CHECK(m >= 1 && n >= 1, "svd: empty matrix");
    cplx* dM = c->scratch[0].get((size_t)(m * n), s);
    TN_CUDA(cudaMemcpyAsync(dM, mat, (size_t)(m * n) * sizeof(cplx), cudaMemcpyHostToDevice, s));
    int k = svd_factor(c->svd, dM, (int)m, (int)n, m, T(tr), s); c->svds++;
    cplx* dU = c->scratch[1].get((size_t)(m * k), s);
    cplx* dV = c->scratch[2].get((size_t)(k * n), s);
    svd_gather_U(c->svd, dU, m, false, s);
    svd_gather_Vh(c->svd, dV, k, false, s);
    TN_CUDA(cudaMemcpyAsync(U, dU, (size_t)(m * k) * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(Vh, dV, (size_t)(k * n) * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(S, c->svd.sig, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s));
    c->sync();
    *k_out = k;
    if (sweeps_out) *sweeps_out = c->svd.sweeps;
  });
}

// The split the MPS drivers use (gmps.jl:60-82, 218-256): one orthonormal factor and the other one multiplied by S.
// side = 1: U (m x k) orthonormal, Vh <- S V^H;  side = 2: Vh (k x n) orthonormal, U <- U S.  repeat > 1 re-runs the
// factorisation on the resident matrix (timing without the PCIe copies); *ms_out (optional) = device time of the last run.
int32_t tn_svd_trunc_split(tn_ctx* ctx, const tn_cplx* mat, int64_t m, int64_t n, tn_trunc_t tr, int32_t side, tn_cplx* U, double* S,
                           tn_cplx* Vh, int64_t* k_out, int32_t* sweeps_out, int32_t repeat, double* ms_out) {
  return guard([&] { TN_CHECK(ctx, "tn_svd_trunc_split: null handle"); use_device(ctx);
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    TN_CHECK(m >= 1 && n >= 1 && (side == 1 || side == 2) && mat && U && S && Vh && k_out, "svd split: bad arguments");
    cplx* dM = c->scratch[0].get((size_t)(m * n), s);
    TN_CUDA(cudaMemcpyAsync(dM, mat, (size_t)(m * n) * sizeof(cplx), cudaMemcpyHostToDevice, s));
    const size_t kmax = (size_t)std::min(m, n);
    cplx* dU = c->scratch[1].get((size_t)m * kmax, s);
    cplx* dV = c->scratch[2].get(kmax * (size_t)n, s);
    cudaEvent_t e0, e1; TN_CUDA(cudaEventCreate(&e0)); TN_CUDA(cudaEventCreate(&e1));
    int k = 0;
    for (int r = 0; r < std::max(1, (int)repeat); ++r) {
      TN_CUDA(cudaEventRecord(e0, s));
      k = svd_factor(c->svd, dM, (int)m, (int)n, m, T(tr), s, side); c->svds++;
      svd_gather_U(c->svd, dU, m, side == 2, s);
      svd_gather_Vh(c->svd, dV, k, side == 1, s);
      TN_CUDA(cudaEventRecord(e1, s));
    }
    TN_CUDA(cudaMemcpyAsync(U, dU, (size_t)(m * k) * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(Vh, dV, (size_t)(k * n) * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(S, c->svd.sig, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s));
    c->sync();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms;
    *k_out = k;
    if (sweeps_out) *sweeps_out = c->svd.sweeps;
  });
}

int32_t tn_svd_trunc_batched(tn_ctx* ctx, int32_t B, const tn_cplx* mats, int64_t m, int64_t n, tn_trunc_t tr, tn_cplx* U, double* S,
                             tn_cplx* Vh, int64_t* k_out, int32_t* sweeps_out) {
  return guard([&] { TN_CHECK(ctx, "tn_svd_trunc_batched: null handle"); use_device(ctx);
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    TN_CHECK(B >= 1 && m >= 1 && n >= 1 && mats && U && S && Vh && k_out, "batched svd: bad arguments");
    const size_t mn = (size_t)(m * n), kmax = (size_t)std::min(m, n);
    cplx* dM = c->scratch[0].get((size_t)B * mn, s);
    TN_CUDA(cudaMemcpyAsync(dM, mats, (size_t)B * mn * sizeof(cplx), cudaMemcpyHostToDevice, s));
    std::vector<const cplx*> ptrs(B);
    for (int b = 0; b < B; ++b) ptrs[b] = dM + (size_t)b * mn;
    svd_batched_factor(c->svdb, B, ptrs.data(), (int)m, (int)n, m, T(tr), s); c->svds += B;
    cplx* dU = c->scratch[1].get((size_t)m * kmax, s);
    cplx* dV = c->scratch[2].get(kmax * (size_t)n, s);
    for (int b = 0; b < B; ++b) {
      const int k = c->svdb.k[b];
      svd_batched_gather_U(c->svdb, b, dU, m, false, s);
      svd_batched_gather_Vh(c->svdb, b, dV, k, false, s);
      TN_CUDA(cudaMemcpyAsync(U + (size_t)b * m * kmax, dU, (size_t)m * k * sizeof(cplx), cudaMemcpyDeviceToHost, s));
      TN_CUDA(cudaMemcpyAsync(Vh + (size_t)b * kmax * n, dV, (size_t)k * n * sizeof(cplx), cudaMemcpyDeviceToHost, s));
      TN_CUDA(cudaMemcpyAsync(S + (size_t)b * kmax, c->svdb.sig + (size_t)b * c->svdb.npad, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s));
      c->sync();                      // dU / dV are reused by the next problem
      k_out[b] = k;
    }
    if (sweeps_out) *sweeps_out = c->svdb.sweeps;
  });
}

int32_t tn_svd_set_precond(int32_t mode) { return guard([&] { svd_set_precond(mode); }); }

int32_t tn_jacobi_qr_update_pass(tn_ctx* ctx, const tn_cplx* Q, int64_t rows, int64_t ncols, int32_t panel, tn_cplx* C_out, tn_cplx* Q_out) {
  return guard([&] { TN_CHECK(ctx && Q && C_out && Q_out, "tn_jacobi_qr_update_pass: null pointer"); use_device(ctx);
    TN_CHECK(rows >= 1 && rows < (1ll << 30) && ncols >= 128 && ncols % 64 == 0 && ncols < (1ll << 20), "tn_jacobi_qr_update_pass: bad shape");
    const int npanels = (int)(ncols / 64), ntiles = npanels - panel - 1;
    TN_CHECK(panel >= 0 && ntiles >= 1, "tn_jacobi_qr_update_pass: the panel needs at least one trailing tile");
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    const size_t qe = (size_t)rows * ncols, ce = (size_t)64 * 64 * ntiles;
    cplx* dQ = c->scratch[0].get(qe, s);
    cplx* dC = c->scratch[1].get(ce, s);
    int* dI = reinterpret_cast<int*>(c->scratch[3].get((size_t)npanels, s));      // 16 B per entry: 2 ints per panel fit twice over
    std::vector<int> tab((size_t)2 * npanels);
    for (int i = 0; i < 2 * npanels; ++i) tab[i] = i;                              // panel p = column blocks (2p, 2p + 1)
    TN_CUDA(cudaMemcpyAsync(dQ, Q, qe * sizeof(cplx), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemcpyAsync(dI, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemsetAsync(dC, 0, ce * sizeof(cplx), s));
    jacobi_cross64(dQ, rows, dI + 2 * panel, dQ, rows, (int)rows, dI + 2 * (panel + 1), 1, ntiles, dC, 64, 0, 16, s);
    jacobi_update64(dQ, rows, dI + 2 * panel, dQ, rows, (int)rows, dI + 2 * (panel + 1), 1, ntiles, dC, 64, 0, s);
    TN_CUDA(cudaMemcpyAsync(C_out, dC, ce * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(Q_out, dQ, qe * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
  });
}
int32_t tn_svd_split_schedule(int32_t nblocks, int32_t groups, int32_t* out5, int64_t capacity, int64_t* npairs_out) {
  return guard([&] { TN_CHECK(out5 && npairs_out && nblocks >= 2 && nblocks <= 4096 && groups >= 1 && groups <= 4, "tn_svd_split_schedule: bad argument");
    const long long n = svd_split_schedule_dump(nblocks, groups, out5, capacity);
    TN_CHECK(n >= 0, "tn_svd_split_schedule: output buffer too small");
    *npairs_out = n;
  });
}
int32_t tn_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
  return guard([&] { TN_CHECK(ctr4 && key2 && out4, "tn_philox4x32_10: null pointer"); philox4x32_10(ctr4, key2, out4); });
}
int32_t tn_qjmc_uniform(uint64_t seed, uint64_t trajectory, uint64_t step, uint64_t slot, double* out) {
  return guard([&] { TN_CHECK(out, "tn_qjmc_uniform: null pointer"); *out = counter_uniform(seed, trajectory, step, slot); });
}

int32_t tn_jacobi_pair_pass(tn_ctx* ctx, const tn_cplx* Z, int64_t rows, int64_t ncols, const int32_t* pairs, int32_t npairs,
                            const tn_cplx* J, const int32_t* skip, tn_cplx* G_out, tn_cplx* Z_out) {
  return guard([&] { TN_CHECK(ctx && Z && pairs && J && G_out && Z_out, "tn_jacobi_pair_pass: null pointer"); use_device(ctx);
    TN_CHECK(rows >= 1 && rows < (1ll << 30) && ncols >= 64 && ncols % 32 == 0 && ncols < (1ll << 24) && npairs >= 1, "tn_jacobi_pair_pass: bad shape");
    const int nb = (int)(ncols / 32);
    std::vector<char> seen((size_t)nb, 0);
    for (int i = 0; i < 2 * npairs; ++i) {
      TN_CHECK(pairs[i] >= 0 && pairs[i] < nb && !seen[pairs[i]], "tn_jacobi_pair_pass: column blocks must be in range and disjoint");
      seen[pairs[i]] = 1;
    }
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    const size_t ze = (size_t)rows * ncols, je = (size_t)npairs * 64 * 64;
    cplx* dZ = c->scratch[0].get(ze, s);
    cplx* dJ = c->scratch[1].get(je, s);
    cplx* dG = c->scratch[2].get(je, s);
    int* dI = reinterpret_cast<int*>(c->scratch[3].get((size_t)npairs, s));      // 16 B per entry: room for 2 + 1 ints per pair
    int* dskip = dI + 2 * npairs;
    TN_CUDA(cudaMemcpyAsync(dZ, Z, ze * sizeof(cplx), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemcpyAsync(dJ, J, je * sizeof(cplx), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemcpyAsync(dI, pairs, (size_t)2 * npairs * sizeof(int), cudaMemcpyHostToDevice, s));
    if (skip) TN_CUDA(cudaMemcpyAsync(dskip, skip, (size_t)npairs * sizeof(int), cudaMemcpyHostToDevice, s));
    jacobi_gram64(dZ, rows, (int)rows, dI, npairs, dG, 32, s);
    jacobi_rot64(dZ, rows, (int)rows, dI, npairs, dJ, skip ? dskip : nullptr, s);
    TN_CUDA(cudaMemcpyAsync(G_out, dG, je * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    TN_CUDA(cudaMemcpyAsync(Z_out, dZ, ze * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
  });
}

static Idx2 I2(tn_idx2_t i) { return Idx2{(int)std::min<int64_t>(i.n0, 0x7fffffff), (long long)i.s0, (long long)i.s1, nullptr, 0}; }

int32_t tn_contract_strided(tn_ctx* ctx, int64_t M, int64_t N, int64_t K, const tn_cplx* A, int64_t ae, tn_idx2_t am, tn_idx2_t ak, int32_t conjA,
                            const tn_cplx* B, int64_t be, tn_idx2_t bk, tn_idx2_t bn, int32_t conjB, tn_cplx* Cc, int64_t ce,
                            tn_idx2_t cm, tn_idx2_t cn, tn_cplx alpha) {
  return guard([&] { TN_CHECK(ctx, "tn_contract_strided: null handle"); use_device(ctx);
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    cplx* dA = c->scratch[0].get((size_t)ae, s);
    cplx* dB = c->scratch[1].get((size_t)be, s);
    cplx* dC = c->scratch[2].get((size_t)ce, s);
    TN_CUDA(cudaMemcpyAsync(dA, A, (size_t)ae * sizeof(cplx), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemcpyAsync(dB, B, (size_t)be * sizeof(cplx), cudaMemcpyHostToDevice, s));
    TN_CUDA(cudaMemsetAsync(dC, 0, (size_t)ce * sizeof(cplx), s));
    GemmDesc g{};
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.A = dA; g.am = I2(am); g.ak = I2(ak); g.conjA = conjA;
    g.B = dB; g.bk = I2(bk); g.bn = I2(bn); g.conjB = conjB;
    g.C = dC; g.cm = I2(cm); g.cn = I2(cn);
    g.alpha = cplx{alpha.re, alpha.im}; g.beta = cplx{0, 0};
    g.batch = 1; g.ksplit = 1; g.kchunk = (int)K;
    zgemm_auto(g, s);
    TN_CUDA(cudaMemcpyAsync(Cc, dC, (size_t)ce * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
  });
}

int32_t tn_contract_strided_dev(tn_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* A, tn_idx2_t am, tn_idx2_t ak, int32_t conjA,
                                const void* B, tn_idx2_t bk, tn_idx2_t bn, int32_t conjB, void* Cc, tn_idx2_t cm, tn_idx2_t cn,
                                tn_cplx alpha, tn_cplx beta) {
  return guard([&] { TN_CHECK(ctx, "tn_contract_strided_dev: null handle"); use_device(ctx);
    TN_CHECK(M >= 0 && N >= 0 && K >= 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "contract: extent out of range");
    GemmDesc g{};
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.A = (const cplx*)A; g.am = I2(am); g.ak = I2(ak); g.conjA = conjA;
    g.B = (const cplx*)B; g.bk = I2(bk); g.bn = I2(bn); g.conjB = conjB;
    g.C = (cplx*)Cc; g.cm = I2(cm); g.cn = I2(cn);
    g.alpha = cplx{alpha.re, alpha.im}; g.beta = cplx{beta.re, beta.im};
    g.batch = 1; g.ksplit = 1; g.kchunk = (int)K;
    zgemm_auto(g, ctx->c.stream);
  });
}

// ---- environments ------------------------------------------------------------------------------
int32_t tn_env_create(tn_ctx* ctx, tn_mps* bra, tn_mps* mpo, tn_mps* ket, tn_cplx coeff, int32_t center, tn_env** out) {
  return guard([&] { TN_CHECK(ctx && bra && ket, "tn_env_create: null handle"); use_device(ctx);
    TN_CHECK(ctx && bra && ket && out, "null pointer");
    auto* h = new tn_env();
    try { h->e = env_create(&ctx->c, bra->m, mpo ? mpo->m : nullptr, ket->m, cplx{coeff.re, coeff.im}, center); } catch (...) { delete h; throw; }
    *out = h;
  });
}
int32_t tn_env_free(tn_env* e) { return guard([&] { if (e) { env_free(e->e); delete e; } }); }
int32_t tn_env_buildleft(tn_env* e, int32_t idx) { return guard([&] { TN_CHECK(e, "tn_env_buildleft: null handle"); use_device(e); env_buildleft(e->e, idx); }); }
int32_t tn_env_buildright(tn_env* e, int32_t idx) { return guard([&] { TN_CHECK(e, "tn_env_buildright: null handle"); use_device(e); env_buildright(e->e, idx); }); }
int32_t tn_env_movecenter(tn_env* e, int32_t idx) { return guard([&] { TN_CHECK(e, "tn_env_movecenter: null handle"); use_device(e); env_movecenter(e->e, idx); }); }
int32_t tn_env_center(tn_env* e, int32_t* out) { return guard([&] { TN_CHECK(e, "tn_env_center: null handle"); use_device(e); *out = e->e->center; }); }
int32_t tn_env_block_dims(tn_env* e, int32_t idx, int64_t* dims3) {
  return guard([&] { TN_CHECK(e, "tn_env_block_dims: null handle"); use_device(e); const Tensor& t = env_block(e->e, idx); for (int k = 0; k < 3; ++k) dims3[k] = t.dims[k]; });
}
int32_t tn_env_block_download(tn_env* e, int32_t idx, tn_cplx* out) {
  return guard([&] { TN_CHECK(e, "tn_env_block_download: null handle"); use_device(e);
    const Tensor& t = env_block(e->e, idx);
    TN_CUDA(cudaMemcpyAsync(out, t.p, (size_t)t.size() * sizeof(cplx), cudaMemcpyDeviceToHost, e->e->ctx->stream));
    e->e->ctx->sync();
  });
}
int32_t tn_env_block_upload(tn_env* eh, int32_t idx, const int64_t* dims3, const tn_cplx* data) {
  return guard([&] { TN_CHECK(eh, "tn_env_block_upload: null handle"); use_device(eh);
    Env* e = eh->e; Ctx* c = e->ctx;
    TN_CHECK(idx >= 1 && idx <= e->ket->N, "block index out of range");
    Tensor& t = e->blocks[idx - 1];
    c->alloc(t, {(long long)dims3[0], (long long)dims3[1], (long long)dims3[2]});
    TN_CUDA(cudaMemcpyAsync(t.p, data, (size_t)t.size() * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
    c->sync();
  });
}
int32_t tn_env_set_center(tn_env* eh, int32_t center) {
  return guard([&] { TN_CHECK(eh, "tn_env_set_center: null handle"); use_device(eh); TN_CHECK(center >= 0 && center <= eh->e->ket->N, "center out of range"); eh->e->center = center; });
}
static int product_site(Env* e, int direction) {   // projmps.jl:109
  TN_CHECK(e->center >= 1, "the environment centre is not set");
  int site = direction ? e->center - 1 : e->center;
  TN_CHECK(site >= 1 && site + 1 <= e->ket->N, "product: the two sites fall outside the chain");
  return site;
}
int32_t tn_env_product(tn_env* eh, const tn_cplx* theta, int32_t direction, tn_cplx* out) {
  return guard([&] { TN_CHECK(eh, "tn_env_product: null handle"); use_device(eh);
    Env* e = eh->e;
    int site = product_site(e, direction);
    env_product_host(e, C(theta), site, C(out));
  });
}
int32_t tn_env_product_dev(tn_env* eh, const void* theta_dev, int32_t direction, void* out_dev, int32_t reps) {
  return guard([&] { TN_CHECK(eh, "tn_env_product_dev: null handle"); use_device(eh);
    Env* e = eh->e;
    int site = product_site(e, direction);
    for (int r = 0; r < std::max(1, reps); ++r) env_product_dev(e, (const cplx*)theta_dev, site, (cplx*)out_dev);
  });
}
int32_t tn_env_product_profile(tn_env* eh, const void* theta_dev, int32_t direction, void* out_dev, int32_t reps, double* stage_ms3) {
  return guard([&] { TN_CHECK(eh, "tn_env_product_profile: null handle"); use_device(eh);
    Env* e = eh->e; Ctx* c = e->ctx;
    int site = product_site(e, direction);
    reps = std::max(1, reps);
    std::vector<cudaEvent_t> ev((size_t)4 * reps);
    for (auto& x : ev) TN_CUDA(cudaEventCreate(&x));
    for (int r = 0; r < reps; ++r) env_product_dev(e, (const cplx*)theta_dev, site, (cplx*)out_dev, &ev[(size_t)4 * r]);
    c->sync();
    for (int k = 0; k < 3; ++k) stage_ms3[k] = 0;
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 3; ++k) { float ms = 0; TN_CUDA(cudaEventElapsedTime(&ms, ev[(size_t)4 * r + k], ev[(size_t)4 * r + k + 1])); stage_ms3[k] += ms; }
    for (auto& x : ev) cudaEventDestroy(x);
  });
}
int32_t tn_env_calculate(tn_env* e, tn_cplx* out) { return guard([&] { TN_CHECK(e, "tn_env_calculate: null handle"); use_device(e); cplx v = env_calculate(e->e); out->re = v.x; out->im = v.y; }); }

// ---- drivers -----------------------------------------------------------------------------------
int32_t tn_dmrg_sweep(tn_mps* psi, tn_env* env, int32_t direction, tn_lanczos_t lz, tn_trunc_t tr, double* energy, int64_t* maxbond) {
  return guard([&] { TN_CHECK(psi && env, "tn_dmrg_sweep: null handle"); use_device(psi);
    long long mb = 0;
    dmrg_halfsweep(psi->m, env->e, direction != 0, Lanczos{lz.krylovdim, lz.maxiter, lz.tol}, T(tr), energy, &mb);
    psi->m->ctx->sync();
    if (maxbond) *maxbond = mb;
  });
}
int32_t tn_eigsolve(tn_env* eh, const tn_cplx* theta0, int32_t direction, tn_lanczos_t lz, double* eig, tn_cplx* theta_out, int32_t* numops) {
  return guard([&] { TN_CHECK(eh, "tn_eigsolve: null handle"); use_device(eh);
    Env* e = eh->e; Ctx* c = e->ctx; cudaStream_t s = c->stream;
    int site = product_site(e, direction);
    long long n = e->ket->chiL(site) * e->ket->d * e->ket->d * e->ket->chiR(site + 1);
    cplx* din = c->scratch[13].get((size_t)n, s);
    cplx* dout = c->scratch[14].get((size_t)n, s);
    TN_CUDA(cudaMemcpyAsync(din, theta0, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, s));
    int ops = 0;
    double v = lanczos_lowest(e, site, din, dout, n, Lanczos{lz.krylovdim, lz.maxiter, lz.tol}, &ops);
    TN_CUDA(cudaMemcpyAsync(theta_out, dout, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
    *eig = v;
    if (numops) *numops = ops;
  });
}
int32_t tn_gates_upload(tn_ctx* ctx, int32_t d, int32_t nrows, const int32_t* counts, const int32_t* sites, const int32_t* nsites,
                        const tn_cplx* const* gate_ptrs, tn_gates** out) {
  return guard([&] { TN_CHECK(ctx, "tn_gates_upload: null handle"); use_device(ctx);
    auto* h = new tn_gates();
    try { h->g = gates_create(&ctx->c, d, nrows, counts, sites, nsites, reinterpret_cast<const cplx* const*>(gate_ptrs)); } catch (...) { delete h; throw; }
    *out = h;
  });
}
int32_t tn_gates_free(tn_gates* g) { return guard([&] { if (g) { gates_free(g->g); delete g; } }); }
int32_t tn_apply_gates(tn_mps* psi, tn_gates* gates, tn_trunc_t tr) {
  return guard([&] { TN_CHECK(psi && gates, "tn_apply_gates: null handle"); use_device(psi); apply_gates(psi->m, gates->g, T(tr)); psi->m->ctx->sync(); });
}
int32_t tn_apply_gates_fidelity(tn_mps* psi, tn_gates* gates, tn_trunc_t tr, double* fidelity_out) {
  return guard([&] { TN_CHECK(psi && gates, "tn_apply_gates_fidelity: null handle"); use_device(psi);
    TN_CHECK(psi && gates && fidelity_out, "null pointer");
    apply_gates(psi->m, gates->g, T(tr), fidelity_out);
    psi->m->ctx->sync();
  });
}
int32_t tn_qjmc_run(tn_mps* psi, tn_gates* gates, int32_t njump, const int32_t* jump_sites, const tn_cplx* jump_ops,
                    const double* jump_coeffs, int32_t steps, double dt, tn_trunc_t tr, const double* uniforms, uint64_t seed,
                    uint64_t trajectory, const tn_cplx* obs_op, int32_t save_every, tn_cplx* obs_out, int32_t* jumps_out,
                    double* jumptimes_out, int32_t jump_cap, int32_t* njumps_out, int32_t classical) {
  return guard([&] { TN_CHECK(psi && gates, "tn_qjmc_run: null handle"); use_device(psi);
    int nj = qjmc_run(psi->m, gates->g, njump, jump_sites, C(jump_ops), jump_coeffs, steps, dt, T(tr), uniforms, seed, trajectory,
                      obs_op ? C(obs_op) : nullptr, save_every, obs_out ? C(obs_out) : nullptr, jumps_out, jumptimes_out, jump_cap,
                      classical != 0);
    if (njumps_out) *njumps_out = nj;
  });
}

// ---- projector sums / squared projectors / project / one-site branch / vmps (tn_projsum.cu) ----------------------
int32_t tn_env_create_squared(tn_ctx* ctx, tn_mps* V, tn_mps* psi, tn_cplx coeff, int32_t center, tn_env** out) {
  return guard([&] { TN_CHECK(ctx && V && psi, "tn_env_create_squared: null handle"); use_device(ctx);
    TN_CHECK(ctx && V && psi && out, "null pointer");
    auto* h = new tn_env();
    try { h->e = env_create_squared(&ctx->c, V->m, psi->m, cplx{coeff.re, coeff.im}, center); } catch (...) { delete h; throw; }
    *out = h;
  });
}
static int nsite_site(int center, int N, int direction, int nsites) {   // projmps.jl:109, :155
  TN_CHECK(nsites == 1 || nsites == 2, "nsites must be 1 or 2");
  TN_CHECK(center >= 1, "the environment centre is not set");
  int site = direction ? center - nsites + 1 : center;
  TN_CHECK(site >= 1 && site + nsites - 1 <= N, "the sites fall outside the chain");
  return site;
}
static long long nsite_size(Mps* ket, int site, int nsites) {
  long long n = ket->chiL(site) * ket->chiR(site + nsites - 1);
  for (int i = 0; i < nsites; ++i) n *= ket->d;
  return n;
}
static void envsum_product_host(EnvSum* es, const tn_cplx* A, int direction, int nsites, tn_cplx* out) {
  Ctx* c = es->ctx; cudaStream_t s = c->stream;
  Mps* ket = es->projs[0]->ket;
  int site = nsite_site(es->center, ket->N, direction, nsites);
  long long n = nsite_size(ket, site, nsites);
  cplx* din = c->scratch[13].get((size_t)n, s);
  cplx* dout = c->scratch[14].get((size_t)n, s);
  TN_CUDA(cudaMemcpyAsync(din, A, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, s));
  envsum_prepare(es, site, nsites);
  envsum_apply(es, din, site, nsites, dout);
  TN_CUDA(cudaMemcpyAsync(out, dout, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToHost, s));
  c->sync();
}
static void envsum_project_host(EnvSum* es, int direction, int nsites, tn_cplx* out) {
  Ctx* c = es->ctx; cudaStream_t s = c->stream;
  Mps* ket = es->projs[0]->ket;
  int site = nsite_site(es->center, ket->N, direction, nsites);
  long long n = nsite_size(ket, site, nsites);
  cplx* d = c->scratch[14].get((size_t)n, s);
  envsum_project_phi(es, site, nsites, d);
  zconj_inplace(n, d, s);                    // project() itself, not its conjugate
  TN_CUDA(cudaMemcpyAsync(out, d, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToHost, s));
  c->sync();
}
int32_t tn_env_product_n(tn_env* eh, const tn_cplx* A, int32_t direction, int32_t nsites, tn_cplx* out) {
  return guard([&] { TN_CHECK(eh, "tn_env_product_n: null handle"); use_device(eh);
    TN_CHECK(eh && A && out, "null pointer");
    EnvSum one{eh->e->ctx, {eh->e}, eh->e->center};
    envsum_product_host(&one, A, direction, nsites, out);
  });
}
int32_t tn_env_project(tn_env* eh, int32_t direction, int32_t nsites, tn_cplx* out) {
  return guard([&] { TN_CHECK(eh, "tn_env_project: null handle"); use_device(eh);
    TN_CHECK(eh && out, "null pointer");
    EnvSum one{eh->e->ctx, {eh->e}, eh->e->center};
    envsum_project_host(&one, direction, nsites, out);
  });
}
int32_t tn_envsum_create(tn_ctx* ctx, int32_t n, tn_env* const* envs, int32_t center, tn_envsum** out) {
  return guard([&] { TN_CHECK(ctx, "tn_envsum_create: null handle"); use_device(ctx);
    TN_CHECK(ctx && envs && out && n >= 1, "ProjMPSSum: bad arguments");
    std::vector<Env*> v;
    for (int i = 0; i < n; ++i) { TN_CHECK(envs[i] != nullptr, "ProjMPSSum: null projection"); v.push_back(envs[i]->e); }
    auto* h = new tn_envsum();
    try { h->s = envsum_create(&ctx->c, n, v.data(), center); } catch (...) { delete h; throw; }
    *out = h;
  });
}
int32_t tn_envsum_free(tn_envsum* s) { return guard([&] { if (s) { envsum_free(s->s); delete s; } }); }
int32_t tn_envsum_movecenter(tn_envsum* s, int32_t idx) { return guard([&] { TN_CHECK(s, "tn_envsum_movecenter: null handle"); use_device(s); envsum_movecenter(s->s, idx); }); }
int32_t tn_envsum_calculate(tn_envsum* s, tn_cplx* out) { return guard([&] { TN_CHECK(s, "tn_envsum_calculate: null handle"); use_device(s); cplx v = envsum_calculate(s->s); out->re = v.x; out->im = v.y; }); }
int32_t tn_envsum_product(tn_envsum* s, const tn_cplx* A, int32_t direction, int32_t nsites, tn_cplx* out) {
  return guard([&] { TN_CHECK(s, "tn_envsum_product: null handle"); use_device(s); TN_CHECK(s && A && out, "null pointer"); envsum_product_host(s->s, A, direction, nsites, out); });
}
int32_t tn_envsum_project(tn_envsum* s, int32_t direction, int32_t nsites, tn_cplx* out) {
  return guard([&] { TN_CHECK(s, "tn_envsum_project: null handle"); use_device(s); TN_CHECK(s && out, "null pointer"); envsum_project_host(s->s, direction, nsites, out); });
}
int32_t tn_dmrg_sweep_sum(tn_mps* psi, tn_envsum* Hs, int32_t direction, int32_t nsites, tn_lanczos_t lz, tn_trunc_t tr,
                          double* energy, int64_t* maxbond) {
  return guard([&] { TN_CHECK(psi && Hs, "tn_dmrg_sweep_sum: null handle"); use_device(psi);
    TN_CHECK(psi && Hs, "null pointer");
    long long mb = 0;
    dmrg_halfsweep_sum(psi->m, Hs->s, direction != 0, nsites, Lanczos{lz.krylovdim, lz.maxiter, lz.tol}, T(tr), energy, &mb);
    psi->m->ctx->sync();
    if (maxbond) *maxbond = mb;
  });
}
int32_t tn_vmps_sweep(tn_mps* psi, tn_envsum* Vs, int32_t direction, int32_t nsites, tn_trunc_t tr, int64_t* maxbond) {
  return guard([&] { TN_CHECK(psi && Vs, "tn_vmps_sweep: null handle"); use_device(psi);
    TN_CHECK(psi && Vs, "null pointer");
    long long mb = 0;
    vmps_halfsweep(psi->m, Vs->s, direction != 0, nsites, T(tr), &mb);
    psi->m->ctx->sync();
    if (maxbond) *maxbond = mb;
  });
}

// ---- device-pointer entry points for callers that orchestrate the sweep themselves (multi-GPU sharded DMRG) -------
int32_t tn_mps_site_ptr(tn_mps* m, int32_t site, void** dev_out) {
  return guard([&] { TN_CHECK(m, "tn_mps_site_ptr: null handle"); use_device(m);
    TN_CHECK(m && dev_out, "null pointer");
    TN_CHECK(site >= 1 && site <= m->m->N, "site index out of range");
    m->m->ctx->sync();
    *dev_out = (void*)m->m->sites[site - 1].p;
  });
}
int32_t tn_mps_replacesites_dev(tn_mps* m, const void* theta_dev, int32_t site, int32_t direction, int32_t normalize, tn_trunc_t tr) {
  return guard([&] { TN_CHECK(m, "tn_mps_replacesites_dev: null handle"); use_device(m);
    TN_CHECK(m && theta_dev, "null pointer");
    mps_replacesites2(m->m, (const cplx*)theta_dev, site, direction != 0, normalize != 0, T(tr));
    m->m->ctx->sync();
  });
}
int32_t tn_mps_upload_site_dev(tn_mps* m, int32_t site, const int64_t* dims, const void* data_dev) {
  return guard([&] { TN_CHECK(m, "tn_mps_upload_site_dev: null handle"); use_device(m);
    Mps* p = m->m; Ctx* c = p->ctx;
    TN_CHECK(site >= 1 && site <= p->N, "site index out of range");
    std::vector<long long> dd(dims, dims + p->rank + 2);
    c->alloc(p->sites[site - 1], dd);
    TN_CUDA(cudaMemcpyAsync(p->sites[site - 1].p, data_dev, (size_t)p->sites[site - 1].size() * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
    c->sync();
  });
}
int32_t tn_memcpy_dev(tn_ctx* ctx, void* dst_dev, const void* src_dev, int64_t nbytes) {
  return guard([&] { TN_CHECK(ctx, "tn_memcpy_dev: null handle"); use_device(ctx);
    TN_CHECK(ctx && dst_dev && src_dev && nbytes >= 0, "memcpy: bad arguments");
    TN_CUDA(cudaMemcpyAsync(dst_dev, src_dev, (size_t)nbytes, cudaMemcpyDeviceToDevice, ctx->c.stream));
    ctx->c.sync();
  });
}
// ---- truncated SVD in three calls with a caller-owned pair schedule (distributed Jacobi sweeps) ---------------------------------
int32_t tn_svd_dist_begin(tn_ctx* ctx, const void* mat_dev, int64_t m, int64_t n, void** Z_dev, int64_t* ldz, int64_t* zrows,
                          int32_t* nblocks, double* tol) {
  return guard([&] {
    TN_CHECK(ctx && mat_dev && Z_dev && ldz && zrows && nblocks && tol, "tn_svd_dist_begin: null pointer");
    Ctx* c = &ctx->c;
    const int nb = svd_dist_begin(c->svd, (const cplx*)mat_dev, (int)m, (int)n, m, c->stream); c->svds++;
    *Z_dev = (void*)c->svd.Z; *ldz = c->svd.ldz; *zrows = c->svd.jrows + c->svd.ncols_pad; *nblocks = nb; *tol = svd_dist_tol(c->svd);
  });
}
int32_t tn_svd_dist_step(tn_ctx* ctx, const int32_t* pairs, int32_t npairs, double* offmax) {
  return guard([&] {
    TN_CHECK(ctx && pairs && offmax, "tn_svd_dist_step: null pointer");
    *offmax = svd_dist_step(ctx->c.svd, pairs, npairs, ctx->c.stream);
  });
}
int32_t tn_svd_dist_finish(tn_ctx* ctx, tn_trunc_t tr, int32_t sweeps, int64_t* k_out) {
  return guard([&] {
    TN_CHECK(ctx && k_out, "tn_svd_dist_finish: null pointer");
    *k_out = svd_dist_finish(ctx->c.svd, T(tr), sweeps, ctx->c.stream);
  });
}
int32_t tn_svd_dist_factors(tn_ctx* ctx, void* U_dev, void* S_dev, void* Vh_dev) {
  return guard([&] {
    TN_CHECK(ctx, "tn_svd_dist_factors: null handle"); use_device(ctx);
    Ctx* c = &ctx->c; cudaStream_t s = c->stream;
    if (U_dev) svd_gather_U(c->svd, (cplx*)U_dev, c->svd.m, false, s);
    if (Vh_dev) svd_gather_Vh(c->svd, (cplx*)Vh_dev, c->svd.k, false, s);
    if (S_dev) svd_copy_S(c->svd, (double*)S_dev, s);
    c->sync();
  });
}
int32_t tn_mps_replacesites_factored(tn_mps* m, int32_t site, int32_t direction, int32_t normalize) {
  return guard([&] {
    TN_CHECK(m, "tn_mps_replacesites_factored: null handle"); use_device(m);
    mps_replacesites2_factored(m->m, site, direction != 0, normalize != 0);
    m->m->ctx->sync();
  });
}

int32_t tn_eigsolve_fn(tn_ctx* ctx, int64_t n, const void* theta0_dev, void* theta_out_dev, tn_lanczos_t lz, tn_apply_fn apply, void* user,
                       double* eig, int32_t* numops) {
  return guard([&] { TN_CHECK(ctx, "tn_eigsolve_fn: null handle"); use_device(ctx);
    TN_CHECK(ctx && theta0_dev && theta_out_dev && apply && eig && n >= 1, "eigsolve: bad arguments");
    Ctx* c = &ctx->c;
    int ops = 0;
    double v = lanczos_core(c, [&](const cplx* in, cplx* out) {
      c->sync();                                   // `in` is complete before the caller's map reads it
      int32_t st = apply(user, (const void*)in, (void*)out);
      if (st != 0) throw tn::Error(TN_ERR_INVALID, "eigsolve: the caller's linear map returned status " + std::to_string(st));
    }, (const cplx*)theta0_dev, (cplx*)theta_out_dev, n, Lanczos{lz.krylovdim, lz.maxiter, lz.tol}, &ops);
    c->sync();
    *eig = v;
    if (numops) *numops = ops;
  });
}

// ---- infinite TEBD (Vidal form, two-site cell) ------------------------------------------------------------------
int32_t tn_imps_create(tn_ctx* ctx, int32_t d, int32_t L, const int64_t* dims, const tn_cplx* const* site_ptrs, const double* const* sing_ptrs,
                       tn_imps** out) {
  return guard([&] {
    TN_CHECK(ctx && dims && site_ptrs && sing_ptrs && out, "tn_imps_create: null pointer");
    std::vector<long long> dd((size_t)3 * L);
    for (size_t i = 0; i < dd.size(); ++i) dd[i] = dims[i];
    auto* h = new tn_imps();
    try { h->m = imps_create(&ctx->c, d, L, dd.data(), reinterpret_cast<const cplx* const*>(site_ptrs), sing_ptrs); } catch (...) { delete h; throw; }
    *out = h;
  });
}
int32_t tn_imps_free(tn_imps* p) { return guard([&] { if (p) { imps_free(p->m); delete p; } }); }
int32_t tn_imps_dims(tn_imps* p, int64_t* dims) {
  return guard([&] {
    TN_CHECK(p && dims, "tn_imps_dims: null pointer");
    for (int i = 0; i < p->m->L; ++i) for (int k = 0; k < 3; ++k) dims[3 * i + k] = p->m->gam[i].dims[k];
  });
}
int32_t tn_imps_download(tn_imps* p, int32_t site, tn_cplx* tensor_out, double* singulars_out, double* norm_out) {
  return guard([&] {
    TN_CHECK(p, "tn_imps_download: null handle"); use_device(p);
    IMps* m = p->m; Ctx* c = m->ctx;
    TN_CHECK(site >= 1 && site <= m->L, "site index out of range");
    if (tensor_out) TN_CUDA(cudaMemcpyAsync(tensor_out, m->gam[site - 1].p, (size_t)m->gam[site - 1].size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    if (singulars_out) TN_CUDA(cudaMemcpyAsync(singulars_out, m->sing[site - 1], (size_t)m->nsing[site - 1] * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    if (norm_out) *norm_out = m->norms[site - 1];
  });
}
int32_t tn_itebd_apply_gate(tn_imps* p, const tn_cplx* gate_host, int32_t nsteps, tn_trunc_t tr) {
  return guard([&] {
    TN_CHECK(p && gate_host && nsteps >= 0, "tn_itebd_apply_gate: bad arguments");
    IMps* m = p->m; Ctx* c = m->ctx;
    const size_t ne = (size_t)m->d * m->d * m->d * m->d;
    cplx* g = c->scratch[15].get(ne, c->stream);
    TN_CUDA(cudaMemcpyAsync(g, gate_host, ne * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
    c->sync();
    for (int i = 0; i < nsteps; ++i) itebd_apply_gate2(m, g, T(tr));
    c->sync();
  });
}

// ---- MPO-bond-sharded H_eff application over several GPUs of this process (tn_shard.cu) ------------------------------------
int32_t tn_heff_sharded_create(int32_t ngpus, const int32_t* device_ids, int64_t chi, int64_t chi2, int32_t d, int64_t w, int64_t w1, int64_t w2,
                               const tn_cplx* L, const tn_cplx* R, const tn_cplx* M1, const tn_cplx* M2, tn_cplx coeff, tn_shard** out) {
  return guard([&] {
    TN_CHECK(out && device_ids, "tn_heff_sharded_create: null pointer");
    std::vector<int> dev(device_ids, device_ids + std::max(0, (int)ngpus));
    Shard* s = shard_create(ngpus, dev.data(), chi, chi2, d, w, w1, w2, C(L), C(R), C(M1), C(M2), cplx{coeff.re, coeff.im});
    *out = new tn_shard{s};
  });
}
int32_t tn_heff_sharded_apply(tn_shard* h, const tn_cplx* theta, tn_cplx* out) {
  return guard([&] { TN_CHECK(h && theta && out, "tn_heff_sharded_apply: null handle"); shard_apply(h->s, C(theta), C(out)); });
}
int32_t tn_heff_sharded_free(tn_shard* h) { return guard([&] { if (h) { shard_free(h->s); delete h; } }); }

int32_t tn_inner_oplist(tn_mps* psi, tn_mps* phi, int32_t nterms, const int32_t* nops, const int32_t* op_sites, const tn_cplx* ops_host,
                        const tn_cplx* coeffs, tn_cplx* out) {
  return guard([&] { TN_CHECK(psi && phi, "tn_inner_oplist: null handle"); use_device(psi);
    TN_CHECK(psi && phi && (nterms == 0 || (nops && op_sites && ops_host && coeffs && out)), "inner: null pointer");
    TN_CHECK(psi->m->ctx == phi->m->ctx, "inner: both MPSs must live in the same context");
    inner_oplist(psi->m, phi->m, nterms, nops, op_sites, C(ops_host), C(coeffs), C(out));
  });
}

int32_t tn_qjmc_ensemble(int32_t device, int32_t nworkers, int32_t ntraj, const uint64_t* traj_ids,
                         int32_t d, int32_t N, const int64_t* dims, const tn_cplx* const* site_ptrs, int32_t center,
                         int32_t nrows, const int32_t* counts, const int32_t* gate_sites, const int32_t* gate_nsites,
                         const tn_cplx* const* gate_ptrs, int32_t njump, const int32_t* jump_sites, const tn_cplx* jump_ops,
                         const double* jump_coeffs, int32_t steps, double dt, tn_trunc_t tr, uint64_t seed, const tn_cplx* obs_op,
                         int32_t save_every, tn_cplx* obs_out, int32_t* njumps_out, int32_t* jumps_out, double* jumptimes_out,
                         int32_t jump_cap, int32_t classical) {
  return guard([&] {
    TN_CHECK(ntraj >= 0 && nworkers >= 1 && dims && site_ptrs && counts && gate_sites && gate_nsites && gate_ptrs, "qjmc_ensemble: bad arguments");
    TN_CHECK(!obs_op || (obs_out && save_every > 0), "qjmc_ensemble: obs_out / save_every missing");
    std::vector<long long> dd((size_t)N * 3);
    for (size_t i = 0; i < dd.size(); ++i) dd[i] = dims[i];
    const int nsaves = (obs_op && save_every > 0) ? steps / save_every : 0;
    const int nw = std::min<int>(nworkers, std::max(1, ntraj));
    std::atomic<int> next{0};
    std::mutex err_mu; std::string err; int err_code = 0;
    // TN_QJMC_BATCH=1: the workers' truncated SVDs are collected into batching rounds (tn_svd.cuh: SvdBatcher) -- every round
    // factorises the pending SVD of every active trajectory, grouped by shape, in one batched pipeline instead of nw
    // concurrent single-problem pipelines.  On by default (C4 shapes on a B200, 32 trajectories x 2 steps: 1.00 -> 2.49
    // trajectory-steps/s and 11x fewer launches, profiles/r02_first_run_qjmc_batching_and_sharded1.jsonl); TN_QJMC_BATCH=0
    // restores the independent multi-stream workers.
    const char* benv = getenv("TN_QJMC_BATCH");
    // The workers are split into groups with one batcher each (default 2 groups from 16 workers on, 4 from 64 on, TN_QJMC_GROUPS): while one group's
    // round sits in its latency-bound phases (pair EVD, panel Cholesky) the other group's GEMMs fill the SMs -- the only overlap
    // available to a Jacobi step, whose own phases depend on each other.
    int ngroups = 1;
    if (!(benv && benv[0] == '0') && nw > 1) {
      const char* genv = getenv("TN_QJMC_GROUPS");
      ngroups = genv ? std::max(1, atoi(genv)) : (nw >= 64 ? 4 : (nw >= 16 ? 2 : 1));   // 64 workers: 4.5 (2 groups) -> 5.0 (4) trajectory-steps/s at C4
      ngroups = std::min(ngroups, nw / 2);
      ngroups = std::max(ngroups, 1);
    }
    std::vector<SvdBatcher*> batchers;
    if (!(benv && benv[0] == '0') && nw > 1)
      for (int g = 0; g < ngroups; ++g) batchers.push_back(svd_batcher_create(nw / ngroups + (g < nw % ngroups ? 1 : 0)));
    std::atomic<int> next_worker{0};
    auto worker = [&]() {
      const int widx = next_worker.fetch_add(1);
      SvdBatcher* batcher = batchers.empty() ? nullptr : batchers[widx % ngroups];
      Ctx c;
      Gates* g = nullptr; Mps* psi = nullptr;
      bool attached = false;
      try {
        ctx_init(c, device);
        if (batcher) { svd_batcher_attach(batcher); attached = true; }
        g = gates_create(&c, d, nrows, counts, gate_sites, gate_nsites, reinterpret_cast<const cplx* const*>(gate_ptrs));
        for (;;) {
          { std::lock_guard<std::mutex> lk(err_mu); if (err_code) break; }
          const int i = next.fetch_add(1);
          if (i >= ntraj) break;
          psi = mps_create(&c, 1, d, N, dd.data(), reinterpret_cast<const cplx* const*>(site_ptrs), center);
          int nj = qjmc_run(psi, g, njump, jump_sites, C(jump_ops), jump_coeffs, steps, dt, T(tr), nullptr, seed,
                            traj_ids ? traj_ids[i] : (uint64_t)i, obs_op ? C(obs_op) : nullptr, save_every,
                            obs_op ? C(obs_out) + (size_t)i * nsaves * N : nullptr,
                            jumps_out ? jumps_out + (size_t)i * jump_cap : nullptr,
                            jumptimes_out ? jumptimes_out + (size_t)i * jump_cap : nullptr, jump_cap, classical != 0);
          if (njumps_out) njumps_out[i] = nj;
          mps_free(psi); psi = nullptr;
        }
      } catch (const tn::Error& e) {
        std::lock_guard<std::mutex> lk(err_mu); if (!err_code) { err_code = e.code; err = e.what(); }
      } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(err_mu); if (!err_code) { err_code = TN_ERR_INTERNAL; err = e.what(); }
      }
      if (batcher) {                       // leave the rounds (also when ctx_init failed: this worker was counted as active)
        if (!attached) svd_batcher_attach(batcher);
        svd_batcher_detach(c.stream);
      }
      if (psi) mps_free(psi);
      if (g) gates_free(g);
      ctx_release(c);
    };
    std::vector<std::thread> pool;
    for (int k = 0; k < nw; ++k) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    for (SvdBatcher* batcher : batchers) {
      long long probs = 0; const long long rounds = svd_batcher_rounds(batcher, &probs);
      if (getenv("TN_QJMC_BATCH_STATS")) fprintf(stderr, "{\"qjmc_batching\": {\"workers\": %d, \"groups\": %d, \"rounds\": %lld, \"svds\": %lld}}\n", nw, ngroups, rounds, probs);
      cudaSetDevice(device);
      svd_batcher_destroy(batcher);
    }
    if (err_code) throw tn::Error(err_code, "qjmc_ensemble: " + err);
  });
}

}  // extern "C"
