// Host-side orchestration of the MPS hot path; all tensors stay resident in HBM.
// Reference call sites replaced (all under /root/reference/src):
//   structures/mps/projmps.jl:50-66   buildleft!   -> env_buildleft
//   structures/mps/projmps.jl:74-95   buildright!  -> env_buildright
//   structures/mps/projmps.jl:107-134 product      -> env_product_dev   (flop-optimal order (L.Theta).W.R)
//   structures/mps/projmps.jl:192-216 calculate    -> env_calculate
//   structures/mps/gmps.jl:29-112     norm / normalize! / move*!        -> mps_norm / mps_normalize / mps_movecenter
//   structures/mps/gmps.jl:199-267    replacesites! -> mps_replacesites2
//   structures/mps/gatelist.jl:137-227 applygate / applygates!          -> apply_gates
//   algorithms/mps/dmrg.jl:34-63      sweep body   -> dmrg_halfsweep (+ KrylovKit eigsolve -> lanczos_lowest)
#include "tn_mps.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace tn {
void count_launch(int n);

static const cplx ONE = {1.0, 0.0};
cplx* Buf::get(size_t n, cudaStream_t s) {
  if (n > cap) {
    if (p) TN_CUDA(cudaFreeAsync(p, s));
    size_t want = n + n / 8;
    TN_CUDA(cudaMallocAsync((void**)&p, want * sizeof(cplx), s));
    cap = want;
  }
  return p;
}
void Buf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

void Ctx::alloc(Tensor& t, const std::vector<long long>& dims) {
  long long n = 1; for (auto d : dims) n *= d;
  if ((size_t)n > t.cap) {
    if (t.p) TN_CUDA(cudaFreeAsync(t.p, stream));
    TN_CUDA(cudaMallocAsync((void**)&t.p, (size_t)n * sizeof(cplx), stream));
    t.cap = (size_t)n;
  }
  t.dims = dims;
}
void Ctx::free(Tensor& t) { if (t.p) cudaFreeAsync(t.p, stream); t.p = nullptr; t.cap = 0; t.dims.clear(); }

// ================================================================================================
// MPS container
// ================================================================================================
long long Mps::maxbonddim() const {
  long long D = 0;
  for (int i = 1; i < N; ++i) D = std::max(D, sites[i].dims.front());
  return D;
}

Mps* mps_create(Ctx* c, int rank, int d, int N, const long long* dims, const cplx* const* host_sites, int center) {
  TN_CHECK(rank == 1 || rank == 2, "GMPS rank must be 1 (MPS) or 2 (MPO)");
  TN_CHECK(N >= 1 && d >= 1, "bad MPS size");
  auto m = std::make_unique<Mps>();
  m->ctx = c; m->rank = rank; m->d = d; m->N = N; m->center = center;
  m->sites.resize(N);
  for (int i = 0; i < N; ++i) {
    std::vector<long long> dd(dims + (size_t)i * (rank + 2), dims + (size_t)(i + 1) * (rank + 2));
    for (int k = 1; k <= rank; ++k) TN_CHECK(dd[k] == d, "physical dimension mismatch");
    if (i > 0) TN_CHECK(dd.front() == m->sites[i - 1].dims.back(), "bond dimensions of neighbouring sites differ");
    c->alloc(m->sites[i], dd);
    TN_CUDA(cudaMemcpyAsync(m->sites[i].p, host_sites[i], (size_t)m->sites[i].size() * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  }
  c->sync();
  return m.release();
}
void mps_free(Mps* m) {
  if (!m) return;
  for (auto& t : m->sites) m->ctx->free(t);
  delete m;
}
void mps_download_site(Mps* m, int i, cplx* host) {
  TN_CHECK(i >= 1 && i <= m->N, "site index out of range");
  TN_CUDA(cudaMemcpyAsync(host, m->sites[i - 1].p, (size_t)m->sites[i - 1].size() * sizeof(cplx), cudaMemcpyDeviceToHost, m->ctx->stream));
  m->ctx->sync();
}
void mps_upload_site(Mps* m, int i, const long long* dims, const cplx* host) {
  TN_CHECK(i >= 1 && i <= m->N, "site index out of range");
  std::vector<long long> dd(dims, dims + m->rank + 2);
  m->ctx->alloc(m->sites[i - 1], dd);
  TN_CUDA(cudaMemcpyAsync(m->sites[i - 1].p, host, (size_t)m->sites[i - 1].size() * sizeof(cplx), cudaMemcpyHostToDevice, m->ctx->stream));
  m->ctx->sync();
}

cplx read_scalar(Ctx* c, int slot) {
  TN_CUDA(cudaMemcpyAsync(c->hscal + slot, c->dscal + slot, sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  return c->hscal[slot];
}

cplx mps_norm(Mps* m) {   // gmps.jl:29-37
  Ctx* c = m->ctx;
  if (m->center == 0) mps_movecenter(m, 1, Trunc{0.0, 0, 1});
  Tensor& A = m->sites[m->center - 1];
  const cplx* xs[1] = {A.p};
  zdots(A.size(), 1, xs, A.p, c->dscal, c->partials, c->stream);
  cplx v = read_scalar(c, 0);
  // complex square root of <A,A> (principal branch), as Julia's prod[1]^0.5
  double r = std::hypot(v.x, v.y), re = std::sqrt(0.5 * (r + v.x)), im = std::sqrt(std::max(0.0, 0.5 * (r - v.x)));
  return cplx{re, v.y < 0 ? -im : im};
}
void mps_normalize(Mps* m) {   // gmps.jl:46-51
  cplx n = mps_norm(m);
  double den = n.x * n.x + n.y * n.y;
  cplx inv = {n.x / den, -n.y / den};
  Tensor& A = m->sites[m->center - 1];
  zscal(A.size(), inv, A.p, m->ctx->stream);
}

// gauge move to the right (gmps.jl:75-82): psi[i] = U, psi[i+1] = (S V^H) psi[i+1]
// (s_stays: the MPO compression sweep of mpo.jl:443-449 keeps S on the site instead: O[i] = U S, O[i+1] = V^H O[i+1])
static void moveright(Mps* m, int i, Trunc tr, bool s_stays = false) {
  if (!(0 < i && i < m->N)) return;
  Ctx* c = m->ctx; cudaStream_t s = c->stream;
  Tensor& A = m->sites[i - 1]; Tensor& Bn = m->sites[i];
  long long rows = A.size() / A.dims.back(), cols = A.dims.back();
  int k = svd_factor(c->svd, A.p, (int)rows, (int)cols, rows, tr, s, s_stays ? 2 : 1, s_stays); c->svds++;   // the factor gathered without S is the isometry; a plain gauge move needs no singular values
  Tensor U; std::vector<long long> du = A.dims; du.back() = k;
  c->alloc(U, du);
  svd_gather_U(c->svd, U.p, rows, s_stays, s);
  cplx* SV = c->scratch[0].get((size_t)k * cols, s);
  svd_gather_Vh(c->svd, SV, k, !s_stays, s);
  long long tail = Bn.size() / Bn.dims.front();
  Tensor Nn; std::vector<long long> dn = Bn.dims; dn.front() = k;
  c->alloc(Nn, dn);
  zgemm_auto(mk(k, (int)tail, (int)cols, SV, idx1(1), idx1(k), 0, Bn.p, idx1(1), idx1(cols), 0, Nn.p, idx1(1), idx1(k)), s);
  c->free(A); c->free(Bn);
  m->sites[i - 1] = U; m->sites[i] = Nn;
}
// gauge move to the left (gmps.jl:60-67): psi[i] = V^H reshaped, psi[i-1] = psi[i-1] (U S)
// (s_stays: mpo.jl:451-457 keeps S on the site: O[i] = S V^H, O[i-1] = O[i-1] U)
static void moveleft(Mps* m, int i, Trunc tr, bool s_stays = false) {
  if (!(1 < i && i <= m->N)) return;
  Ctx* c = m->ctx; cudaStream_t s = c->stream;
  Tensor& A = m->sites[i - 1]; Tensor& Pv = m->sites[i - 2];
  long long rows = A.dims.front(), cols = A.size() / rows;
  int k = svd_factor(c->svd, A.p, (int)rows, (int)cols, rows, tr, s, s_stays ? 1 : 2, s_stays); c->svds++;
  Tensor V; std::vector<long long> dv = A.dims; dv.front() = k;
  c->alloc(V, dv);
  svd_gather_Vh(c->svd, V.p, k, s_stays, s);
  cplx* US = c->scratch[0].get((size_t)rows * k, s);
  svd_gather_U(c->svd, US, rows, !s_stays, s);
  long long head = Pv.size() / Pv.dims.back();
  Tensor Nn; std::vector<long long> dn = Pv.dims; dn.back() = k;
  c->alloc(Nn, dn);
  zgemm_auto(mk((int)head, k, (int)rows, Pv.p, idx1(1), idx1(head), 0, US, idx1(1), idx1(rows), 0, Nn.p, idx1(1), idx1(head)), s);
  c->free(A); c->free(Pv);
  m->sites[i - 1] = V; m->sites[i - 2] = Nn;
}

void mps_movecenter(Mps* m, int idx, Trunc tr) {   // gmps.jl:90-112
  TN_CHECK(idx >= 1 && idx <= m->N, "The index is out of range.");
  if (m->center == 0) {
    for (int i = 1; i <= idx - 1; ++i) moveright(m, i, tr);
    for (int i = 1; i <= m->N - idx; ++i) moveleft(m, m->N + 1 - i, tr);
  } else if (idx > m->center) {
    for (int i = m->center; i <= idx - 1; ++i) moveright(m, i, tr);
  } else if (idx < m->center) {
    for (int i = 1; i <= m->center - idx; ++i) moveleft(m, m->center + 1 - i, tr);
  }
  m->center = idx;
}

// Bond compression of a freshly assembled MPO, mpo.jl:443-457 (also addMPOs, mpo.jl:296-311): a right-going then a left-going
// sweep of truncated SVDs in which the singular values stay on the site that was factorised.
void mpo_compress(Mps* m, Trunc tr) {
  for (int i = 1; i <= m->N - 1; ++i) moveright(m, i, tr, true);
  for (int i = m->N; i >= 2; --i) moveleft(m, i, tr, true);
  m->center = 0;
}

// applyMPO(O, psi) = O * psi (mpo.jl:105-143, MPOMPSProduct): exact site products phi[i]((w,a), s, (w',b)) = sum_t M(w,s,t,w') A(a,t,b)
// (one strided GEMM with K = d per site; fused bonds with the MPO index fastest), a right-going untruncated gauge sweep while the
// sites are built, then movecenter!(phi, 1; kwargs...) from an unset centre, i.e. a truncating left-going sweep.
Mps* mpo_apply(Mps* O, Mps* psi, Trunc tr) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  TN_CHECK(O->rank == 2 && psi->rank == 1, "Unallowed combinations of MPS ranks.");
  TN_CHECK(O->d == psi->d && O->N == psi->N, "GMPS must share the same physical dims and length.");
  TN_CHECK(O->ctx == psi->ctx, "applyMPO: both arguments must live in the same context");
  const int N = psi->N, d = psi->d;
  auto phi = std::make_unique<Mps>();
  phi->ctx = c; phi->rank = 1; phi->d = d; phi->N = N; phi->center = 0;
  phi->sites.resize(N);
  for (int i = 1; i <= N; ++i) {
    const Tensor& M = O->sites[i - 1]; const Tensor& A = psi->sites[i - 1];
    const long long W = M.dims[0], X = M.dims[3], a = A.dims[0], b = A.dims[2];
    // the bond entering site i may have been changed by the gauge move of site i-1: the site is built with the exact product bond
    // and moveright(i-1) below contracts its S V^H into it, exactly as the reference does
    c->alloc(phi->sites[i - 1], {W * a, (long long)d, X * b});
    zgemm_auto(mk((int)(W * d * X), (int)(a * b), d, M.p, idx2((int)(W * d), 1, W * d * d), idx1(W * d), 0,
                  A.p, idx1(a), idx2((int)a, 1, a * d), 0,
                  phi->sites[i - 1].p, idx2((int)W, 1, W * a), idx2((int)a, W, W * a * d * X)), s);
    if (i > 1) moveright(phi.get(), i - 1, Trunc{0.0, 0, 1});
  }
  mps_movecenter(phi.get(), 1, tr);
  return phi.release();
}

// Split a two-site tensor theta (chi_l, p, p, chi_r), p = d^rank, back into two sites (gmps.jl:215-266).
void mps_replacesites2(Mps* m, const cplx* theta, int site, bool direction, bool normalize, Trunc tr) {
  Ctx* c = m->ctx; cudaStream_t s = c->stream;
  TN_CHECK(site >= 1 && site + 1 <= m->N, "replacesites: site out of range");
  Tensor& A = m->sites[site - 1]; Tensor& B = m->sites[site];
  long long chil = A.dims.front(), chir = B.dims.back(), p = m->phys();
  long long rows = chil * p, cols = p * chir;
  int k = svd_factor(c->svd, theta, (int)rows, (int)cols, rows, tr, s, direction ? 2 : 1); c->svds++;
  std::vector<long long> da = A.dims, db = B.dims; da.back() = k; db.front() = k;
  c->alloc(A, da); c->alloc(B, db);
  // sweeping right: left-orthonormal site + (S V^H); sweeping left: (U S) + right-orthonormal site
  svd_gather_U(c->svd, A.p, rows, direction, s);
  svd_gather_Vh(c->svd, B.p, k, !direction, s);
  m->center = direction ? site : site + 1;
  if (normalize) {          // normalize! (gmps.jl:46-51) without a host round trip: <A,A> of the centre is real, A <- A / sqrt(<A,A>)
    Tensor& Cn = m->sites[m->center - 1];
    const cplx* xs[1] = {Cn.p};
    zdots(Cn.size(), 1, xs, Cn.p, c->dscal + 50, c->partials, s);
    zscale_invnorm(Cn.size(), Cn.p, c->dscal + 50, Cn.p, s);
  }
}

// The part of mps_replacesites2 after the factorisation, for factors that are already in the context's SvdWork (a caller that
// ran the SVD itself, e.g. the distributed sweeps): same gathers, same centre / normalisation.
void mps_replacesites2_factored(Mps* m, int site, bool direction, bool normalize) {
  Ctx* c = m->ctx; cudaStream_t s = c->stream;
  TN_CHECK(site >= 1 && site + 1 <= m->N, "replacesites: site out of range");
  Tensor& A = m->sites[site - 1]; Tensor& B = m->sites[site];
  long long chil = A.dims.front(), chir = B.dims.back(), p = m->phys();
  long long rows = chil * p, cols = p * chir;
  TN_CHECK(c->svd.m == rows && c->svd.n == cols, "replacesites: the factorisation in the context does not belong to these sites");
  const int k = c->svd.k;
  std::vector<long long> da = A.dims, db = B.dims; da.back() = k; db.front() = k;
  c->alloc(A, da); c->alloc(B, db);
  svd_gather_U(c->svd, A.p, rows, direction, s);
  svd_gather_Vh(c->svd, B.p, k, !direction, s);
  m->center = direction ? site : site + 1;
  if (normalize) mps_normalize(m);
}

void mps_applyop1(Mps* m, int site, const cplx* op_dev) {   // mps.jl:141-152
  Ctx* c = m->ctx;
  Tensor& A = m->sites[site - 1];
  Tensor Nn; c->alloc(Nn, A.dims);
  int inner = m->rank == 2 ? m->d : 1;
  op_apply1(A.p, Nn.p, op_dev, A.dims.front(), m->d, inner, A.dims.back(), c->stream);
  c->free(A);
  m->sites[site - 1] = Nn;
}

void mps_bond_spectrum(Mps* m, int site, std::vector<double>& out) {   // gmps.jl:184-189 (the SVD inside entropy())
  Ctx* c = m->ctx;
  mps_movecenter(m, site, Trunc{0.0, 0, 1});
  Tensor& A = m->sites[site - 1];
  long long rows = A.size() / A.dims.back(), cols = A.dims.back();
  int k = svd_factor(c->svd, A.p, (int)rows, (int)cols, rows, Trunc{0.0, 0, 1}, c->stream);
  out.resize(k);
  TN_CUDA(cudaMemcpyAsync(out.data(), c->svd.sig, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
}

// ================================================================================================
// Environments
// ================================================================================================
__global__ void fill_ones_kernel(cplx* p, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = make_double2(1.0, 0.0); }

Env* env_create(Ctx* c, Mps* bra, Mps* mpo, Mps* ket, cplx coeff, int center) {
  TN_CHECK(bra && ket, "The projection must have a braket structure");
  TN_CHECK(bra->rank == 1 && ket->rank == 1, "The projection must have a braket structure");
  TN_CHECK(bra->N == ket->N && bra->d == ket->d, "GMPS must share the same length and physical dim.");
  if (mpo) TN_CHECK(mpo->rank == 2 && mpo->N == ket->N && mpo->d == ket->d, "MPO must be rank 2 and share length / physical dim.");
  auto e = std::make_unique<Env>();
  e->ctx = c; e->bra = bra; e->mpo = mpo; e->ket = ket; e->coeff = coeff; e->center = 0;
  e->blocks.resize(ket->N);
  c->alloc(e->edge, {1, 1, 1});
  fill_ones_kernel<<<1, 32, 0, c->stream>>>(e->edge.p, 1);
  count_launch(1);
  Env* raw = e.release();
  env_movecenter(raw, center);
  return raw;
}
void env_free(Env* e) {
  if (!e) return;
  for (auto& t : e->blocks) e->ctx->free(t);
  e->ctx->free(e->edge);
  e->ctx->free(e->phi);
  delete e;
}
const Tensor& env_block(Env* e, int idx) {   // abstractprojmps.jl:43-46
  if (idx < 1 || idx > e->ket->N) return e->edge;
  TN_CHECK(e->blocks[idx - 1].p != nullptr, "environment block has not been built");
  return e->blocks[idx - 1];
}

void env_buildleft(Env* e, int idx) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  const Tensor& L = env_block(e, idx - 1);
  const Tensor& A1 = e->bra->sites[idx - 1];
  const Tensor& A2 = e->ket->sites[idx - 1];
  int d = e->ket->d;
  int ca = (int)A1.dims[0], ca2 = (int)A1.dims[2], cb = (int)A2.dims[0], cb2 = (int)A2.dims[2];
  int w = e->mpo ? (int)e->mpo->sites[idx - 1].dims[0] : 1, w2 = e->mpo ? (int)e->mpo->sites[idx - 1].dims[3] : 1;
  TN_CHECK(L.dims[0] == ca && L.dims[1] == w && L.dims[2] == cb, "buildleft: block / site dimension mismatch");
  // X1[(a,w),(s',b')] = L[(a,w), b] A2[b,(s',b')]
  cplx* X1 = c->scratch[1].get((size_t)ca * w * d * cb2, s);
  zgemm_auto(mk(ca * w, d * cb2, cb, L.p, idx1(1), idx1((long long)ca * w), 0, A2.p, idx1(1), idx1(cb), 0, X1, idx1(1), idx1((long long)ca * w)), s);
  cplx* X2 = X1;   // without an MPO layer (a, s, b') == (a, w=1, s', b')
  if (e->mpo) {
    // X2(a,s,w',b') = sum_{w,s'} X1(a,w,s',b') M(w,s,s',w'); rows m = (a,b'), k = (w,s'), n = (s,w')
    const Tensor& M = e->mpo->sites[idx - 1];
    X2 = c->scratch[2].get((size_t)ca * d * w2 * cb2, s);
    zgemm_auto(mk(ca * cb2, d * w2, w * d,
                  X1, idx2(ca, 1, (long long)ca * w * d), idx1(ca), 0,
                  M.p, idx2(w, 1, (long long)w * d), idx2(d, w, (long long)w * d * d), 0,
                  X2, idx2(ca, 1, (long long)ca * d * w2), idx1(ca)), s);
  }
  // L'(a',(w',b')) = sum_{(a,s)} conj(A1[(a,s),a']) X2[(a,s),(w',b')]
  Tensor& out = e->blocks[idx - 1];
  c->alloc(out, {ca2, w2, cb2});
  zgemm_auto(mk(ca2, w2 * cb2, ca * d, A1.p, idx1((long long)ca * d), idx1(1), 1, X2, idx1(1), idx1((long long)ca * d), 0, out.p, idx1(1), idx1(ca2)), s);
}

void env_buildright(Env* e, int idx) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  const Tensor& R = env_block(e, idx + 1);
  const Tensor& A1 = e->bra->sites[idx - 1];
  const Tensor& A2 = e->ket->sites[idx - 1];
  int d = e->ket->d;
  int cal = (int)A1.dims[0], ca = (int)A1.dims[2], cbl = (int)A2.dims[0], cb = (int)A2.dims[2];
  int wl = e->mpo ? (int)e->mpo->sites[idx - 1].dims[0] : 1, w = e->mpo ? (int)e->mpo->sites[idx - 1].dims[3] : 1;
  TN_CHECK(R.dims[0] == ca && R.dims[1] == w && R.dims[2] == cb, "buildright: block / site dimension mismatch");
  // Y1(b_l, a, w, s') = sum_b A2(b_l,s',b) R(a,w,b); rows m = (b_l,s') written with a 2-level stride
  cplx* Y1 = c->scratch[1].get((size_t)cbl * ca * w * d, s);
  zgemm_auto(mk(cbl * d, ca * w, cb, A2.p, idx1(1), idx1((long long)cbl * d), 0, R.p, idx1((long long)ca * w), idx1(1), 0,
                Y1, idx2(cbl, 1, (long long)cbl * ca * w), idx1(cbl)), s);
  cplx* Y2 = Y1;   // without an MPO layer Y1(b_l, a, s') already is Y2(b_l, a, s)
  if (e->mpo) {
    // Y2[(b_l,a),(s,w_l)] = sum_{(w,s')} Y1[(b_l,a),(w,s')] M(w_l,s,s',w)
    const Tensor& M = e->mpo->sites[idx - 1];
    Y2 = c->scratch[2].get((size_t)cbl * ca * d * wl, s);
    zgemm_auto(mk(cbl * ca, d * wl, w * d, Y1, idx1(1), idx1((long long)cbl * ca), 0,
                  M.p, idx2(w, (long long)wl * d * d, (long long)wl * d), idx2(d, wl, 1), 0,
                  Y2, idx1(1), idx1((long long)cbl * ca)), s);
  }
  // R'(a_l, w_l, b_l) = sum_{(a,s)} conj(A1)(a_l,s,a) Y2(b_l,a,s,w_l); k = (a,s) with a fastest (uniform stride in Y2),
  // n = (b_l, w_l) with b_l fastest so that both operands are read along their unit-stride index
  Tensor& out = e->blocks[idx - 1];
  c->alloc(out, {cal, wl, cbl});
  zgemm_auto(mk(cal, cbl * wl, d * ca, A1.p, idx1(1), idx2(ca, (long long)cal * d, cal), 1,
                Y2, idx1(cbl), idx2(cbl, 1, (long long)cbl * ca * d), 0,
                out.p, idx1(1), idx2(cbl, (long long)cal * wl, cal)), s);
}

void env_movecenter(Env* e, int idx) {   // abstractprojmps.jl:60-81
  int N = e->ket->N;
  TN_CHECK(idx >= 1 && idx <= N, "The index is out of range.");
  if (e->center == 0) {
    for (int i = 1; i <= idx - 1; ++i) env_buildleft(e, i);
    for (int i = 1; i <= N - idx; ++i) env_buildright(e, N + 1 - i);
  } else if (idx > e->center) {
    for (int i = 1; i <= idx - e->center; ++i) env_buildleft(e, e->center - 1 + i);
  } else if (idx < e->center) {
    for (int i = 1; i <= e->center - idx; ++i) env_buildright(e, e->center + 1 - i);
  }
  e->center = idx;
}

// W[(w,s1',s2'),(s1,s2,w2)] = sum_{w1} M1(w,s1,s1',w1) M2(w1,s2,s2',w2)
__global__ void build_w2_kernel(const cplx* __restrict__ M1, const cplx* __restrict__ M2, cplx* __restrict__ W, int w, int w1, int w2, int d) {
  int K = w * d * d, Nn = d * d * w2;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * Nn) return;
  int kk = e % K, nn = e / K;
  int iw = kk % w, s1p = (kk / w) % d, s2p = kk / (w * d);
  int s1 = nn % d, s2 = (nn / d) % d, iw2 = nn / (d * d);
  double xr = 0, xi = 0;
  for (int j = 0; j < w1; ++j) {
    cplx a = M1[iw + w * (s1 + d * (s1p + d * (long long)j))];
    cplx b = M2[j + w1 * (s2 + d * (s2p + d * (long long)iw2))];
    xr += a.x * b.x - a.y * b.y; xi += a.x * b.y + a.y * b.x;
  }
  W[e] = make_double2(xr, xi);
}

void build_w2(const cplx* M1, const cplx* M2, cplx* W, int w, int w1, int w2, int d, cudaStream_t s) {
  int tot = w * d * d * d * d * w2;
  build_w2_kernel<<<(tot + 127) / 128, 128, 0, s>>>(M1, M2, W, w, w1, w2, d);
  count_launch(1);
}

// H_eff * theta for sites (site, site+1):  out(a,s1,s2,a') = coeff * sum L M1 M2 theta R   (projmps.jl:107-134, :144)
// The matvec is split into two pieces so that the host-buffer entry point can pipeline PCIe copies with compute:
//   stage 1+2 for a slice b' in [b0,b1) of theta's right bond (theta, T1 and T2 all have b' as their slowest index, so a
//   slice is a contiguous chunk of each), stage 3 for a slice a' in [a0,a1) of the output's right bond.
struct HeffDims { int d, d2, ca, w, cb, ca2, w2, cb2, w1; };
static HeffDims heff_dims(Env* e, int site) {
  TN_CHECK(e->mpo != nullptr, "product: the rank-2 branch needs an MPO layer");
  TN_CHECK(site >= 1 && site + 1 <= e->ket->N, "product: site out of range");
  const Tensor& L = env_block(e, site - 1);
  const Tensor& R = env_block(e, site + 2);
  const Tensor& M1 = e->mpo->sites[site - 1];
  const Tensor& M2 = e->mpo->sites[site];
  HeffDims h;
  h.d = e->ket->d; h.d2 = h.d * h.d;
  h.ca = (int)L.dims[0]; h.w = (int)L.dims[1]; h.cb = (int)L.dims[2];
  h.ca2 = (int)R.dims[0]; h.w2 = (int)R.dims[1]; h.cb2 = (int)R.dims[2];
  h.w1 = (int)M1.dims[3];
  TN_CHECK(M1.dims[0] == h.w && M2.dims[0] == h.w1 && M2.dims[3] == h.w2, "product: MPO / block bond mismatch");
  return h;
}
void heff_prepare(Env* e, int site) {   // W = M1.M2 and the intermediates' storage
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  HeffDims h = heff_dims(e, site);
  cplx* W = c->scratch[3].get((size_t)h.w * h.d2 * h.d2 * h.w2, s);
  int tot = h.w * h.d2 * h.d2 * h.w2;
  build_w2_kernel<<<(tot + 127) / 128, 128, 0, s>>>(e->mpo->sites[site - 1].p, e->mpo->sites[site].p, W, h.w, h.w1, h.w2, h.d);
  count_launch(1);
  c->scratch[4].get((size_t)h.ca * h.w * h.d2 * h.cb2, s);
  c->scratch[5].get((size_t)h.ca * h.d2 * h.w2 * h.cb2, s);
}
void heff_stage12(Env* e, const cplx* theta, int site, int b0, int b1, cudaEvent_t* ev_mid) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  HeffDims h = heff_dims(e, site);
  const Tensor& L = env_block(e, site - 1);
  const int nb = b1 - b0, ca = h.ca, w = h.w, d2 = h.d2, cb = h.cb, w2 = h.w2;
  cplx* W = c->scratch[3].p;
  cplx* T1 = c->scratch[4].p + (size_t)ca * w * d2 * b0;
  cplx* T2 = c->scratch[5].p + (size_t)ca * d2 * w2 * b0;
  const cplx* th = theta + (size_t)cb * d2 * b0;
  // T1[(a,w),(s1',s2',b')] = L[(a,w),b] theta[b,(s1',s2',b')]
  zgemm_auto(mk(ca * w, d2 * nb, cb, L.p, idx1(1), idx1((long long)ca * w), 0, th, idx1(1), idx1(cb), 0, T1, idx1(1), idx1((long long)ca * w)), s);
  if (ev_mid) TN_CUDA(cudaEventRecord(*ev_mid, s));
  // T2(a,s1,s2,w2,b') = sum_{(w,s1',s2')} T1(a,(w,s1',s2'),b') W; rows m = (a,b')
  zgemm_auto(mk(ca * nb, d2 * w2, w * d2, T1, idx2(ca, 1, (long long)ca * w * d2), idx1(ca), 0,
                W, idx1(1), idx1((long long)w * d2), 0,
                T2, idx2(ca, 1, (long long)ca * d2 * w2), idx1(ca)), s);
}
void heff_stage3(Env* e, int site, int a0, int a1, cplx* out) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  HeffDims h = heff_dims(e, site);
  const Tensor& R = env_block(e, site + 2);
  const int ca = h.ca, d2 = h.d2;
  // out[(a,s1,s2),a'] = coeff * sum_{(w2,b')} T2[(a,s1,s2),(w2,b')] R[a',(w2,b')]
  zgemm_auto(mk(ca * d2, a1 - a0, h.w2 * h.cb2, c->scratch[5].p, idx1(1), idx1((long long)ca * d2), 0, R.p + a0, idx1(h.ca2), idx1(1), 0,
                out + (size_t)ca * d2 * a0, idx1(1), idx1((long long)ca * d2), e->coeff), s);
}

// H_eff * theta for sites (site, site+1):  out(a,s1,s2,a') = coeff * sum L M1 M2 theta R   (projmps.jl:107-134, :144)
void env_product_dev(Env* e, const cplx* theta, int site, cplx* out, cudaEvent_t* ev4, bool prepared) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  HeffDims h = heff_dims(e, site);
  if (!prepared) heff_prepare(e, site);   // W = M1.M2 + intermediates; the Lanczos loop prepares once per bond
  if (ev4) TN_CUDA(cudaEventRecord(ev4[0], s));
  heff_stage12(e, theta, site, 0, h.cb2, ev4 ? &ev4[1] : nullptr);
  if (ev4) TN_CUDA(cudaEventRecord(ev4[2], s));
  heff_stage3(e, site, 0, h.ca2, out);
  if (ev4) TN_CUDA(cudaEventRecord(ev4[3], s));
  c->matvecs++;
}

// Host-buffer matvec with the PCIe copies pipelined against the contractions: theta is uploaded in slices of its right
// bond on a copy stream while stage 1+2 of the previous slice runs; the result is downloaded slice by slice while
// stage 3 computes the next one.  (theta_host / out_host should be pinned for the copies to be asynchronous.)
void env_product_host(Env* e, const cplx* theta_host, int site, cplx* out_host) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  HeffDims h = heff_dims(e, site);
  const long long n_in = (long long)h.cb * h.d2 * h.cb2, n_out = (long long)h.ca * h.d2 * h.ca2;
  cplx* din = c->scratch[13].get((size_t)n_in, s);
  cplx* dout = c->scratch[14].get((size_t)n_out, s);
  // slices of Theta's / the result's right bond: enough of them that the first upload and the last download (the only
  // copies not hidden behind a contraction stage) are short, but each slice still a few hundred columns wide
  // (TN_MATVEC_CHUNKS overrides)
  static int want_chunks = -1;
  if (want_chunks < 0) { const char* e = getenv("TN_MATVEC_CHUNKS"); want_chunks = e ? std::max(1, std::min((int)Ctx::MAX_CHUNKS, atoi(e))) : 0; }
  int nchunk = 1;
  if (n_in >= (1 << 20) && h.cb2 >= 64 && h.ca2 >= 64) nchunk = want_chunks ? want_chunks : std::max(4, std::min(8, std::min(h.cb2, h.ca2) / 128));
  nchunk = std::min(nchunk, std::min(h.cb2, h.ca2));
  if (nchunk == 1) {
    TN_CUDA(cudaMemcpyAsync(din, theta_host, (size_t)n_in * sizeof(cplx), cudaMemcpyHostToDevice, s));
    env_product_dev(e, din, site, dout, nullptr);
    TN_CUDA(cudaMemcpyAsync(out_host, dout, (size_t)n_out * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
    return;
  }
  if (!c->copy_stream) {
    TN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c->copy_ev) TN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  heff_prepare(e, site);
  TN_CUDA(cudaEventRecord(c->copy_ev[2 * Ctx::MAX_CHUNKS], s));     // scratch (re)allocation is ordered on the main stream
  TN_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[2 * Ctx::MAX_CHUNKS], 0));
  for (int k = 0; k < nchunk; ++k) {
    int b0 = (int)((long long)h.cb2 * k / nchunk), b1 = (int)((long long)h.cb2 * (k + 1) / nchunk);
    size_t off = (size_t)h.cb * h.d2 * b0, cnt = (size_t)h.cb * h.d2 * (b1 - b0);
    TN_CUDA(cudaMemcpyAsync(din + off, theta_host + off, cnt * sizeof(cplx), cudaMemcpyHostToDevice, c->copy_stream));
    TN_CUDA(cudaEventRecord(c->copy_ev[k], c->copy_stream));
  }
  for (int k = 0; k < nchunk; ++k) {
    int b0 = (int)((long long)h.cb2 * k / nchunk), b1 = (int)((long long)h.cb2 * (k + 1) / nchunk);
    TN_CUDA(cudaStreamWaitEvent(s, c->copy_ev[k], 0));
    heff_stage12(e, din, site, b0, b1, nullptr);
  }
  for (int k = 0; k < nchunk; ++k) {
    int a0 = (int)((long long)h.ca2 * k / nchunk), a1 = (int)((long long)h.ca2 * (k + 1) / nchunk);
    heff_stage3(e, site, a0, a1, dout);
    TN_CUDA(cudaEventRecord(c->copy_ev[Ctx::MAX_CHUNKS + k], s));
    TN_CUDA(cudaStreamWaitEvent(c->copy_stream, c->copy_ev[Ctx::MAX_CHUNKS + k], 0));
    size_t off = (size_t)h.ca * h.d2 * a0, cnt = (size_t)h.ca * h.d2 * (a1 - a0);
    TN_CUDA(cudaMemcpyAsync(out_host + off, dout + off, cnt * sizeof(cplx), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  c->matvecs++;
  TN_CUDA(cudaStreamSynchronize(c->copy_stream));
  c->sync();
}

cplx env_calculate(Env* e) {   // projmps.jl:192-216
  Ctx* c = e->ctx;
  int site = e->center;
  TN_CHECK(site >= 1, "calculate: the environment centre is not set");
  Tensor saved = e->blocks[site - 1];
  e->blocks[site - 1] = Tensor{};
  env_buildleft(e, site);                 // left block extended over the centre site (a', w', b')
  Tensor tmp = e->blocks[site - 1];
  e->blocks[site - 1] = saved;
  const Tensor& R = env_block(e, site + 1);
  TN_CHECK(tmp.size() == R.size(), "calculate: block size mismatch");
  // un-conjugated full contraction sum tmp .* R: zdots conjugates its first argument, so conjugate tmp first
  zconj_inplace(tmp.size(), tmp.p, c->stream);
  const cplx* xs[1] = {tmp.p};
  zdots(tmp.size(), 1, xs, R.p, c->dscal, c->partials, c->stream);
  cplx v = read_scalar(c, 0);
  c->free(tmp);
  return cplx{e->coeff.x * v.x - e->coeff.y * v.y, e->coeff.x * v.y + e->coeff.y * v.x};
}

// ================================================================================================
// Lanczos (KrylovKit eigsolve schedule, dmrg.jl:51-53) -- all vectors and coefficients on device
// ================================================================================================
static void eigh_sym3(int K, const double T[3][3], double* D, double U[3][3]) {
  // cyclic Jacobi on a real symmetric K x K (K <= 3) matrix; eigenvalues ascending
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
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int k = 0; k < K; ++k) { double x = A[k][p], y = A[k][q]; A[k][p] = cs * x - sn * y; A[k][q] = sn * x + cs * y; }
        for (int k = 0; k < K; ++k) { double x = A[p][k], y = A[q][k]; A[p][k] = cs * x - sn * y; A[q][k] = sn * x + cs * y; }
        for (int k = 0; k < K; ++k) { double x = U[k][p], y = U[k][q]; U[k][p] = cs * x - sn * y; U[k][q] = sn * x + cs * y; }
      }
  }
  int ord[3] = {0, 1, 2};
  for (int i = 1; i < K; ++i)            // insertion sort of <= 3 indices by eigenvalue
    for (int j = i; j > 0 && A[ord[j]][ord[j]] < A[ord[j - 1]][ord[j - 1]]; --j) std::swap(ord[j], ord[j - 1]);
  double Us[3][3];
  for (int j = 0; j < K; ++j) { D[j] = A[ord[j]][ord[j]]; for (int i = 0; i < K; ++i) Us[i][j] = U[i][ord[j]]; }
  for (int i = 0; i < K; ++i) for (int j = 0; j < K; ++j) U[i][j] = Us[i][j];
}

double lanczos_lowest(Env* e, int site, const cplx* theta0, cplx* theta_out, long long n, Lanczos lz, int* numops_out) {
  heff_prepare(e, site);    // once per bond: the <= 5 H_eff applications below share W and the intermediates
  return lanczos_core(e->ctx, [&](const cplx* in, cplx* out) { env_product_dev(e, in, site, out, nullptr, true); }, theta0, theta_out, n, lz, numops_out);
}

double lanczos_core(Ctx* c, const ApplyFn& apply, const cplx* theta0, cplx* theta_out, long long n, Lanczos lz, int* numops_out) {
  cudaStream_t s = c->stream;
  TN_CHECK(lz.krylovdim >= 1 && lz.krylovdim <= 3, "krylovdim must be 1..3 (the reference uses 3)");
  const int KD = lz.krylovdim;
  cplx* V[3]; for (int j = 0; j < 3; ++j) V[j] = c->scratch[6 + j].get((size_t)n, s);
  cplx* w = c->scratch[9].get((size_t)n, s);
  cplx* tmpv[3]; for (int j = 0; j < 3; ++j) tmpv[j] = c->scratch[10 + j].get((size_t)n, s);
  double T[3][3] = {{0}};
  cplx* ds = c->dscal;   // device scalars: [0..3] dots, [8] norm^2, [16+..] log of alpha/beta per step
  // v1 = theta0 / ||theta0||
  { const cplx* xs[1] = {theta0}; zdots(n, 1, xs, theta0, ds + 8, c->partials, s); zscale_invnorm(n, theta0, ds + 8, V[0], s); }
  int K = 1, numops = 0, numiter = 1;
  int logn = 0;   // scalars recorded on device: pairs (alpha_i at ds[16+2i], beta_i^2 at ds[17+2i])
  auto expand = [&](int Kc) {
    // w = H V[Kc-1]; alpha = Re<v,w>; w -= sum_j <v_j,w> v_j (twice); beta^2 = <w,w>
    apply(V[Kc - 1], w);
    numops++;
    const cplx* xs[3] = {V[0], V[1], V[2]};
    zdots(n, Kc, xs, w, ds, c->partials, s);
    TN_CUDA(cudaMemcpyAsync(ds + 16 + 2 * logn, ds + (Kc - 1), sizeof(cplx), cudaMemcpyDeviceToDevice, s));   // alpha (complex; real part used)
    zsubproj(n, Kc, xs, ds, w, s);
    zdots(n, Kc, xs, w, ds + 4, c->partials, s);
    zsubproj(n, Kc, xs, ds + 4, w, s);
    const cplx* ws[1] = {w};
    zdots(n, 1, ws, w, ds + 17 + 2 * logn, c->partials, s);
    logn++;
  };
  auto fetch = [&](int cnt, double* alpha, double* beta) {
    TN_CUDA(cudaMemcpyAsync(c->hscal + 16, ds + 16, sizeof(cplx) * 2 * cnt, cudaMemcpyDeviceToHost, s));
    c->sync();
    for (int i = 0; i < cnt; ++i) { alpha[i] = c->hscal[16 + 2 * i].x; beta[i] = std::sqrt(std::max(0.0, c->hscal[17 + 2 * i].x)); }
  };
  // ---- round 1: build the KD-dimensional factorisation without host round trips
  double alpha[3], beta[3];
  expand(1);
  for (int k = 2; k <= KD; ++k) {
    zscale_invnorm(n, w, ds + 17 + 2 * (logn - 1), V[k - 1], s);
    expand(k);
  }
  fetch(KD, alpha, beta);
  K = KD;
  for (int i = 0; i < KD; ++i) { T[i][i] = alpha[i]; if (i + 1 < KD) T[i][i + 1] = T[i + 1][i] = beta[i]; }
  double bet = beta[KD - 1];
  // early invariant-subspace exit inside round 1 (beta_j <= tol): truncate the factorisation there
  for (int i = 0; i < KD - 1; ++i) if (beta[i] <= lz.tol) { K = i + 1; bet = beta[i]; break; }
  double D[3], U[3][3];
  int converged = 0;
  while (true) {
    eigh_sym3(K, T, D, U);
    converged = 0;
    while (converged < K && std::fabs(U[K - 1][converged] * bet) <= lz.tol) converged++;
    if (converged >= 1 || bet <= lz.tol) break;
    if (K < KD) break;   // only reachable through the early-exit path
    if (numiter == lz.maxiter) break;
    // ---- thick restart: keep = div(3*KD + 2*converged, 5) Ritz vectors + the residual
    int keep = (3 * KD + 2 * converged) / 5;
    if (keep >= KD) keep = KD - 1;
    if (keep < 1) {   // krylovdim == 1: restart from the residual direction is meaningless; stop
      break;
    }
    const cplx* xs[3] = {V[0], V[1], V[2]};
    for (int j = 0; j < keep; ++j) { double cj[3] = {U[0][j], U[1][j], U[2][j]}; zlincomb(n, K, xs, cj, tmpv[j], s); }
    double f[3];
    for (int j = 0; j < keep; ++j) f[j] = U[K - 1][j] * bet;
    for (int j = 0; j < keep; ++j) std::swap(V[j], tmpv[j]);
    zscale_invnorm(n, w, ds + 17 + 2 * (logn - 1), V[keep], s);
    std::memset(T, 0, sizeof(T));
    for (int j = 0; j < keep; ++j) { T[j][j] = D[j]; T[j][keep] = T[keep][j] = f[j]; }
    logn = 0;
    int first = keep + 1;
    expand(first);
    for (int k = first + 1; k <= KD; ++k) {
      zscale_invnorm(n, w, ds + 17 + 2 * (logn - 1), V[k - 1], s);
      expand(k);
    }
    int cnt = KD - keep;
    fetch(cnt, alpha, beta);
    for (int i = 0; i < cnt; ++i) {
      int r = keep + i;
      T[r][r] = alpha[i];
      if (r + 1 < KD) T[r][r + 1] = T[r + 1][r] = beta[i];
    }
    bet = beta[cnt - 1];
    K = KD;
    numiter++;
  }
  // Ritz vector of the lowest Ritz value, renormalised
  {
    const cplx* xs[3] = {V[0], V[1], V[2]};
    double c0[3] = {U[0][0], U[1][0], U[2][0]};
    zlincomb(n, K, xs, c0, w, s);
    const cplx* ws[1] = {w};
    zdots(n, 1, ws, w, ds + 8, c->partials, s);
    zscale_invnorm(n, w, ds + 8, theta_out, s);
  }
  if (numops_out) *numops_out = numops;
  return D[0];
}

// ================================================================================================
// DMRG half sweep (dmrg.jl:35-63): all bonds in one direction, then environments to the end
// ================================================================================================
void dmrg_halfsweep(Mps* psi, Env* e, bool direction, Lanczos lz, Trunc tr, double* energy, long long* maxbond) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  TN_CHECK(psi->rank == 1, "Psi must be a GMPS of rank 1 (vector).");
  TN_CHECK(e->ket == psi && e->bra == psi, "dmrg: the environment must be built on psi");
  int N = psi->N, d = psi->d;
  double cost = 0;
  // small bonds: Theta0 and the eigsolve run in one single-CTA launch (tn_small.cu) that leaves the energy and the number of H_eff
  // applications in device memory (dscal slots 48 / 49) -- read back once, at the end of the half sweep
  double* energy_dev = reinterpret_cast<double*>(c->dscal + 48);
  int* numops_dev = reinterpret_cast<int*>(c->dscal + 49);
  TN_CUDA(cudaMemsetAsync(numops_dev, 0, sizeof(int), s));
  bool last_small = false, any_small = false;
  for (int j = 1; j <= N - 1; ++j) {
    int site = direction ? N + 1 - j : j;
    int site1 = direction ? site - 1 : site;
    env_movecenter(e, site);
    Tensor& A = psi->sites[site1 - 1]; Tensor& B = psi->sites[site1];
    int cl = (int)A.dims[0], cm = (int)A.dims[2], cr = (int)B.dims[2];
    long long n = (long long)cl * d * d * cr;
    cplx* th1 = c->scratch[14].get((size_t)n, s);
    last_small = lanczos_small(e, site1, energy_dev, numops_dev, th1, lz);
    if (last_small) any_small = true;
    else {
      cplx* th0 = c->scratch[13].get((size_t)n, s);
      zgemm_auto(mk(cl * d, d * cr, cm, A.p, idx1(1), idx1((long long)cl * d), 0, B.p, idx1(1), idx1(cm), 0, th0, idx1(1), idx1((long long)cl * d)), s);
      cost = lanczos_lowest(e, site1, th0, th1, n, lz, nullptr);
    }
    mps_replacesites2(psi, th1, site1, direction, true, tr);
  }
  env_movecenter(e, direction ? 1 : N);
  if (any_small) {
    TN_CUDA(cudaMemcpyAsync(c->hscal + 48, c->dscal + 48, 2 * sizeof(cplx), cudaMemcpyDeviceToHost, s));
    c->sync();
    if (last_small) cost = *reinterpret_cast<double*>(c->hscal + 48);
    c->matvecs += *reinterpret_cast<int*>(c->hscal + 49);
  }
  if (energy) *energy = cost;
  if (maxbond) *maxbond = psi->maxbonddim();
}

// ================================================================================================
// Gates (gatelist.jl)
// ================================================================================================
void gates_free(Gates* g);
Gates* gates_create(Ctx* c, int d, int nrows, const int* counts, const int* sites, const int* nsites, const cplx* const* host_gates) {
  auto g = std::make_unique<Gates>();
  g->ctx = c; g->d = d; g->rows.resize(nrows);
  int idx = 0;
  try {
    for (int r = 0; r < nrows; ++r)
      for (int k = 0; k < counts[r]; ++k, ++idx) {
        TN_CHECK(nsites[idx] == 1 || nsites[idx] == 2, "only one- and two-site gates are supported");
        size_t ne = 1; for (int q = 0; q < 2 * nsites[idx]; ++q) ne *= d;
        Gate gt{sites[idx], nsites[idx], nullptr};
        TN_CUDA(cudaMalloc((void**)&gt.dev, ne * sizeof(cplx)));
        g->rows[r].push_back(gt);                       // owned by the list from here on (freed below if a later gate fails)
        TN_CUDA(cudaMemcpyAsync(gt.dev, host_gates[idx], ne * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
      }
    c->sync();
  } catch (...) {
    gates_free(g.release());
    throw;
  }
  return g.release();
}
void gates_free(Gates* g) {
  if (!g) return;
  for (auto& r : g->rows) for (auto& gt : r) cudaFree(gt.dev);
  delete g;
}

// gatelist.jl:137-171.  fid (optional): multiplied by the fidelity |<Theta', Theta'_truncated>|^2 of a two-site gate (error=true, :160-168)
static void applygate(Mps* psi, const Gate& g, bool direction, Trunc tr, double* fid = nullptr) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  int d = psi->d, inner = psi->rank == 2 ? d : 1;
  TN_CHECK(g.site >= 1 && g.site + g.nsites - 1 <= psi->N, "gate site out of range for this MPS");
  if (g.nsites == 1) {
    mps_applyop1(psi, g.site, g.dev);
    // replacesites!, one-site branch (gmps.jl:204-213): move the centre one site on, untruncated
    int nxt = g.site + 1 - 2 * (direction ? 1 : 0);
    if (0 < nxt && nxt <= psi->N) mps_movecenter(psi, nxt, Trunc{0.0, 0, 1});
    return;
  }
  Tensor& A = psi->sites[g.site - 1]; Tensor& B = psi->sites[g.site];
  long long cl = A.dims.front(), cm = A.dims.back(), cr = B.dims.back(), p = psi->phys();
  long long n = cl * p * p * cr;
  cplx* th0 = c->scratch[13].get((size_t)n, s);
  cplx* th1 = c->scratch[14].get((size_t)n, s);
  zgemm_auto(mk((int)(cl * p), (int)(p * cr), (int)cm, A.p, idx1(1), idx1(cl * p), 0, B.p, idx1(1), idx1(cm), 0, th0, idx1(1), idx1(cl * p)), s);
  gate_mix2(th0, th1, g.dev, cl, d, inner, cr, s);
  mps_replacesites2(psi, th1, g.site, direction, false, tr);
  if (fid) {
    // contract the two new site tensors again and take the overlap with the untruncated Theta' (th1 is still intact)
    Tensor& A2 = psi->sites[g.site - 1]; Tensor& B2 = psi->sites[g.site];
    const long long k = A2.dims.back();
    zgemm_auto(mk((int)(cl * p), (int)(p * cr), (int)k, A2.p, idx1(1), idx1(cl * p), 0, B2.p, idx1(1), idx1(k), 0, th0, idx1(1), idx1(cl * p)), s);
    const cplx* xs[1] = {th1};
    zdots(n, 1, xs, th0, c->dscal + 40, c->partials, s);
    cplx v = read_scalar(c, 40);
    *fid *= v.x * v.x + v.y * v.y;
  }
}

void apply_gates(Mps* psi, Gates* g, Trunc tr, double* fid) {   // gatelist.jl:191-227 (fid != nullptr: applygates with error=true)
  if (fid) *fid = 1.0;
  TN_CHECK(g->d == psi->d, "gate / MPS physical dimension mismatch");
  for (auto& row : g->rows)      // validate the whole list before the first gate changes psi
    for (auto& gt : row) TN_CHECK(gt.site >= 1 && gt.site + gt.nsites - 1 <= psi->N, "gate site out of range for this MPS");
  for (auto& row : g->rows) {
    if (row.empty()) continue;
    int firstsite = row.front().site;
    int lastsite = row.back().site + row.back().nsites - 1;
    bool direction = !(std::abs(psi->center - firstsite) < std::abs(psi->center - lastsite));
    int n = (int)row.size();
    for (int i = 1; i <= n; ++i) {
      const Gate& gt = row[(direction ? n + 1 - i : i) - 1];
      int ctr = direction ? gt.site + gt.nsites - 1 : gt.site;
      mps_movecenter(psi, ctr, tr);
      applygate(psi, gt, direction, tr, fid);
    }
  }
}

// ================================================================================================
// Local expectation values <psi| O_site |psi> for single-site operators (mps.jl:87-134 restricted to
// one-site terms; qjmc.jl:170-220 uses it with O = L^dag L): overlap blocks built exactly as the
// reference does (ProjMPS(psi, psi)), no canonical-form shortcut.
// ================================================================================================
__global__ void zconj_kernel(long long n, cplx* x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i].y = -x[i].y;
}
void zconj_inplace(long long n, cplx* x, cudaStream_t s) {
  if (n <= 0) return;
  int blocks = (int)std::min<long long>(148 * 8, (n + 255) / 256);
  zconj_kernel<<<blocks, 256, 0, s>>>(n, x);
  count_launch(1);
}

void expect_local(Mps* psi, int nops, const int* sites, const cplx* ops_host, cplx* out_host) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  TN_CHECK(psi->rank == 1, "expect_local: rank-1 MPS only");
  int d = psi->d, N = psi->N;
  TN_CHECK(nops >= 0 && nops <= 1 << 20, "bad operator count");
  if (nops == 0) return;
  for (int k = 0; k < nops; ++k) TN_CHECK(sites[k] >= 1 && sites[k] <= N, "expect_local: operator site out of range");
  cplx* ops = c->scratch[15].get((size_t)nops * d * d + 64, s);
  TN_CUDA(cudaMemcpyAsync(ops, ops_host, (size_t)nops * d * d * sizeof(cplx), cudaMemcpyHostToDevice, s));
  Env* e = env_create(c, psi, nullptr, psi, ONE, 1);
  cplx* dall; TN_CUDA(cudaMallocAsync((void**)&dall, sizeof(cplx) * nops, s));
  TN_CUDA(cudaMemsetAsync(dall, 0, sizeof(cplx) * nops, s));
  for (int site = 1; site <= N; ++site) {
    bool any = false;
    for (int k = 0; k < nops; ++k) if (sites[k] == site) any = true;
    if (!any) continue;
    env_movecenter(e, site);
    const Tensor& L = env_block(e, site - 1);
    const Tensor& R = env_block(e, site + 1);
    const Tensor& A = psi->sites[site - 1];
    int cl = (int)A.dims[0], cr = (int)A.dims[2];
    for (int k = 0; k < nops; ++k) {
      if (sites[k] != site) continue;
      // OA = O A ; P1[a,(s,d)] = L[a,b] OA[b,(s,d)] ; P2[c,d] = sum_{(a,s)} conj(A[(a,s),c]) P1[(a,s),d] ; <..> = sum P2 .* R
      cplx* OA = c->scratch[1].get((size_t)A.size(), s);
      op_apply1(A.p, OA, ops + (size_t)k * d * d, cl, d, 1, cr, s);
      cplx* P1 = c->scratch[2].get((size_t)A.size(), s);
      zgemm_auto(mk(cl, d * cr, cl, L.p, idx1(1), idx1(cl), 0, OA, idx1(1), idx1(cl), 0, P1, idx1(1), idx1(cl)), s);
      cplx* P2 = c->scratch[4].get((size_t)cr * cr, s);
      zgemm_auto(mk(cr, cr, cl * d, A.p, idx1((long long)cl * d), idx1(1), 1, P1, idx1(1), idx1((long long)cl * d), 0, P2, idx1(1), idx1(cr)), s);
      zconj_inplace((long long)cr * cr, P2, s);
      const cplx* xs[1] = {P2};
      zdots((long long)cr * cr, 1, xs, R.p, dall + k, c->partials, s);
    }
  }
  TN_CUDA(cudaMemcpyAsync(out_host, dall, sizeof(cplx) * nops, cudaMemcpyDeviceToHost, s));
  c->sync();
  TN_CUDA(cudaFreeAsync(dall, s));
  env_free(e);
}

// ================================================================================================
// inner(st, psi, oplist, phi) (mps.jl:87-134): <psi| O_t |phi> for operator strings O_t = prod_k O_{t,k}(site_{t,k}),
// evaluated exactly as the reference does -- overlap blocks ProjMPS(psi, phi), the left block carried through the sites
// of the string, closed with the right block.  The energy measurement of tebd.jl:51,90 and the observers go through this.
// ================================================================================================
void inner_oplist(Mps* bra, Mps* ket, int nterms, const int* nops, const int* op_sites, const cplx* ops_host, const cplx* coeffs,
                  cplx* out_host) {
  Ctx* c = ket->ctx; cudaStream_t s = c->stream;
  TN_CHECK(bra->rank == 1 && ket->rank == 1, "inner: rank-1 MPS only");
  TN_CHECK(bra->N == ket->N && bra->d == ket->d, "inner: the MPSs must share length and physical dimension");
  TN_CHECK(nterms >= 0 && nterms <= (1 << 24), "inner: bad term count");
  if (nterms == 0) return;
  const int d = ket->d, N = ket->N;
  std::vector<long long> off(nterms + 1, 0);
  for (int t = 0; t < nterms; ++t) {
    TN_CHECK(nops[t] >= 1, "inner: every term needs at least one operator");
    off[t + 1] = off[t] + nops[t];
    for (long long k = off[t]; k < off[t + 1]; ++k) {
      TN_CHECK(op_sites[k] >= 1 && op_sites[k] <= N, "inner: operator site out of range");
      TN_CHECK(k == off[t] || op_sites[k] > op_sites[k - 1], "inner: operator sites of a term must be strictly ascending");
    }
  }
  const long long nop = off[nterms];
  cplx* ops = c->scratch[15].get((size_t)nop * d * d + 64, s);
  TN_CUDA(cudaMemcpyAsync(ops, ops_host, (size_t)nop * d * d * sizeof(cplx), cudaMemcpyHostToDevice, s));
  cplx* dall; TN_CUDA(cudaMallocAsync((void**)&dall, sizeof(cplx) * nterms, s));
  Env* e = env_create(c, bra, nullptr, ket, ONE, 1);
  for (int site = 1; site <= N; ++site) {
    bool moved = false;
    for (int t = 0; t < nterms; ++t) {
      if (op_sites[off[t]] != site) continue;
      if (!moved) { env_movecenter(e, site); moved = true; }
      const int last = op_sites[off[t + 1] - 1];
      const Tensor& L = env_block(e, site - 1);
      const Tensor& R = env_block(e, last + 1);
      const cplx* cur = L.p;                     // (chi_bra, chi_ket) overlap block left of the string
      long long k = off[t];
      int flip = 0;
      for (int q = site; q <= last; ++q) {
        const Tensor& A1 = bra->sites[q - 1];
        const Tensor& A2 = ket->sites[q - 1];
        const int ca = (int)A1.dims[0], ca2 = (int)A1.dims[2], cb = (int)A2.dims[0], cb2 = (int)A2.dims[2];
        const cplx* B = A2.p;
        if (k < off[t + 1] && op_sites[k] == q) {          // B = O A2 on the physical index (mps.jl:116-119)
          cplx* OA = c->scratch[1].get((size_t)A2.size(), s);
          op_apply1(A2.p, OA, ops + (size_t)k * d * d, cb, d, 1, cb2, s);
          B = OA; ++k;
        }
        // X1[a,(s,b')] = cur[a,b] B[b,(s,b')] ;  new[a',b'] = sum_{(a,s)} conj(A1[(a,s),a']) X1[(a,s),b']
        cplx* X1 = c->scratch[2].get((size_t)ca * d * cb2, s);
        zgemm_auto(mk(ca, d * cb2, cb, cur, idx1(1), idx1(ca), 0, B, idx1(1), idx1(cb), 0, X1, idx1(1), idx1(ca)), s);
        cplx* nxt = c->scratch[4 + flip].get((size_t)ca2 * cb2, s);
        zgemm_auto(mk(ca2, cb2, ca * d, A1.p, idx1((long long)ca * d), idx1(1), 1, X1, idx1(1), idx1((long long)ca * d), 0, nxt, idx1(1), idx1(ca2)), s);
        cur = nxt; flip ^= 1;
      }
      TN_CHECK(R.size() == (long long)bra->sites[last - 1].dims[2] * ket->sites[last - 1].dims[2], "inner: block size mismatch");
      // un-conjugated sum cur .* R: zdots conjugates its first argument, so conjugate a copy of cur first
      cplx* tmp = c->scratch[1].get((size_t)R.size(), s);
      TN_CUDA(cudaMemcpyAsync(tmp, cur, (size_t)R.size() * sizeof(cplx), cudaMemcpyDeviceToDevice, s));
      zconj_inplace(R.size(), tmp, s);
      const cplx* xs[1] = {tmp};
      zdots(R.size(), 1, xs, R.p, dall + t, c->partials, s);
    }
  }
  std::vector<cplx> raw(nterms);
  TN_CUDA(cudaMemcpyAsync(raw.data(), dall, sizeof(cplx) * nterms, cudaMemcpyDeviceToHost, s));
  c->sync();
  TN_CUDA(cudaFreeAsync(dall, s));
  env_free(e);
  for (int t = 0; t < nterms; ++t)
    out_host[t] = cplx{coeffs[t].x * raw[t].x - coeffs[t].y * raw[t].y, coeffs[t].x * raw[t].y + coeffs[t].y * raw[t].x};
}

}  // namespace tn
