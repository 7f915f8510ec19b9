// Sums of projections and the projector branch of the MPS hot path; everything is a composition of the strided ZGEMM
// (tn_zgemm.cu), the vector kernels (tn_vec.cu) and the drivers of tn_mps.cu -- no new contraction kernel.
// Reference call sites replaced (all under /root/reference/src):
//   structures/mps/projmpssum.jl:27-108  buildleft!/buildright!/movecenter!/product/project/calculate on a ProjMPSSum
//   structures/mps/projmps.jl:135-143    product, squared branch (rank-1 projector penalty |V><V| in psi's local basis)
//   structures/mps/projmps.jl:153-185    project (no MPO layer and one MPO layer; one and two sites)
//   structures/mps/projmps.jl:107-134    product, rank-2 branch with nsites = 1
//   structures/mps/gmps.jl:204-213       replacesites!, one-site branch
//   algorithms/mps/dmrg.jl:34-63,128-154 sweep body over a ProjMPSSum (several MPOs, MPS penalties), nsites = 1 or 2
//   algorithms/mps/vmps.jl:36-62         sweep body of the variational MPS optimisation
// The index wiring of every GEMM below was checked on the CPU with tools/gemm_emul.py + tools/proto_projsum.py.
#include "tn_mps.cuh"
#include <algorithm>
#include <cmath>

namespace tn {
void count_launch(int n);
void build_w2(const cplx* M1, const cplx* M2, cplx* W, int w, int w1, int w2, int d, cudaStream_t s);

static const cplx ONE = {1.0, 0.0};

Env* env_create_squared(Ctx* c, Mps* V, Mps* psi, cplx coeff, int center) {   // dmrg.jl:144-145
  Env* e = env_create(c, V, nullptr, psi, coeff, center);
  e->squared = true;
  return e;
}

static long long ipow(long long d, int n) { long long p = 1; for (int i = 0; i < n; ++i) p *= d; return p; }

// phi = conj(project(projV, A, direction, nsites)) for the sites site .. site+nsites-1 (A is not used by the reference):
//   phi(b, s1'.., b') = sum conj(L)(a,w,b) bra(a,s1,a') [bra(a',s2,a'')] conj(M1)(w,s1,s1',w1) [conj(M2)] conj(R)(a'',w2,b')
void env_project_phi(Env* e, int site, int nsites, cplx* phi) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  TN_CHECK(nsites == 1 || nsites == 2, "project: nsites must be 1 or 2");
  TN_CHECK(site >= 1 && site + nsites - 1 <= e->ket->N, "project: site out of range");
  const Tensor& L = env_block(e, site - 1);
  const Tensor& R = env_block(e, site + nsites);
  const Tensor& V1 = e->bra->sites[site - 1];
  const int d = e->ket->d, d2 = d * d;
  const int ca = (int)L.dims[0], w = (int)L.dims[1], cb = (int)L.dims[2];
  const int ca2 = (int)R.dims[0], w2 = (int)R.dims[1], cb2 = (int)R.dims[2];
  const int ca1 = (int)V1.dims[2];
  TN_CHECK(V1.dims[0] == ca, "project: block / bra bond mismatch");
  if (!e->mpo) {
    // X1[b,(s1,a')] = sum_a conj(L[a,b]) V1[a,(s1,a')]
    cplx* X1 = c->scratch[16].get((size_t)cb * d * ca1, s);
    zgemm_auto(mk(cb, d * ca1, ca, L.p, idx1(ca), idx1(1), 1, V1.p, idx1(1), idx1(ca), 0, X1, idx1(1), idx1(cb)), s);
    if (nsites == 1) {
      TN_CHECK(ca2 == ca1, "project: block / bra bond mismatch");
      // phi[(b,s),b'] = sum_a' X1[(b,s),a'] conj(R[a',b'])
      zgemm_auto(mk(cb * d, cb2, ca1, X1, idx1(1), idx1((long long)cb * d), 0, R.p, idx1(1), idx1(ca1), 1, phi, idx1(1), idx1((long long)cb * d)), s);
      return;
    }
    const Tensor& V2 = e->bra->sites[site];
    TN_CHECK(V2.dims[0] == ca1 && V2.dims[2] == ca2, "project: block / bra bond mismatch");
    // X2[(b,s1),(s2,a'')] = X1[(b,s1),a'] V2[a',(s2,a'')]
    cplx* X2 = c->scratch[17].get((size_t)cb * d2 * ca2, s);
    zgemm_auto(mk(cb * d, d * ca2, ca1, X1, idx1(1), idx1((long long)cb * d), 0, V2.p, idx1(1), idx1(ca1), 0, X2, idx1(1), idx1((long long)cb * d)), s);
    // phi[(b,s1,s2),b'] = sum_a'' X2[(b,s1,s2),a''] conj(R[a'',b'])
    zgemm_auto(mk(cb * d2, cb2, ca2, X2, idx1(1), idx1((long long)cb * d2), 0, R.p, idx1(1), idx1(ca2), 1, phi, idx1(1), idx1((long long)cb * d2)), s);
    return;
  }
  const Tensor& M1 = e->mpo->sites[site - 1];
  TN_CHECK(M1.dims[0] == w, "project: MPO / block bond mismatch");
  if (nsites == 1) {
    TN_CHECK(ca2 == ca1 && M1.dims[3] == w2, "project: block / site bond mismatch");
    // T1(b,w,s,a') = sum_a conj(L(a,w,b)) V1(a,s,a'); rows m = (w,b) in L's order, written transposed
    cplx* T1 = c->scratch[16].get((size_t)cb * w * d * ca1, s);
    zgemm_auto(mk(w * cb, d * ca1, ca, L.p, idx1(ca), idx1(1), 1, V1.p, idx1(1), idx1(ca), 0, T1, idx2(w, cb, 1), idx1((long long)w * cb)), s);
    // T2(b,s',w2,a') = sum_{(w,s)} T1(b,(w,s),a') conj(M(w,s,s',w2)); rows m = (b,a')
    cplx* T2 = c->scratch[17].get((size_t)cb * d * w2 * ca1, s);
    zgemm_auto(mk(cb * ca1, d * w2, w * d, T1, idx2(cb, 1, (long long)cb * w * d), idx1(cb), 0,
                  M1.p, idx1(1), idx1((long long)w * d), 1, T2, idx2(cb, 1, (long long)cb * d * w2), idx1(cb)), s);
    // phi[(b,s'),b'] = sum_{(w2,a')} T2[(b,s'),(w2,a')] conj(R(a',w2,b'))
    zgemm_auto(mk(cb * d, cb2, w2 * ca1, T2, idx1(1), idx1((long long)cb * d), 0, R.p, idx2(w2, ca1, 1), idx1((long long)ca1 * w2), 1,
                  phi, idx1(1), idx1((long long)cb * d)), s);
    return;
  }
  const Tensor& V2 = e->bra->sites[site];
  const Tensor& M2 = e->mpo->sites[site];
  const int w1 = (int)M1.dims[3];
  TN_CHECK(V2.dims[0] == ca1 && V2.dims[2] == ca2 && M2.dims[0] == w1 && M2.dims[3] == w2, "project: block / site bond mismatch");
  // W[(w,s1',s2'),(s1,s2,w2)] = sum_{w1} M1 M2 (same layout as the matvec's)
  cplx* W = c->scratch[20].get((size_t)w * d2 * d2 * w2, s);
  build_w2(M1.p, M2.p, W, w, w1, w2, d, s);
  // thb[(a,s1),(s2,a'')] = V1[(a,s1),a'] V2[a',(s2,a'')]
  cplx* thb = c->scratch[18].get((size_t)ca * d2 * ca2, s);
  zgemm_auto(mk(ca * d, d * ca2, ca1, V1.p, idx1(1), idx1((long long)ca * d), 0, V2.p, idx1(1), idx1(ca1), 0, thb, idx1(1), idx1((long long)ca * d)), s);
  // T1(b,w,s1,s2,a'') = sum_a conj(L(a,w,b)) thb(a,s1,s2,a'')
  cplx* T1 = c->scratch[16].get((size_t)cb * w * d2 * ca2, s);
  zgemm_auto(mk(w * cb, d2 * ca2, ca, L.p, idx1(ca), idx1(1), 1, thb, idx1(1), idx1(ca), 0, T1, idx2(w, cb, 1), idx1((long long)w * cb)), s);
  // T2(b,s1',s2',w2,a'') = sum_{(w,s1,s2)} T1(b,(w,s1,s2),a'') conj(W[(w,s1',s2'),(s1,s2,w2)]); k = (w,(s1,s2)), n = ((s1',s2'),w2)
  cplx* T2 = c->scratch[17].get((size_t)cb * d2 * w2 * ca2, s);
  zgemm_auto(mk(cb * ca2, d2 * w2, w * d2, T1, idx2(cb, 1, (long long)cb * w * d2), idx1(cb), 0,
                W, idx2(w, 1, (long long)w * d2), idx2(d2, w, (long long)w * d2 * d2), 1,
                T2, idx2(cb, 1, (long long)cb * d2 * w2), idx1(cb)), s);
  // phi[(b,s1',s2'),b'] = sum_{(w2,a'')} T2[(b,s1',s2'),(w2,a'')] conj(R(a'',w2,b'))
  zgemm_auto(mk(cb * d2, cb2, w2 * ca2, T2, idx1(1), idx1((long long)cb * d2), 0, R.p, idx2(w2, ca2, 1), idx1((long long)ca2 * w2), 1,
                phi, idx1(1), idx1((long long)cb * d2)), s);
}

// One-site H_eff: out(a,s,a') = coeff * sum L(a,w,b) M(w,s,s',w') A(b,s',b') R(a',w',b')   (projmps.jl:107-134 with nsites = 1)
void env_product1_dev(Env* e, const cplx* A, int site, cplx* out) {
  Ctx* c = e->ctx; cudaStream_t s = c->stream;
  TN_CHECK(e->mpo != nullptr, "product: the rank-2 branch needs an MPO layer");
  TN_CHECK(site >= 1 && site <= e->ket->N, "product: site out of range");
  const Tensor& L = env_block(e, site - 1);
  const Tensor& R = env_block(e, site + 1);
  const Tensor& M = e->mpo->sites[site - 1];
  const int d = e->ket->d;
  const int ca = (int)L.dims[0], w = (int)L.dims[1], cb = (int)L.dims[2];
  const int ca2 = (int)R.dims[0], w2 = (int)R.dims[1], cb2 = (int)R.dims[2];
  TN_CHECK(M.dims[0] == w && M.dims[3] == w2, "product: MPO / block bond mismatch");
  // T1[(a,w),(s',b')] = L[(a,w),b] A[b,(s',b')]
  cplx* T1 = c->scratch[4].get((size_t)ca * w * d * cb2, s);
  zgemm_auto(mk(ca * w, d * cb2, cb, L.p, idx1(1), idx1((long long)ca * w), 0, A, idx1(1), idx1(cb), 0, T1, idx1(1), idx1((long long)ca * w)), s);
  // T2(a,s,w',b') = sum_{(w,s')} T1(a,(w,s'),b') M(w,s,s',w'); rows m = (a,b')
  cplx* T2 = c->scratch[5].get((size_t)ca * d * w2 * cb2, s);
  zgemm_auto(mk(ca * cb2, d * w2, w * d, T1, idx2(ca, 1, (long long)ca * w * d), idx1(ca), 0,
                M.p, idx2(w, 1, (long long)w * d), idx2(d, w, (long long)w * d * d), 0,
                T2, idx2(ca, 1, (long long)ca * d * w2), idx1(ca)), s);
  // out[(a,s),a'] = coeff * sum_{(w',b')} T2[(a,s),(w',b')] R[a',(w',b')]
  zgemm_auto(mk(ca * d, ca2, w2 * cb2, T2, idx1(1), idx1((long long)ca * d), 0, R.p, idx1(ca2), idx1(1), 0,
                out, idx1(1), idx1((long long)ca * d), e->coeff), s);
  c->matvecs++;
}

// replacesites!(psi, A, site, direction, normalize) with a one-site tensor (gmps.jl:204-213): psi[site] = A, then the
// centre moves one site on in the sweep direction by an untruncated SVD gauge move.  A has the site's current dimensions.
void mps_replacesite1(Mps* m, const cplx* A, int site, bool direction, bool normalize) {
  Ctx* c = m->ctx;
  TN_CHECK(site >= 1 && site <= m->N, "replacesites: site out of range");
  Tensor& T = m->sites[site - 1];
  if (A != T.p) TN_CUDA(cudaMemcpyAsync(T.p, A, (size_t)T.size() * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
  int nxt = site + 1 - 2 * (direction ? 1 : 0);
  if (0 < nxt && nxt <= m->N) mps_movecenter(m, nxt, Trunc{0.0, 0, 1});
  if (normalize) mps_normalize(m);
}

// ================================================================================================
// ProjMPSSum
// ================================================================================================
EnvSum* envsum_create(Ctx* c, int n, Env* const* projs, int center) {   // projmpssum.jl:11-19
  TN_CHECK(n >= 1 && n <= 16, "ProjMPSSum: between 1 and 16 projections");
  auto es = std::make_unique<EnvSum>();
  es->ctx = c; es->center = 0;
  for (int i = 0; i < n; ++i) {
    TN_CHECK(projs[i] != nullptr, "ProjMPSSum: null projection");
    TN_CHECK(projs[i]->ctx == c, "ProjMPSSum: every projection must live in the same context");
    TN_CHECK(projs[i]->ket == projs[0]->ket, "ProjMPSSum: every projection must share the ket MPS");
    es->projs.push_back(projs[i]);
  }
  EnvSum* raw = es.release();
  try { envsum_movecenter(raw, center); } catch (...) { delete raw; throw; }
  return raw;
}
void envsum_free(EnvSum* es) { delete es; }   // the member projections are owned by their own handles
void envsum_movecenter(EnvSum* es, int idx) {   // projmpssum.jl:51-55
  for (Env* e : es->projs) env_movecenter(e, idx);
  es->center = idx;
}
cplx envsum_calculate(EnvSum* es) {   // projmpssum.jl:97-108
  cplx tot = {0.0, 0.0};
  for (Env* e : es->projs) { cplx v = env_calculate(e); tot.x += v.x; tot.y += v.y; }
  return tot;
}

static long long local_size(Env* e, int site, int nsites) {
  TN_CHECK(site >= 1 && site + nsites - 1 <= e->ket->N, "product: the sites fall outside the chain");
  return e->ket->chiL(site) * ipow(e->ket->d, nsites) * e->ket->chiR(site + nsites - 1);
}
static int count_mpo(EnvSum* es) { int n = 0; for (Env* e : es->projs) if (!e->squared) ++n; return n; }

// Once per bond: phi of every squared projection, W = M1.M2 of the MPO layer when there is exactly one.
void envsum_prepare(EnvSum* es, int site, int nsites) {
  Ctx* c = es->ctx;
  TN_CHECK(nsites == 1 || nsites == 2, "product: nsites must be 1 or 2");
  for (Env* e : es->projs) {
    if (e->squared) {
      const Tensor& L = env_block(e, site - 1);
      const Tensor& R = env_block(e, site + nsites);
      std::vector<long long> dd;
      dd.push_back(L.dims[2]);
      for (int i = 0; i < nsites; ++i) dd.push_back(e->ket->d);
      dd.push_back(R.dims[2]);
      c->alloc(e->phi, dd);
      env_project_phi(e, site, nsites, e->phi.p);
    } else {
      TN_CHECK(e->mpo != nullptr, "product: a rank-1 projection without squared=true has no product on a site tensor");
    }
  }
  if (nsites == 2 && count_mpo(es) == 1)
    for (Env* e : es->projs) if (!e->squared) heff_prepare(e, site);
}

// out = sum_k product(projs[k], in, direction, nsites)  (projmpssum.jl:63-73); envsum_prepare must have run for this bond.
void envsum_apply(EnvSum* es, const cplx* in, int site, int nsites, cplx* out) {
  Ctx* c = es->ctx; cudaStream_t s = c->stream;
  Env* e0 = es->projs[0];
  const long long n = local_size(e0, site, nsites);
  const bool prepared = nsites == 2 && count_mpo(es) == 1;
  bool first = true;
  for (Env* e : es->projs) {
    if (e->squared) continue;
    cplx* dst = first ? out : c->scratch[19].get((size_t)n, s);
    if (nsites == 2) env_product_dev(e, in, site, dst, nullptr, prepared);
    else env_product1_dev(e, in, site, dst);
    if (!first) zaxpy(n, ONE, dst, out, s);
    first = false;
  }
  int sq = 0;
  for (Env* e : es->projs) {
    if (!e->squared) continue;
    TN_CHECK(e->phi.p != nullptr && e->phi.size() == n, "product: the projection has not been prepared for these sites");
    if (first) { TN_CUDA(cudaMemsetAsync(out, 0, (size_t)n * sizeof(cplx), s)); first = false; }
    // out += coeff * <phi, in> * phi
    const cplx* xs[1] = {e->phi.p};
    zdots(n, 1, xs, in, c->dscal + 32 + sq, c->partials, s);
    zaxpy_dev(n, e->coeff, c->dscal + 32 + sq, e->phi.p, out, s);
    ++sq;
  }
}

// out = conj(sum_k project(projs[k], ., direction, nsites))  (projmpssum.jl:81-91; no coefficients, like the reference)
void envsum_project_phi(EnvSum* es, int site, int nsites, cplx* out) {
  Ctx* c = es->ctx; cudaStream_t s = c->stream;
  const long long n = local_size(es->projs[0], site, nsites);
  bool first = true;
  for (Env* e : es->projs) {
    const Tensor& L = env_block(e, site - 1);
    const Tensor& R = env_block(e, site + nsites);
    TN_CHECK(L.dims[2] * ipow(e->ket->d, nsites) * R.dims[2] == n, "project: block / ket bond mismatch");
    cplx* dst = first ? out : c->scratch[19].get((size_t)n, s);
    env_project_phi(e, site, nsites, dst);
    if (!first) zaxpy(n, ONE, dst, out, s);
    first = false;
  }
}

// ================================================================================================
// Sweeps
// ================================================================================================
static void sweep_sites(int N, int nsites, bool direction, int j, int* site, int* site1) {   // dmrg.jl:36-38, vmps.jl:38-40
  *site = direction ? N + 1 - j : j;
  *site1 = direction ? *site + 1 - nsites : *site;
}

void dmrg_halfsweep_sum(Mps* psi, EnvSum* es, bool direction, int nsites, Lanczos lz, Trunc tr, double* energy, long long* maxbond) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  TN_CHECK(psi->rank == 1, "Psi must be a GMPS of rank 1 (vector).");
  TN_CHECK(nsites == 1 || nsites == 2, "dmrg: nsites must be 1 or 2");
  for (Env* e : es->projs) {
    TN_CHECK(e->ket == psi, "dmrg: every projection must be built on psi");
    TN_CHECK(e->squared || (e->bra == psi && e->mpo != nullptr), "dmrg: the Hamiltonian must be composed of MPOs or squared MPS projections");
  }
  const int N = psi->N, d = psi->d;
  double cost = 0;
  for (int j = 1; j <= N + 1 - nsites; ++j) {
    int site, site1; sweep_sites(N, nsites, direction, j, &site, &site1);
    envsum_movecenter(es, site);
    long long n;
    const cplx* th0;
    if (nsites == 2) {
      Tensor& A = psi->sites[site1 - 1]; Tensor& B = psi->sites[site1];
      int cl = (int)A.dims[0], cm = (int)A.dims[2], cr = (int)B.dims[2];
      n = (long long)cl * d * d * cr;
      cplx* t = c->scratch[13].get((size_t)n, s);
      zgemm_auto(mk(cl * d, d * cr, cm, A.p, idx1(1), idx1((long long)cl * d), 0, B.p, idx1(1), idx1(cm), 0, t, idx1(1), idx1((long long)cl * d)), s);
      th0 = t;
    } else {
      n = psi->sites[site1 - 1].size();
      th0 = psi->sites[site1 - 1].p;
    }
    cplx* th1 = c->scratch[14].get((size_t)n, s);
    envsum_prepare(es, site1, nsites);
    cost = lanczos_core(c, [&](const cplx* in, cplx* out) { envsum_apply(es, in, site1, nsites, out); }, th0, th1, n, lz, nullptr);
    if (nsites == 2) mps_replacesites2(psi, th1, site1, direction, true, tr);
    else mps_replacesite1(psi, th1, site1, direction, true);
  }
  envsum_movecenter(es, direction ? 1 : N);
  if (energy) *energy = cost;
  if (maxbond) *maxbond = psi->maxbonddim();
}

void vmps_halfsweep(Mps* psi, EnvSum* es, bool direction, int nsites, Trunc tr, long long* maxbond) {
  Ctx* c = psi->ctx; cudaStream_t s = c->stream;
  TN_CHECK(psi->rank == 1, "vmps: rank-1 MPS only");
  TN_CHECK(nsites == 1 || nsites == 2, "vmps: nsites must be 1 or 2");
  for (Env* e : es->projs) TN_CHECK(e->ket == psi && !e->squared, "vmps: every projection must be ProjMPS(psi_k, [H,] psi)");
  const int N = psi->N;
  for (int j = 1; j <= N + 1 - nsites; ++j) {
    int site, site1; sweep_sites(N, nsites, direction, j, &site, &site1);
    envsum_movecenter(es, site);
    const long long n = local_size(es->projs[0], site1, nsites);
    cplx* vec = c->scratch[14].get((size_t)n, s);
    envsum_project_phi(es, site1, nsites, vec);          // vec = conj(project(Vs, A0, direction, nsites)): vmps.jl:51
    if (nsites == 2) mps_replacesites2(psi, vec, site1, direction, false, tr);
    else mps_replacesite1(psi, vec, site1, direction, false);
  }
  envsum_movecenter(es, direction ? 1 : N);
  if (maxbond) *maxbond = psi->maxbonddim();
}

}  // namespace tn
