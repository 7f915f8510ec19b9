// Infinite TEBD step in Vidal form on the device, two-site unit cell: reference algorithms/mps/itebd.jl:71-119
// (`_itebd_apply_gates_mps!`) on an iGMPS (structures/mps/igmps.jl:8-33).  Same building blocks as the finite gate step
// (Theta GEMM, gate_mix2, truncated SVD, gathers) plus the diagonal singular-value scalings, which are an elementwise kernel
// (scale_lr) instead of the reference's diagm() + contract.  singulars[i] sits to the left of tensors[i]; the cell is periodic.
#include "tn_mps.cuh"
#include <cmath>

namespace tn {

IMps* imps_create(Ctx* c, int d, int L, const long long* dims, const cplx* const* host_sites, const double* const* host_sing) {
  TN_CHECK(L == 2, "iTEBD on the device supports a two-site unit cell");
  TN_CHECK(d >= 2 && d <= 4, "physical dimension must be 2, 3 or 4");
  auto m = std::make_unique<IMps>();
  m->ctx = c; m->d = d; m->L = L;
  m->gam.resize(L); m->sing.assign(L, nullptr); m->nsing.assign(L, 0); m->norms.assign(L, 0.0);
  for (int i = 0; i < L; ++i) {
    std::vector<long long> dd(dims + 3 * i, dims + 3 * (i + 1));
    TN_CHECK(dd[1] == d, "physical dimension mismatch");
    TN_CHECK(dd[2] == dims[3 * ((i + 1) % L)], "bond dimensions of neighbouring cell sites differ");
    c->alloc(m->gam[i], dd);
    TN_CUDA(cudaMemcpyAsync(m->gam[i].p, host_sites[i], (size_t)m->gam[i].size() * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
    m->nsing[i] = dd[0];
    TN_CUDA(cudaMalloc((void**)&m->sing[i], (size_t)dd[0] * sizeof(double)));
    TN_CUDA(cudaMemcpyAsync(m->sing[i], host_sing[i], (size_t)dd[0] * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  c->sync();
  return m.release();
}
void imps_free(IMps* m) {
  if (!m) return;
  for (auto& t : m->gam) m->ctx->free(t);
  for (double* p : m->sing) if (p) cudaFree(p);
  delete m;
}

// One pass over the cell (itebd.jl:73-118 for length(psi) == 2): for (a, b) = (1, 2), (2, 1):
//   Theta = S_a G_a S_b G_b S_a -> gate -> SVD -> G_a = S_a^-1 U, S_b = s / |s|, G_b = V^H S_a^-1, norms[b] += log |s|
void itebd_apply_gate2(IMps* m, const cplx* gate_dev, Trunc tr) {
  Ctx* c = m->ctx; cudaStream_t s = c->stream;
  const int d = m->d;
  for (int i = 0; i < 2; ++i) {
    const int a = i, b = 1 - i;
    Tensor& Ga = m->gam[a]; Tensor& Gb = m->gam[b];
    const long long Da = Ga.dims[0], Dm = Ga.dims[2], Dr = Gb.dims[2];
    TN_CHECK(Gb.dims[0] == Dm && Dr == Da && m->nsing[a] == Da && m->nsing[b] == Dm, "iTEBD: inconsistent cell dimensions");
    // X = S_a G_a S_b,  Y = G_b S_a   (itebd.jl:78-83)
    cplx* X = c->scratch[1].get((size_t)Ga.size(), s);
    cplx* Y = c->scratch[2].get((size_t)Gb.size(), s);
    scale_lr(Ga.p, X, Da, d, Dm, m->sing[a], false, m->sing[b], false, s);
    scale_lr(Gb.p, Y, Dm, d, Dr, nullptr, false, m->sing[a], false, s);
    const long long n = Da * d * d * Dr;
    cplx* th0 = c->scratch[13].get((size_t)n, s);
    cplx* th1 = c->scratch[14].get((size_t)n, s);
    zgemm_auto(mk((int)(Da * d), (int)(d * Dr), (int)Dm, X, idx1(1), idx1(Da * d), 0, Y, idx1(1), idx1(Dm), 0, th0, idx1(1), idx1(Da * d)), s);
    gate_mix2(th0, th1, gate_dev, Da, d, 1, Dr, s);                       // itebd.jl:86-87
    const long long rows = Da * d, cols = d * Dr;
    const int k = svd_factor(c->svd, th1, (int)rows, (int)cols, rows, tr, s); c->svds++;     // itebd.jl:92-94
    // U and V^H without singular values, then the inverse of S_a on the outer bonds (itebd.jl:101-102)
    cplx* U = c->scratch[1].get((size_t)rows * k, s);
    cplx* Vh = c->scratch[2].get((size_t)k * cols, s);
    svd_gather_U(c->svd, U, rows, false, s);
    svd_gather_Vh(c->svd, Vh, k, false, s);
    Tensor Na, Nb;
    c->alloc(Na, {Da, (long long)d, (long long)k});
    c->alloc(Nb, {(long long)k, (long long)d, Dr});
    scale_lr(U, Na.p, Da, d, k, m->sing[a], true, nullptr, false, s);
    scale_lr(Vh, Nb.p, k, d, Dr, nullptr, false, m->sing[a], true, s);
    // new singular values of the inner bond, normalised; the log of their norm is accumulated (itebd.jl:105-109)
    std::vector<double> sv(k);
    double* dsv = reinterpret_cast<double*>(c->scratch[0].get((size_t)k / 2 + 1, s));
    svd_copy_S(c->svd, dsv, s);
    TN_CUDA(cudaMemcpyAsync(sv.data(), dsv, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s));
    c->sync();
    double nrm2 = 0; for (double x : sv) nrm2 += x * x;
    const double nrm = std::sqrt(nrm2);
    m->norms[b] += std::log(nrm);
    for (double& x : sv) x /= nrm;
    if (m->sing[b]) cudaFree(m->sing[b]);
    TN_CUDA(cudaMalloc((void**)&m->sing[b], (size_t)k * sizeof(double)));
    TN_CUDA(cudaMemcpyAsync(m->sing[b], sv.data(), (size_t)k * sizeof(double), cudaMemcpyHostToDevice, s));
    c->sync();                                                            // sv is a host temporary
    m->nsing[b] = k;
    c->free(Ga); c->free(Gb);
    m->gam[a] = Na; m->gam[b] = Nb;
  }
}

}  // namespace tn
