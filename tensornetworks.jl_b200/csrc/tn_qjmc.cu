// One quantum-jump Monte Carlo trajectory on the device: reference algorithms/mps/qjmc.jl:59-164
// (classical = true branch :88-112, and the norm-based branch classical = false :65-87) with the emission rates of :170-220.
// Uniforms are indexed (step, slot): classical mode draws slot 0 (unused, :64), slot 1 (jump test), slot 2 (channel);
// the norm-based mode draws slot 0 (jump test against the decayed norm^2) and slot 1 (channel).
#include "tn_mps.cuh"
#include <cmath>
#include <algorithm>

namespace tn {

// Counter-based generator for throughput runs: Philox4x32-10 (Salmon et al., SC'11; the generator behind cuRAND's Philox and
// Random123), key = seed, counter = (trajectory, step, slot).  Every (seed, trajectory, step, slot) has its own uniform, independent of how
// trajectories are dealt to GPUs, worker threads or batching rounds.  Known-answer vectors: tests/test_cabi.py.
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in [0, 1) with 53 random bits; step < 2^30 and slot < 4 share one counter word
double counter_uniform(uint64_t seed, uint64_t traj, uint64_t step, uint64_t slot) {
  const uint32_t ctr[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)step, (uint32_t)((step >> 32) << 2 | (slot & 3))};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  philox4x32_10(ctr, key, o);
  const uint64_t bits = ((uint64_t)o[0] << 21) ^ ((uint64_t)o[1] >> 11);      // 32 + 21 = 53 bits
  return (double)bits * (1.0 / 9007199254740992.0);
}

int qjmc_run(Mps* psi, Gates* gates, int njump, const int* jump_sites, const cplx* jump_ops, const double* jump_coeffs,
             int steps, double dt, Trunc tr, const double* uniforms, uint64_t seed, uint64_t traj,
             const cplx* obs_op, int save_every, cplx* obs_out, int* jumps_out, double* jumptimes_out, int jump_cap, bool classical) {
  Ctx* c = psi->ctx;
  int d = psi->d, N = psi->N;
  TN_CHECK(psi->rank == 1, "qjmc: psi must be an MPS");
  TN_CHECK(njump >= 0, "qjmc: negative jump-operator count");
  for (int k = 0; k < njump; ++k) TN_CHECK(jump_sites[k] >= 1 && jump_sites[k] <= N, "qjmc: jump-operator site out of range");
  // escape operators L^dag L (qjmc.jl:11-20) and device copies of the jump operators
  std::vector<cplx> esc((size_t)njump * d * d);
  for (int k = 0; k < njump; ++k)
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        double xr = 0, xi = 0;
        for (int q = 0; q < d; ++q) {
          cplx a = jump_ops[(size_t)k * d * d + q + d * i];   // L(q,i), conj -> L^dag(i,q)
          cplx b = jump_ops[(size_t)k * d * d + q + d * j];   // L(q,j)
          xr += a.x * b.x + a.y * b.y; xi += a.x * b.y - a.y * b.x;
        }
        esc[(size_t)k * d * d + i + d * j] = cplx{xr, xi};
      }
  struct DevBuf { cplx* p = nullptr; ~DevBuf() { if (p) cudaFree(p); } } djump_holder;     // released on every exit path
  TN_CUDA(cudaMalloc((void**)&djump_holder.p, sizeof(cplx) * (size_t)std::max(njump, 1) * d * d));
  cplx* djump = djump_holder.p;
  if (njump > 0) TN_CUDA(cudaMemcpyAsync(djump, jump_ops, sizeof(cplx) * (size_t)njump * d * d, cudaMemcpyHostToDevice, c->stream));
  std::vector<cplx> ex(njump);
  std::vector<double> rates(njump);
  std::vector<int> all_sites(N);
  std::vector<cplx> obs_ops;
  if (obs_op) { for (int i = 0; i < N; ++i) { all_sites[i] = i + 1; for (int q = 0; q < d * d; ++q) obs_ops.push_back(obs_op[q]); } }
  int njumps = 0;
  double time = 0;
  auto draw = [&](int step, int slot) { return uniforms ? uniforms[(size_t)3 * step + slot] : counter_uniform(seed, traj, step, slot); };
  auto emission_rates = [&]() {                                    // qjmc.jl:170-220 on the normalised state
    if (njump == 0) return 0.0;                                    // no channels: rates is empty, the trajectory never jumps
    expect_local(psi, njump, jump_sites, esc.data(), ex.data());
    double er = 0;
    for (int k = 0; k < njump; ++k) {
      double c2 = jump_coeffs[k] * jump_coeffs[k];
      rates[k] = std::hypot(c2 * ex[k].x, c2 * ex[k].y);
      er += rates[k];
    }
    return er;
  };
  for (int i = 1; i <= steps; ++i) {
    apply_gates(psi, gates, tr);                                   // qjmc.jl:61
    const double r0 = draw(i - 1, 0);                              // qjmc.jl:64 (drawn, unused in classical mode)
    bool jump;
    double er = 0;
    if (classical) {
      mps_normalize(psi);                                          // qjmc.jl:90
      er = emission_rates();                                       // qjmc.jl:93
      jump = draw(i - 1, 1) > std::exp(-er * dt);                  // qjmc.jl:95-97
    } else {
      cplx nrm = mps_norm(psi);                                    // qjmc.jl:67: prob = real(norm(psi)^2)
      const double prob = nrm.x * nrm.x - nrm.y * nrm.y;
      mps_normalize(psi);                                          // qjmc.jl:68
      jump = r0 > prob;                                            // qjmc.jl:69
      if (jump) er = emission_rates();                             // qjmc.jl:71
    }
    if (jump && njump > 0) {
      double r = draw(i - 1, classical ? 2 : 1), cum = 0;          // qjmc.jl:74 / :99
      int idx = njump - 1;
      for (int k = 0; k < njump; ++k) { cum += rates[k]; if (r < cum / er) { idx = k; break; } }
      mps_movecenter(psi, 1, Trunc{0.0, 0, 1});                    // qjmc.jl:103-107
      mps_applyop1(psi, jump_sites[idx], djump + (size_t)idx * d * d);
      mps_movecenter(psi, N, Trunc{0.0, 0, 1});
      mps_movecenter(psi, 1, tr);
      mps_normalize(psi);
      if (njumps < jump_cap) { if (jumps_out) jumps_out[njumps] = idx + 1; if (jumptimes_out) jumptimes_out[njumps] = time + dt; }
      njumps++;
    }
    time += dt;
    if (obs_op && save_every > 0 && i % save_every == 0)
      expect_local(psi, N, all_sites.data(), obs_ops.data(), obs_out + (size_t)(i / save_every - 1) * N);
  }
  c->sync();
  return njumps;
}

}  // namespace tn
