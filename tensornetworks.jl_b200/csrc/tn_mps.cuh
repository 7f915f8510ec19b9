// Device-resident MPS / MPO / environment containers and the hot-path algorithms built on the
// strided ZGEMM (tn_zgemm.cu), the Jacobi SVD (tn_svd.cu) and the vector kernels (tn_vec.cu).
// Site numbers are 1-based like the reference.
#pragma once
#include "tn_common.cuh"
#include "tn_svd.cuh"
#include "tn_vec.cuh"
#include <functional>
#include <memory>

namespace tn {

struct Buf {
  cplx* p = nullptr; size_t cap = 0;
  cplx* get(size_t n, cudaStream_t s);
  void release();
};

struct Tensor {           // column-major, first index fastest
  cplx* p = nullptr; size_t cap = 0;
  std::vector<long long> dims;
  long long size() const { long long n = 1; for (auto d : dims) n *= d; return n; }
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  SvdWork svd;
  SvdBatch svdb;             // batched factorisations (same-shape problems of lockstep trajectories)
  Buf scratch[24];           // 0-15: MPS / environment / Lanczos work space (tn_mps.cu); 16-23: projector sums (tn_projsum.cu)
  cplx* dscal = nullptr;     // 64 device scalars
  cplx* hscal = nullptr;     // 64 pinned host scalars
  cplx* partials = nullptr;  // dot-product partial sums
  long long matvecs = 0, svds = 0;
  cudaStream_t copy_stream = nullptr;   // H2D / D2H pipeline of the host-buffer matvec
  static constexpr int MAX_CHUNKS = 16;  // slices of the pipelined host-buffer matvec
  cudaEvent_t copy_ev[2 * MAX_CHUNKS + 1] = {};
  void alloc(Tensor& t, const std::vector<long long>& dims);
  void free(Tensor& t);
  void swap(Tensor& a, Tensor& b) { std::swap(a, b); }
  void sync() { TN_CUDA(cudaStreamSynchronize(stream)); }
};

struct Mps {
  Ctx* ctx; int rank, d, N, center;
  std::vector<Tensor> sites;     // sites[i-1] = (chi_l, d [, d], chi_r)
  long long chiL(int i) const { return sites[i - 1].dims.front(); }
  long long chiR(int i) const { return sites[i - 1].dims.back(); }
  long long phys() const { long long p = 1; for (int k = 0; k < rank; ++k) p *= d; return p; }
  long long maxbonddim() const;
};

struct Env {                     // ProjMPS(bra, mpo, ket; rank=2)  or overlap ProjMPS(bra, ket) when mpo == nullptr
  Ctx* ctx; Mps* bra; Mps* mpo; Mps* ket;
  std::vector<Tensor> blocks;    // (chi_bra, w, chi_ket)
  Tensor edge;                   // ones(1,1,1)
  int center; cplx coeff;
  bool squared = false;          // ProjMPS(V, psi; rank=2, squared=true): product = coeff * phi <phi, A> (projmps.jl:135-143)
  Tensor phi;                    // conj(project(projV, ...)) of the sites being optimised (tn_projsum.cu)
};

struct EnvSum {                  // ProjMPSSum (projmpssum.jl:1-4): every member shares the ket MPS
  Ctx* ctx; std::vector<Env*> projs; int center;
};

struct Gate { int site, nsites; cplx* dev; };
struct Gates { Ctx* ctx; int d; std::vector<std::vector<Gate>> rows; };

struct Lanczos { int krylovdim, maxiter; double tol; };
struct Env;
// tn_small.cu: Theta0 + the whole Lanczos eigsolve of one bond in a single-CTA launch when everything fits in shared memory (small
// bond dimensions); energy and the count of H_eff applications stay on the device.  false = does not fit, use the general path.
bool lanczos_small(Env* e, int site, double* energy_dev, int* numops_dev, cplx* theta_out, Lanczos lz);

// GEMM descriptor for one (unbatched, unsplit) strided contraction
inline GemmDesc mk(int M, int N, int K, const cplx* A, Idx2 am, Idx2 ak, int conjA, const cplx* B, Idx2 bk, Idx2 bn, int conjB,
                   cplx* C, Idx2 cm, Idx2 cn, cplx alpha = cplx{1.0, 0.0}, cplx beta = cplx{0.0, 0.0}) {
  GemmDesc g{};
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.am = am; g.ak = ak; g.conjA = conjA;
  g.B = B; g.bk = bk; g.bn = bn; g.conjB = conjB;
  g.C = C; g.cm = cm; g.cn = cn; g.alpha = alpha; g.beta = beta;
  g.batch = 1; g.ksplit = 1; g.kchunk = K;
  return g;
}

// --- MPS -------------------------------------------------------------------------------------
Mps* mps_create(Ctx* c, int rank, int d, int N, const long long* dims, const cplx* const* host_sites, int center);
void mps_free(Mps* m);
void mps_download_site(Mps* m, int i, cplx* host);
void mps_upload_site(Mps* m, int i, const long long* dims, const cplx* host);
cplx mps_norm(Mps* m);
void mps_normalize(Mps* m);
void mps_movecenter(Mps* m, int idx, Trunc tr);
void mps_replacesites2(Mps* m, const cplx* theta, int site, bool direction, bool normalize, Trunc tr);
void mps_replacesites2_factored(Mps* m, int site, bool direction, bool normalize);
void mps_applyop1(Mps* m, int site, const cplx* op_dev);
void mps_bond_spectrum(Mps* m, int site, std::vector<double>& out);
void mpo_compress(Mps* m, Trunc tr);   // mpo.jl:443-457
Mps* mpo_apply(Mps* O, Mps* psi, Trunc tr);   // applyMPO(O, psi): mpo.jl:105-143

// --- environments -----------------------------------------------------------------------------
Env* env_create(Ctx* c, Mps* bra, Mps* mpo, Mps* ket, cplx coeff, int center);
void env_free(Env* e);
void env_buildleft(Env* e, int idx);
void env_buildright(Env* e, int idx);
void env_movecenter(Env* e, int idx);
const Tensor& env_block(Env* e, int idx);
void env_product_dev(Env* e, const cplx* theta, int site, cplx* out, cudaEvent_t* ev4 = nullptr, bool prepared = false);   // sites (site, site+1)
void env_product_host(Env* e, const cplx* theta_host, int site, cplx* out_host);   // pipelined PCIe copies
cplx env_calculate(Env* e);

// --- drivers ------------------------------------------------------------------------------------
double lanczos_lowest(Env* e, int site, const cplx* theta0, cplx* theta_out, long long n, Lanczos lz, int* numops);
// the same eigensolver for any linear map out = H(in) on n-element device vectors (launched on c->stream)
typedef std::function<void(const cplx* in, cplx* out)> ApplyFn;
double lanczos_core(Ctx* c, const ApplyFn& apply, const cplx* theta0, cplx* theta_out, long long n, Lanczos lz, int* numops);
cplx read_scalar(Ctx* c, int slot);
void heff_prepare(Env* e, int site);
void zconj_inplace(long long n, cplx* x, cudaStream_t s);
void dmrg_halfsweep(Mps* psi, Env* e, bool direction, Lanczos lz, Trunc tr, double* energy, long long* maxbond);
Gates* gates_create(Ctx* c, int d, int nrows, const int* counts, const int* sites, const int* nsites, const cplx* const* host_gates);
void gates_free(Gates* g);
void apply_gates(Mps* psi, Gates* g, Trunc tr, double* fid = nullptr);   // fid: product of the gates' truncation fidelities (error=true)
void expect_local(Mps* psi, int nops, const int* sites, const cplx* ops_host, cplx* out_host);
void inner_oplist(Mps* bra, Mps* ket, int nterms, const int* nops, const int* op_sites, const cplx* ops_host, const cplx* coeffs,
                  cplx* out_host);

// --- projector sums, squared projectors, project(), one-site branch, vmps (tn_projsum.cu) ---------------------------
Env* env_create_squared(Ctx* c, Mps* V, Mps* psi, cplx coeff, int center);               // ProjMPS(V, psi; rank=2, squared=true)
void env_project_phi(Env* e, int site, int nsites, cplx* phi);                           // phi = conj(project(...)): projmps.jl:153-185
void env_product1_dev(Env* e, const cplx* A, int site, cplx* out);                       // product(..., nsites=1), rank-2 branch
void mps_replacesite1(Mps* m, const cplx* A, int site, bool direction, bool normalize);  // replacesites!, one-site branch: gmps.jl:204-213
EnvSum* envsum_create(Ctx* c, int n, Env* const* projs, int center);
void envsum_free(EnvSum* es);
void envsum_movecenter(EnvSum* es, int idx);
cplx envsum_calculate(EnvSum* es);
void envsum_prepare(EnvSum* es, int site, int nsites);
void envsum_apply(EnvSum* es, const cplx* in, int site, int nsites, cplx* out);          // product(projVs, A, direction, nsites)
void envsum_project_phi(EnvSum* es, int site, int nsites, cplx* out);                    // conj(project(projVs, ...))
void dmrg_halfsweep_sum(Mps* psi, EnvSum* es, bool direction, int nsites, Lanczos lz, Trunc tr, double* energy, long long* maxbond);
void vmps_halfsweep(Mps* psi, EnvSum* es, bool direction, int nsites, Trunc tr, long long* maxbond);

// --- infinite TEBD, two-site cell (tn_itebd.cu) -----------------------------------------------------------------------
struct IMps {                      // iGMPS of rank 1 (igmps.jl:8-17): singulars[i] sits to the left of tensors[i], periodic cell
  Ctx* ctx; int d, L;
  std::vector<Tensor> gam;          // Gamma_i (D_{i-1}, d, D_i)
  std::vector<double*> sing;        // device, length D_{i-1}
  std::vector<long long> nsing;
  std::vector<double> norms;        // accumulated log-norms (itebd.jl:106-108)
};
IMps* imps_create(Ctx* c, int d, int L, const long long* dims, const cplx* const* host_sites, const double* const* host_sing);
void imps_free(IMps* m);
void itebd_apply_gate2(IMps* m, const cplx* gate_dev, Trunc tr);

// tn_qjmc.cu: counter-based uniforms of the throughput runs (Philox4x32-10)
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double counter_uniform(uint64_t seed, uint64_t traj, uint64_t step, uint64_t slot);

}  // namespace tn
