# Times the REAL reference on the bench.py workload (bounded sample of the C5-shaped H_eff matvec) on the host's cores:
#   julia -t1 --project=/path/to/TensorNetworks.jl baseline/run_reference.jl CHI W ROWS STEPS WARMUP M1.bin M2.bin
# bench.py --impl reference calls this when `julia` is on the PATH and TN_REFERENCE_JL points at a checkout of the reference
# (cpu_baseline.kind = "julia"); otherwise it times the NumPy port (kind = "port").  UNVERIFIED: no Julia in the build image.
#
# The sample is the one bench.py documents: rows a in [0, ROWS) of one product() call (projmps.jl:103-145) -- the left block is
# restricted to those bra-bond rows, every contraction of the reference's order is linear in that slice.  Inputs: unit-normal
# re/im with Julia's own RNG (the timing does not depend on the values); M1, M2 (w,d,d,w) are read from the two raw column-major
# files bench.py writes (the J1-J2 cylinder MPO tensors of the workload).
using TensorNetworks
using LinearAlgebra
using Random

chi, w, rows, steps, warmup = parse.(Int, ARGS[1:5])
d = 2
BLAS.set_num_threads(Sys.CPU_THREADS)
M1 = Array{ComplexF64}(undef, w, d, d, w); read!(ARGS[6], M1)
M2 = Array{ComplexF64}(undef, w, d, d, w); read!(ARGS[7], M2)
Random.seed!(0)
L = randn(ComplexF64, rows, w, chi) .* sqrt(2)
R = randn(ComplexF64, chi, w, chi) .* sqrt(2)
theta = randn(ComplexF64, chi, d, d, chi) .* sqrt(2)
# a 4-site frame whose blocks 1 and 4 are L and R; product() only reads the blocks, the MPO tensors of sites 2, 3 and theta
psi = GMPS(1, d, Array{ComplexF64}[zeros(ComplexF64, 1, d, 1) for _ in 1:4], 0)
H = GMPS(2, d, Array{ComplexF64}[M1[1:1, :, :, :], M1, M2, M2[:, :, :, 1:1]], 0)
P = ProjMPS(psi, H, psi; rank=2, center=1)
P.blocks[1] = L; P.blocks[4] = R; P.center = 2
for _ in 1:warmup
    product(P, theta, false, 2)
end
t = @elapsed for _ in 1:steps
    product(P, theta, false, 2)
end
flops = 8.0 * (2.0 * chi^3 * d^2 * w + 2.0 * chi^2 * d^3 * w^2) * (rows / chi)
println("{\"seconds_per_step\": $(t / steps), \"tflops\": $(flops * steps / t / 1e12), \"threads\": $(BLAS.get_num_threads()), \"julia\": \"$(VERSION)\"}")
