# baseline/julia_ref.jl — times the REAL reference package on the bench's configuration, for anyone with Julia ≥ 1.9
# (Project.toml:27 of OpenQuantumBase.jl).  Julia is not installed in the build container or on the GPU box, so this
# script has never been executed here; bench.py's `--impl reference` arm times the reference-structured CPU restatement
# (oracle/cpu_baseline.py) instead and says so (kind: "port").
#
#   julia --project=/path/to/OpenQuantumBase.jl -t auto baseline/julia_ref.jl [nqubits=10] [nrho=2]
#
# Workload (bench.py build_c3): H(s) = −(1−s)ΣX_i + s·Σ_{i<j} J_ij Z_iZ_j (unit :h), couplings Z_1…Z_q, Ohmic(1e-4, 4, 16),
# S ≡ 0, s = 0.5, lvl = 2^q.  One "eval" = one (du, u, p, t) call of DiffEqLiouvillian{true,false} on one ρ — the external
# ensemble solver calls it once per member.  Prints one JSON line in bench.py's format.
using OpenQuantumBase, LinearAlgebra, Random, Printf

q = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 10
nrho = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2
n = 2^q
rng = MersenneTwister(3000)   # NOTE: numpy's default_rng(3000) draws different J_ij; timing does not depend on them

Hd = -standard_driver(q)
idx = [[i, j] for i in 1:q for j in i+1:q]
Hp = two_local_term([2rand(rng) - 1 for _ in idx], idx, q)
H = DenseHamiltonian([(s) -> 1 - s, (s) -> s], [Matrix(Hd), Matrix(Hp)], unit=:h)
coupling = ConstantCouplings([join([k == i ? "Z" : "I" for k in 1:q]) for i in 1:q], unit=:h)
bath = Ohmic(1e-4, 4, 16)
davies = DaviesGenerator(coupling, (w) -> spectrum(w, bath), (w) -> 0.0)
Op = DiffEqLiouvillian(H, [davies], [], n)
p = ODEParams(H, 10.0, (tf, t) -> t / tf)

g = randn(rng, ComplexF64, n, 16)
ρ = g * g'
ρ ./= tr(ρ)
du = zero(ρ)

# NOTE: at q = 10 the literal per-gap-group loop of davies.jl:21-47 allocates O(#groups · l²) temporaries and the GapIndices
# builder (opensys_util.jl:25-61) inserts into sorted vectors — expect this call not to finish in reasonable time; run with
# q ≤ 6 to obtain numbers and scale with the exponent bench.py reports in cpu_baseline.literal_davies_loop.
Op(du, ρ, p, 5.0)   # warm-up / compilation
t = @elapsed for _ in 1:nrho
    Op(du, ρ, p, 5.0)
end
@printf("{\"impl\": \"reference-julia\", \"metric\": \"rho_rhs_evals_per_sec\", \"value\": %.6g, \"unit\": \"evals/s\", \"n\": %d, \"threads\": %d, \"blas_threads\": %d}\n",
        nrho / t, n, Threads.nthreads(), BLAS.get_num_threads())
