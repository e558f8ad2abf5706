# bench/ref_cpu.jl — times the REAL reference package (Wandao123/IsingModel.jl) on BASELINE config 1 for anyone who
# has Julia (this repository's build image has none, so bench.py's CPU arm times a C restatement instead).
#
#   julia --project=/path/to/IsingModel.jl bench/ref_cpu.jl [sweeps]
#
# Config 1: 32x32 periodic ferromagnet, Metropolis single-spin flips at T = 2.269.  One reference "step" is one
# single-spin attempt at a uniformly random site (src/SamplingHelper.jl:39-49), so `sweeps` sweeps = sweeps * 1024
# steps.  Prints spin-updates/s, the metric of bench.py.
using IsingModel
using Random
using SparseArrays

function lattice(L)
    N = L * L
    J = spzeros(N, N)
    for y in 0:L-1, x in 0:L-1
        i = x + y * L + 1
        for (dx, dy) in ((1, 0), (-1, 0), (0, 1), (0, -1))
            j = mod(x + dx, L) + mod(y + dy, L) * L + 1
            J[i, j] = 1.0
        end
    end
    J
end

function main()
    sweeps = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 100
    L = 32; N = L * L; T = 2.269
    rng = MersenneTwister(1)
    s0 = 2 .* rand(rng, Bool, N) .- 1
    for J in (lattice(L), Matrix(lattice(L)))          # sparse (as the reference's tests build it) and dense
        ss = SpinSystems.SpinSystem(copy(s0), J, zeros(N))
        ua = SingleSpinFlip.MetropolisMethod(ss, T)
        n = sweeps * N
        for _ in SamplingHelper.makeSampler!(ua, N; rng = rng) end      # warm-up / compilation
        t = @elapsed for _ in SamplingHelper.makeSampler!(ua, n; rng = rng) end
        println(typeof(J), ": ", n, " single-spin updates in ", round(t, digits = 3), " s = ", round(n / t, sigdigits = 4),
                " spin-updates/s (1 thread; the reference has no threading)")
    end
end

main()
