# Compare the committed exact-arithmetic values (tests/golden/exact_mkA_<n>_W.csv, produced by
# tests/golden/make_exact_krylov.py with 60-digit mpmath) against the REAL package.  Not run in CI here (the build
# container has no Julia); a maintainer runs it once from the repository root:
#
#     julia --project=/path/to/ExponentialUtilities.jl tests/golden/check_with_julia.jl
#
# The calls are the reference's own seed-free fixture, test/basictests.jl:859-882: mkA(n), b = [1/i], m = 30,
# expv!(w, 0.1, Ks) and phiv!(w, 0.1, Ks, 3).  Expected output: relative differences at the 1e-14 level.
using ExponentialUtilities, LinearAlgebra, DelimitedFiles

mkA(n) = [i == j ? -2.0 : 0.1 / (1 + abs(i - j)) * (i < j ? 1.0 : 0.5) for i in 1:n, j in 1:n]

for n in (64, 200)
    A = mkA(n)
    b = [1.0 / i for i in 1:n]
    Ks = arnoldi(A, b; m = 30)
    w = zeros(n)
    expv!(w, 0.1, Ks)
    W = Matrix{Float64}(undef, n, 4)
    phiv!(W, 0.1, Ks, 3)
    G = readdlm(joinpath(@__DIR__, "exact_mkA_$(n)_W.csv"), ',', Float64)
    println("n = $n: expv rel. diff ", norm(w - G[:, 1]) / norm(G[:, 1]),
            "   phiv rel. diff ", norm(W - G) / norm(G))
    @assert norm(w - G[:, 1]) / norm(G[:, 1]) < 1e-12
    @assert norm(W - G) / norm(G) < 1e-12
end
println("reference outputs agree with the committed exact values")
