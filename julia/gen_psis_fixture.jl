# gen_psis_fixture.jl — pins SURVEY §8 row a12 (PSIS.psis) for anyone who HAS Julia: the engine's PSIS
# stage (pathfinder_b200/csrc/k6_psis_resample.cu, oracle/psis.py) follows the published algorithm
# because PSIS.jl is not part of the reference tree ("parity unpinned").  This script runs the real
# PSIS.jl on the repo's golden log ratios and writes its smoothed log weights, weights, Pareto k-hat and
# tail length next to them; tests/test_golden.py::test_psis_against_psis_jl_fixture then compares the
# oracle (and, on a GPU box, K6) with it.  Never executed in the build container (no Julia there).
#
#   julia --project=/path/to/Pathfinder.jl julia/gen_psis_fixture.jl tests/golden
#
# Input : <dir>/psis_resample.npz  (written by scripts/make_golden.py; array log_ratios)
# Output: <dir>/psis_jl_fixture.npz (arrays log_ratios, log_weights, weights, pareto_k, tail_length)
using PSIS
using NPZ   # ] add NPZ

dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
g = npzread(joinpath(dir, "psis_resample.npz"))
logr = Vector{Float64}(vec(g["log_ratios"]))
# exactly the call of src/resample.jl:78; the fields read at src/resample.jl:64 and src/multipath.jl:53
r = PSIS.psis(logr)
out = Dict{String,Any}(
    "log_ratios" => logr,
    "log_weights" => collect(r.log_weights),
    "weights" => collect(r.weights),
    "pareto_k" => [Float64(r.pareto_shape)],
    "tail_length" => [Int64(r.tail_length)],
)
npzwrite(joinpath(dir, "psis_jl_fixture.npz"), out)
println("wrote ", joinpath(dir, "psis_jl_fixture.npz"), " (PSIS.jl ", pkgversion(PSIS), ", N = ", length(logr), ")")
