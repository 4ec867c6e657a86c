# gen_golden.jl — reference-side golden vectors for the linearise-and-discretise path.
#
# Run on a machine that has Julia and the packages the reference itself needs (there is no Project.toml in the
# reference: DifferentialEquations, DiffEqSensitivity, ForwardDiff, Zygote, StaticArrays, Interpolations, CSV,
# DataFrames, SymEngine, CommonSubexpressions, MacroTools, MathOptInterface, DiffResults):
#
#     julia tools/gen_golden.jl /path/to/SuccessiveConvexification [/path/to/this/repo]
#
# It evaluates the REFERENCE'S OWN code — `Dynamics.rk4` (dynamics.jl:112-134) and its exact forward-mode Jacobian
# (`Dynamics.sensitivity_zygote`, dynamics.jl:311-313; ForwardDiff.jacobian of the same closure if Zygote is not
# installed) — on the committed inputs tests/golden/julia_inputs.f64 and writes
#     tests/golden/julia_golden.f64   per case, trajectory b and interval i: endpoint (14) then D = d endpoint / d inp
#                                      (14 x 21, column-major), Float64 little-endian
#     tests/golden/julia_golden.txt   manifest: Julia / package versions, one line per case `name B n_nodes offset`
# tests/test_julia_golden.py consumes the two files when present (oracle and CUDA path against them, 1e-10); until then
# parity stays "unpinned by the reference".  Nothing of this repository is loaded: the numbers are the reference's.
#
# The reference does not `include` cleanly as a whole (rocketland.jl:54-57 holds merge-conflict markers and needs a
# Mosek licence), so only the files on the path are loaded, in master.jl's order: the RocketlandDefns module (the text
# of master.jl up to its first include), aerodynamics.jl, symbolic_diff.jl, dynamics.jl, sample_problems.jl.
#
# UNVERIFIED: written without a Julia toolchain (the build container has none).

length(ARGS) >= 1 || error("usage: julia tools/gen_golden.jl /path/to/SuccessiveConvexification [/path/to/repo]")
const REFDIR = abspath(ARGS[1])
const REPO = length(ARGS) >= 2 ? abspath(ARGS[2]) : dirname(@__DIR__)
const GOLDEN = joinpath(REPO, "tests", "golden")

cd(REFDIR)                                       # sample_problems.jl:25 loads "aero/lift_drag.csv" relative to the cwd
let src = read(joinpath(REFDIR, "master.jl"), String)
    cut = findfirst("include(\"aerodynamics.jl\")", src)
    cut === nothing && error("master.jl: include(\"aerodynamics.jl\") not found")
    include_string(Main, src[1:first(cut)-1], "master.jl")          # module RocketlandDefns (master.jl:1-136)
end
include(joinpath(REFDIR, "aerodynamics.jl"))
include(joinpath(REFDIR, "symbolic_diff.jl"))
include(joinpath(REFDIR, "dynamics.jl"))
include(joinpath(REFDIR, "sample_problems.jl"))

import ForwardDiff

const prob = SampleProblems.base_prob_aero_scaled
const info = RocketlandDefns.ProbInfo(prob)
const cache = RocketlandDefns.IntegratorCache(nothing, nothing, nothing, nothing, Any[1.0, info], info)

# (y, D) of one interval by the reference's own functions
function reference_interval(inp::Vector{Float64}, dt::Float64)
    y = Vector{Float64}(Dynamics.rk4(inp, dt, info))
    D = try
        _, JT = Dynamics.sensitivity_zygote(inp, dt, cache)          # (y, J') with J' 21 x 14 (dynamics.jl:311-313)
        Matrix{Float64}(permutedims(JT))
    catch err
        @warn "Dynamics.sensitivity_zygote failed; using ForwardDiff.jacobian of the same closure" err
        Matrix{Float64}(ForwardDiff.jacobian(v -> Dynamics.rk4(v, dt, info), inp))
    end
    size(D) == (14, 21) || error("unexpected Jacobian size $(size(D))")
    return y, D
end

# ---- inputs
raw = read(joinpath(GOLDEN, "julia_inputs.f64"))
data = collect(reinterpret(Float64, raw))
cases = []
for line in eachline(joinpath(GOLDEN, "julia_inputs.txt"))
    (isempty(strip(line)) || startswith(line, "#")) && continue
    f = split(line)
    push!(cases, (name = String(f[1]), B = parse(Int, f[2]), n = parse(Int, f[3]), dt = parse(Float64, f[4]),
                  oX = parse(Int, f[5]), oU = parse(Int, f[6]), oS = parse(Int, f[7])))
end

out = Float64[]
manifest = String[]
push!(manifest, "# produced by tools/gen_golden.jl with Julia $(VERSION)")
try
    import Pkg
    for (_, p) in Pkg.dependencies()
        p.name in ("ForwardDiff", "Zygote", "Interpolations", "StaticArrays", "DifferentialEquations") &&
            push!(manifest, "# $(p.name) $(p.version)")
    end
catch
end
for c in cases
    X = reshape(data[c.oX+1 : c.oX+14*c.n*c.B], 14, c.n, c.B)
    U = reshape(data[c.oU+1 : c.oU+3*c.n*c.B], 3, c.n, c.B)
    S = data[c.oS+1 : c.oS+c.B]
    push!(manifest, "$(c.name) $(c.B) $(c.n) $(length(out))")
    for b = 1:c.B, i = 1:c.n-1
        inp = vcat(X[:, i, b], U[:, i, b], U[:, i+1, b], S[b])       # make_state, dynamics.jl:318-320
        y, D = reference_interval(inp, c.dt)
        append!(out, y)
        append!(out, vec(D))
    end
    println("case $(c.name): $(c.B) x $(c.n - 1) intervals done")
end
write(joinpath(GOLDEN, "julia_golden.f64"), reinterpret(UInt8, out))
open(joinpath(GOLDEN, "julia_golden.txt"), "w") do io
    foreach(l -> println(io, l), manifest)
end
println("wrote ", joinpath(GOLDEN, "julia_golden.f64"), " (", length(out), " doubles)")
