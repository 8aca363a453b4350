# PathfinderB200.jl — the Julia side of the drop-in boundary: a thin `ccall` shim over
# libpfb200.so (C ABI: include/pfb200.h).
#
# STATUS: source only.  Julia is not installed in the build container, so this file has never been
# executed; the same ABI is exercised end to end by the Python ctypes twin
# (pathfinder_b200/_lib.py, engine.py, api.py), which the GPU parity tests call.
#
# What stays in Julia (unchanged Pathfinder.jl code): `optimize_with_trace` (src/optimize.jl:35),
# init sampling, retry policy, result structs.  What moves to the GPU, batched over
# (path x iteration):
#     fit_mvnormals          src/singlepath.jl:301-303
#     maximize_elbo          src/singlepath.jl:306-308
#     success / draws        src/singlepath.jl:309-314, 224-233
#     _compute_psis_result   src/multipath.jl:220-224
#     _resample              src/multipath.jl:225
module PathfinderB200

using Random

# a constant path: ccall resolves (symbol, library) once per call site
const LIB = get(ENV, "PFB200_LIB", "libpfb200.so")

const PFB_MODEL_ISONORMAL = Cint(0)
const PFB_MODEL_FUNNEL = Cint(1)
const PFB_MODEL_DIAGNORMAL = Cint(2)

# mirrors `pfb_config` (include/pfb200.h); defaults = src/Pathfinder.jl:24-27
struct PfbConfig
    device::Int32
    history_length::Int32
    ndraws_elbo::Int32
    materialize_all::Int32
    elbo_mode::Int32
    reserved::Int32
    eps::Float64
end

# mirrors `pfb_elbo_out`: 18 pointers, NULL = skip
mutable struct PfbElboOut
    elbo::Ptr{Float64}
    elbo_se::Ptr{Float64}
    logp::Ptr{Float64}
    logq::Ptr{Float64}
    best_iter::Ptr{Int64}
    success::Ptr{Int32}
    n_rejected::Ptr{Int64}
    draws::Ptr{Float64}
    draws_logp::Ptr{Float64}
    draws_logq::Ptr{Float64}
    fit_mu::Ptr{Float64}
    fit_alpha::Ptr{Float64}
    fit_vh::Ptr{Float64}
    fit_T::Ptr{Float64}
    fit_Vc::Ptr{Float64}
    fit_logdet::Ptr{Float64}
    fit_jeff::Ptr{Int32}
    all_draws::Ptr{Float64}
end
PfbElboOut() = PfbElboOut(ntuple(_ -> C_NULL, 18)...)

# mirrors `pfb_resample_out`
mutable struct PfbResampleOut
    log_weights::Ptr{Float64}
    weights::Ptr{Float64}
    pareto_k::Ptr{Float64}
    tail_len::Ptr{Int64}
    inds::Ptr{Int64}
    ids::Ptr{Int64}
    draws::Ptr{Float64}
end
PfbResampleOut() = PfbResampleOut(ntuple(_ -> C_NULL, 7)...)
# isbits twin (same layout): a Vector of these is the contiguous pfb_resample_out[ndev] of the *_all calls
struct PfbResampleOutC
    log_weights::Ptr{Float64}
    weights::Ptr{Float64}
    pareto_k::Ptr{Float64}
    tail_len::Ptr{Int64}
    inds::Ptr{Int64}
    ids::Ptr{Int64}
    draws::Ptr{Float64}
end
PfbResampleOutC(o::PfbResampleOut) = PfbResampleOutC(o.log_weights, o.weights, o.pareto_k, o.tail_len, o.inds, o.ids, o.draws)

mutable struct Engine
    handle::Ptr{Cvoid}
    n::Int
    K::Int
    J::Int
    function Engine(n::Integer, family::Integer, blob::Vector{Float64}=Float64[];
                    history_length::Integer=6, ndraws_elbo::Integer=5, device::Integer=0,
                    materialize_all::Bool=false, eps::Float64=1e-12)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        cfg = Ref(PfbConfig(device, history_length, ndraws_elbo, materialize_all, 0, 0, eps))
        rc = ccall((:pfb_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{PfbConfig}), h, cfg)
        rc == 0 || throw(ErrorException("pfb_create failed ($rc): " * last_error(C_NULL)))
        e = new(h[], n, ndraws_elbo, history_length)
        finalizer(close, e)
        rc = ccall((:pfb_register_model, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Csize_t),
                   e.handle, family, n, isempty(blob) ? C_NULL : pointer(blob), length(blob))
        check(e, rc)
        return e
    end
end

function Base.close(e::Engine)
    if e.handle != C_NULL
        ccall((:pfb_destroy, LIB), Cint, (Ptr{Cvoid},), e.handle)
        e.handle = C_NULL
    end
    return nothing
end

last_error(h) = unsafe_string(ccall((:pfb_last_error, LIB), Cstring, (Ptr{Cvoid},), h))

# error convention of include/pfb200.h: < 0 argument / shape error, > 0 CUDA runtime error
function check(e::Engine, rc::Integer)
    rc == 0 && return nothing
    msg = last_error(e.handle)
    rc == -2 && throw(DimensionMismatch(msg))
    rc < 0 && throw(ArgumentError(msg))
    throw(ErrorException("CUDA error $rc: $msg"))
end

"""
    elbo_batch(engine, traces, seeds; normals=nothing)

`traces`: vector of `(points, gradients)` with `points::Vector{Vector{Float64}}` exactly as
`OptimizationTrace` holds them (src/optimize.jl:110-114); `seeds[p]`: the `UInt64` seeds
`rand!(rng, Vector{UInt64}(undef, L_p))` of src/elbo.jl:2.  Returns a NamedTuple with, per
(path, iteration): `elbo`, `elbo_se`; per path: `fit_iteration`, `success`,
`num_bfgs_updates_rejected`, `draws[n, K, P]`, `logp`, `logq`, and the best-iteration
`WoodburyPDMat` ingredients (`mu, alpha, vh, T, Vc, logdet, jeff`).
"""
function elbo_batch(e::Engine, traces, seeds::Vector{Vector{UInt64}}; normals=nothing, fallback_seeds=nothing)
    P = length(traces)
    n, K = e.n, e.K
    if fallback_seeds !== nothing
        # one UInt64 per path, drawn from the PATH's rng: a failed path returns
        # rand(rng, fit_distributions[fit_iteration + 1], ndraws) (src/singlepath.jl:224-228)
        fs = Vector{UInt64}(fallback_seeds)
        GC.@preserve fs check(e, ccall((:pfb_set_fallback_seeds, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt64}),
                                       e.handle, length(fs), fs))
    end
    offsets = zeros(Int64, P + 1)
    for (p, (pts, _)) in enumerate(traces)
        offsets[p + 1] = offsets[p] + length(pts)
    end
    T = offsets[end]
    U = T - P
    X = Matrix{Float64}(undef, n, T)      # column-major, what the ABI expects
    G = Matrix{Float64}(undef, n, T)
    for (p, (pts, grads)) in enumerate(traces), (l, (x, g)) in enumerate(zip(pts, grads))
        X[:, offsets[p] + l] .= x
        G[:, offsets[p] + l] .= g
    end
    sd = reduce(vcat, seeds; init=UInt64[])
    length(sd) == U || throw(DimensionMismatch("need one seed per (path, iteration)"))
    kp = ccall((:pfb_kp, LIB), Cint, (Ptr{Cvoid},), e.handle)
    elbo = Vector{Float64}(undef, U); se = similar(elbo)
    best = Vector{Int64}(undef, P); succ = Vector{Int32}(undef, P); rej = Vector{Int64}(undef, P)
    draws = Array{Float64}(undef, n, K, P); lp = Matrix{Float64}(undef, K, P); lq = similar(lp)
    mu = Matrix{Float64}(undef, n, P); alpha = similar(mu); vh = Array{Float64}(undef, n, kp, P)
    Tm = Array{Float64}(undef, kp, kp, P); Vc = similar(Tm)  # row-major per path: transpose on use
    logdet = Vector{Float64}(undef, P); jeff = Vector{Int32}(undef, P)
    out = PfbElboOut()
    GC.@preserve offsets X G sd normals elbo se best succ rej draws lp lq mu alpha vh Tm Vc logdet jeff begin
        out.elbo = pointer(elbo); out.elbo_se = pointer(se)
        out.best_iter = pointer(best); out.success = pointer(succ); out.n_rejected = pointer(rej)
        out.draws = pointer(draws); out.draws_logp = pointer(lp); out.draws_logq = pointer(lq)
        out.fit_mu = pointer(mu); out.fit_alpha = pointer(alpha); out.fit_vh = pointer(vh)
        out.fit_T = pointer(Tm); out.fit_Vc = pointer(Vc)
        out.fit_logdet = pointer(logdet); out.fit_jeff = pointer(jeff)
        rc = ccall((:pfb_elbo_batch, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt64},
                    Ptr{Float64}, Ref{PfbElboOut}),
                   e.handle, n, P, offsets, X, G, sd, normals === nothing ? C_NULL : pointer(normals), out)
        check(e, rc)
    end
    return (; offsets, elbo, elbo_se=se, fit_iteration=best, success=succ .!= 0,
            num_bfgs_updates_rejected=rej, draws, logp=lp, logq=lq, mu, alpha, vh,
            T=permutedims(Tm, (2, 1, 3)), Vc=permutedims(Vc, (2, 1, 3)), logdet, jeff)
end

"""
    psis_resample(engine, seed, ndraws; importance=true, replace=true)

`_compute_psis_result` + `_resample` (src/multipath.jl:220-225) on the pool left on the device
by the last `elbo_batch` (N = P * K draws, draw-fastest / component-slowest).
"""
function psis_resample(e::Engine, P::Integer, seed::UInt64, ndraws::Integer; importance::Bool=true,
                       replace::Bool=true)
    N = P * e.K
    lw = Vector{Float64}(undef, N); w = similar(lw)
    k = Ref(NaN); tl = Ref(Int64(0))
    inds = Vector{Int64}(undef, ndraws); ids = similar(inds)
    draws = Matrix{Float64}(undef, e.n, ndraws)
    out = PfbResampleOut()
    GC.@preserve lw w k tl inds ids draws begin
        if importance
            out.log_weights = pointer(lw); out.weights = pointer(w)
        end
        out.pareto_k = Base.unsafe_convert(Ptr{Float64}, k)
        out.tail_len = Base.unsafe_convert(Ptr{Int64}, tl)
        out.inds = pointer(inds); out.ids = pointer(ids); out.draws = pointer(draws)
        rc = ccall((:pfb_psis_resample, LIB), Cint,
                   (Ptr{Cvoid}, UInt64, Cint, Cint, Cint, Ref{PfbResampleOut}), e.handle, seed, ndraws, importance,
                   replace, out)
        check(e, rc)
    end
    return (; log_weights=lw, weights=w, pareto_shape=k[], tail_length=tl[], sample_inds=inds,
            draw_component_ids=ids, draws)
end

"""
    unit_fits(engine, units) -> NamedTuple

`fit_distributions[l + 1]` of arbitrary (path, iteration) units of the current batch
(src/singlepath.jl:64 keeps every iteration's; the engine exports them on demand) in the
WoodburyPDMat ingredients of `elbo_batch`.  `units` are 0-based, path-major / iteration-minor.
"""
function unit_fits(e::Engine, units::Vector{Int32})
    m = length(units); n = e.n
    kp = ccall((:pfb_kp, LIB), Cint, (Ptr{Cvoid},), e.handle)
    mu = Matrix{Float64}(undef, n, m); alpha = similar(mu); vh = Array{Float64}(undef, n, kp, m)
    Tm = Array{Float64}(undef, kp, kp, m); Vc = similar(Tm)
    logdet = Vector{Float64}(undef, m); jeff = Vector{Int32}(undef, m)
    GC.@preserve units mu alpha vh Tm Vc logdet jeff check(e, ccall((:pfb_unit_fits, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Int32}), e.handle, m, units, mu, alpha, vh, Tm, Vc, logdet, jeff))
    return (; mu, alpha, vh, T=permutedims(Tm, (2, 1, 3)), Vc=permutedims(Vc, (2, 1, 3)), logdet, jeff)
end

# ---- multi-GPU: the pool exchange behind the ABI (NCCL is dlopen'ed by the library) -----------------------

"128-byte communicator id: create it on one rank, ship it to the others (MPI.Bcast!, a file, ...)."
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    rc = ccall((:pfb_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("pfb_comm_unique_id failed with code $rc (is libnccl.so.2 loadable?)")
    return id
end

"One process per GPU: every rank calls this with the same id."
comm_init!(e::Engine, id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(e, ccall((:pfb_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), e.handle, id, rank, world))

"One process, all GPUs: engine i (created with device = i - 1) becomes rank i - 1."
function comm_init_all!(engines::Vector{Engine})
    hs = [e.handle for e in engines]
    rc = GC.@preserve hs ccall((:pfb_comm_init_all, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), hs, length(hs))
    check(engines[1], rc)
end

function _resample_out(n, N, ndraws, importance)
    lw = Vector{Float64}(undef, N); w = similar(lw)
    k = Ref(NaN); tl = Ref(Int64(0))
    inds = Vector{Int64}(undef, ndraws); ids = similar(inds)
    draws = Matrix{Float64}(undef, n, ndraws)
    out = PfbResampleOut()
    if importance
        out.log_weights = pointer(lw); out.weights = pointer(w)
    end
    out.pareto_k = Base.unsafe_convert(Ptr{Float64}, k)
    out.tail_len = Base.unsafe_convert(Ptr{Int64}, tl)
    out.inds = pointer(inds); out.ids = pointer(ids); out.draws = pointer(draws)
    return out, (lw, w, k, tl, inds, ids, draws)
end

"""
    pool_exchange_resample(engine, paths_per_rank, seed, ndraws; importance=true, replace=true)

`_compute_psis_result` + `_resample` (src/multipath.jl:220-225) over the runs of ALL ranks: all-gather of
the pools' per-draw log densities, PSIS and the index draw replicated, owned columns regenerated on their
rank, sum-reduce.  Every rank calls it with the same arguments and receives the same result.
"""
function pool_exchange_resample(e::Engine, paths_per_rank::Vector{Int32}, seed::UInt64, ndraws::Integer;
                                importance::Bool=true, replace::Bool=true)
    N = Int(sum(paths_per_rank)) * e.K
    out, keep = _resample_out(e.n, N, ndraws, importance)
    GC.@preserve keep paths_per_rank check(e, ccall((:pfb_pool_exchange_resample, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, UInt64, Cint, Cint, Cint, Ref{PfbResampleOut}), e.handle, paths_per_rank, seed,
        ndraws, importance, replace, out))
    lw, w, k, tl, inds, ids, draws = keep
    return (; log_weights=lw, weights=w, pareto_shape=k[], tail_length=tl[], sample_inds=inds,
            draw_component_ids=ids, draws)
end

"""
    multipathfinder_b200_multi(optimize_one, family, dim, ndraws; nruns, devices, ...)

`multipathfinder` on several GPUs of ONE process, no MPI: the runs shard over `devices` in contiguous
blocks (the pool keeps the run order of src/multipath.jl:217), every shard's ELBO stage runs on its own
task, and one `pfb_pool_exchange_resample_all` call does the PSIS pool exchange for all of them.
"""
function multipathfinder_b200_multi(optimize_one, family::Integer, dim::Integer, ndraws::Integer;
                                    nruns::Integer, devices::Vector{Int}, ndraws_elbo::Integer=5,
                                    history_length::Integer=6, init_scale::Real=2,
                                    rng::AbstractRNG=Random.default_rng(), importance::Bool=true,
                                    blob::Vector{Float64}=Float64[])
    world = length(devices)
    run_seeds = rand(rng, UInt64, nruns)                       # src/multipath.jl:162
    engines = [Engine(dim, family, blob; history_length, ndraws_elbo, device=d) for d in devices]
    try
        comm_init_all!(engines)
        bounds = [div(nruns * r, world) for r in 0:world]      # contiguous, balanced blocks of runs
        fits = Vector{Any}(undef, world)
        @sync for r in 1:world
            Threads.@spawn begin
                runs = (bounds[r] + 1):bounds[r + 1]
                traces = Vector{Any}(undef, length(runs)); seeds = Vector{Vector{UInt64}}(undef, length(runs))
                fb = Vector{UInt64}(undef, length(runs))
                for (j, p) in enumerate(runs)
                    prng = copy(rng); Random.seed!(prng, run_seeds[p])
                    x0 = (rand(prng, dim) .* 2 .- 1) .* init_scale
                    traces[j] = optimize_one(x0)
                    seeds[j] = rand(prng, UInt64, length(traces[j][1]) - 1)
                    fb[j] = rand(prng, UInt64)
                end
                fits[r] = elbo_batch(engines[r], traces, seeds; fallback_seeds=fb)
            end
        end
        ppr = Int32[bounds[r + 1] - bounds[r] for r in 1:world]
        N = nruns * ndraws_elbo
        outs = Vector{PfbResampleOutC}(undef, world); keeps = Vector{Any}(undef, world)
        for r in 1:world
            o, keeps[r] = _resample_out(dim, N, ndraws, importance)
            outs[r] = PfbResampleOutC(o)
        end
        hs = [e.handle for e in engines]
        GC.@preserve hs keeps ppr outs check(engines[1], ccall((:pfb_pool_exchange_resample_all, LIB), Cint,
            (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int32}, UInt64, Cint, Cint, Cint, Ptr{PfbResampleOutC}), hs, world, ppr,
            rand(rng, UInt64), ndraws, importance, true, outs))
        lw, w, k, tl, inds, ids, draws = keeps[1]              # identical on every rank
        return (; fits, log_weights=lw, weights=w, pareto_shape=k[], tail_length=tl[], sample_inds=inds,
                draw_component_ids=ids, draws)
    finally
        foreach(close, engines)
    end
end

# mirrors `pfb_lbfgs_opts`
struct PfbLbfgsOpts
    maxiters::Int32
    max_points::Int32
    gtol::Float64
    ftol::Float64
end

# the C callback of row f2 (pfb_logp_callback): one tile of draws -> one log density per column
function _logp_tile(user::Ptr{Cvoid}, x::Ptr{Float64}, n::Int64, m::Int64, out::Ptr{Float64})::Cvoid
    logp = unsafe_pointer_to_objref(user).x
    X = unsafe_wrap(Array, x, (n, m)); o = unsafe_wrap(Array, out, m)
    @inbounds for k in 1:m
        o[k] = try logp(view(X, :, k)) catch; NaN end
    end
    return nothing
end
"""
    register_host_model!(engine, logp)

Row f2: make an arbitrary Julia closure `logp(x::AbstractVector)` the target density
(src/singlepath.jl:186; `logp.(eachcol(ϕ))`, src/elbo.jl:15).  The engine calls it on pinned tiles
of draws while the next tile is being sampled.  Keep `logp` alive as long as the engine.
"""
function register_host_model!(e::Engine, logp)
    box = Ref{Any}(logp)
    cb = @cfunction(_logp_tile, Cvoid, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}))
    rc = ccall((:pfb_register_host_model, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
               e.handle, e.n, cb, pointer_from_objref(box))
    check(e, rc)
    return box  # the caller GC.@preserve's this around every engine call
end

"""
    lbfgs_batch(engine, inits; maxiters=1000, gtol=1e-8, ftol=1e-14) -> (npoints, status)

Row f1: the L-BFGS trajectories of all paths in one kernel launch (closed-form families); the
traces stay on the device — follow with `batch_from_lbfgs(engine, seeds)` and the usual run /
download calls, and `lbfgs_download` for the `OptimizationTrace` fields (src/optimize.jl:110-114).
"""
function lbfgs_batch(e::Engine, inits::Matrix{Float64}; maxiters::Integer=1000, gtol=1e-8, ftol=1e-14)
    P = size(inits, 2)
    npts = Vector{Int64}(undef, P); status = Vector{Int32}(undef, P)
    opts = Ref(PfbLbfgsOpts(maxiters, maxiters + 1, gtol, ftol))
    rc = ccall((:pfb_lbfgs_batch, LIB), Cint,
               (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ref{PfbLbfgsOpts}, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}),
               e.handle, e.n, P, inits, opts, npts, status, C_NULL)
    check(e, rc)
    return npts, status
end
function batch_from_lbfgs(e::Engine, seeds::Vector{UInt64})
    check(e, ccall((:pfb_batch_from_lbfgs, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}), e.handle, seeds))
end
function lbfgs_download(e::Engine, npts::Vector{Int64})
    T = sum(npts)
    X = Matrix{Float64}(undef, e.n, T); G = similar(X); fx = Vector{Float64}(undef, T)
    check(e, ccall((:pfb_lbfgs_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   e.handle, X, G, fx))
    return X, G, fx
end

"""
    unit_draws(engine, units) -> (draws, logp, logq)

`ELBOEstimate.draws / .log_densities_target / .log_densities_fit` (src/elbo.jl:22-29) of arbitrary
iterations of the current batch, regenerated on the device from their seeds (0-based unit indices:
`offsets[p] - (p - 1) + l - 1` for iteration `l` of path `p`, with 1-based `p`, 0-based `offsets`).
"""
function unit_draws(e::Engine, units::Vector{Int32})
    m = length(units)
    draws = Array{Float64}(undef, e.n, e.K, m); lp = Matrix{Float64}(undef, e.K, m); lq = similar(lp)
    check(e, ccall((:pfb_unit_draws, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   e.handle, m, units, draws, lp, lq))
    return draws, lp, lq
end

"""
    pool_draws(engine, P) -> (draws, logp, logq)

`PathfinderResult.draws` of every run (src/singlepath.jl:231-232), fetched when needed: the engine
materialises the pool's draws only on request — `psis_resample` regenerates just the resampled columns.
"""
function pool_draws(e::Engine, P::Integer)
    draws = Array{Float64}(undef, e.n, e.K, P); lp = Matrix{Float64}(undef, e.K, P); lq = similar(lp)
    check(e, ccall((:pfb_pool_download, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   e.handle, 0, P, draws, lp, lq))
    return draws, lp, lq
end

"""
    multipathfinder_b200(optimize_one, model_family, dim, ndraws; nruns, ndraws_elbo, rng, ...)

Drop-in for the ELBO-and-resample part of `multipathfinder` (src/multipath.jl:118-245).
`optimize_one(init) -> (points, gradients)` is the caller's unchanged trajectory producer, e.g.

    optimize_one(x0) = let (_, tr) = Pathfinder.optimize_with_trace(SciMLBase.remake(prob; u0=x0), optimizer)
        (tr.points, tr.gradients)
    end

Seeds are drawn exactly where the reference draws them (`run_seeds`, src/multipath.jl:162; per
iteration, src/elbo.jl:2), so the run is reproducible under `Random.seed!` like the reference's.
"""
function multipathfinder_b200(optimize_one, family::Integer, dim::Integer, ndraws::Integer;
                              nruns::Integer, ndraws_elbo::Integer=5, history_length::Integer=6,
                              init_scale::Real=2, rng::AbstractRNG=Random.default_rng(),
                              importance::Bool=true, blob::Vector{Float64}=Float64[], device::Integer=0)
    run_seeds = rand(rng, UInt64, nruns)                       # src/multipath.jl:162
    eng = Engine(dim, family, blob; history_length, ndraws_elbo, device)
    try
        traces = Vector{Any}(undef, nruns)
        seeds = Vector{Vector{UInt64}}(undef, nruns)
        fb = Vector{UInt64}(undef, nruns)
        for p in 1:nruns                                       # host phase A; use @threads / ntasks freely
            prng = copy(rng); Random.seed!(prng, run_seeds[p]) # src/multipath.jl:191-193
            x0 = (rand(prng, dim) .* 2 .- 1) .* init_scale     # UniformSampler, src/singlepath.jl:340-344
            traces[p] = optimize_one(x0)
            seeds[p] = rand(prng, UInt64, length(traces[p][1]) - 1)   # src/elbo.jl:2
            fb[p] = rand(prng, UInt64)                         # the path's rng, used only if the path fails
        end
        fit = elbo_batch(eng, traces, seeds; fallback_seeds=fb)
        res = psis_resample(eng, nruns, rand(rng, UInt64), ndraws; importance)
        return (; fit, res...)
    finally
        close(eng)
    end
end

end # module
