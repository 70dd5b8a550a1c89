# record_tape.jl -- records a run of the REFERENCE (DifferentialEvolutionMCMC.jl, serial `sample` path) as the structured
# replay tape of include/demcmc_b200.h (`demcmc_tape`, SURVEY.md Appendix A), so that `demcmc_replay` can be fed the
# reference's own random draws: identical accept decisions, proposals and log densities within 1e-12 (north_star (a)).
#
# STATUS: NOT EXECUTED.  There is no Julia toolchain in the build image or on the GPU box (SURVEY.md 0.2).  Until this
# file has run somewhere, chain-level parity with Julia stays UNPINNED and the parity tests replay the tape of the C
# restatement (oracle/demcmc_oracle.c) instead.  The format written here is read by tests/tape_io.py, whose round trip
# is tested with an oracle tape.
#
# How it records WITHOUT re-implementing the sampler.  The reference draws everything from the task-local RNG through
# bare `rand()` calls, so its draws cannot be intercepted; but they can be PREDICTED: before each operator call the
# recorder copies the task-local RNG (`copy(Random.default_rng())` is a Xoshiro with the same state) and draws, on the
# copy and through the very same library calls in the very same order (crossover.jl:154-172, 239-257, 301-321;
# mutation.jl:13-25; migration.jl:31-35, 64-70; utilities.jl:55-58, 291-306), the values the operator is about to draw.
# Then the reference's OWN operator runs on the real RNG, and the recorder checks that the copy and the real RNG ended
# in the same state: a draw-count or draw-order mistake in the prediction is caught on the spot (`@assert` below), so a
# tape that was written is a tape the reference really followed.  The recorder owns only the three dispatch loops above
# the operators (iterations / groups / blocks: main.jl:33-38, 84-89, 161-179, 199-207); the per-particle work is the
# reference's.  The reference's hooks (src/structs.jl:71-74) are used for what they can see: `evaluate_fitness!` and
# `update_particle!` wrappers capture every proposal, its weight, log_adj and the accept decision (the trace the parity
# tests compare), and the `sample` wrapper cross-checks the predicted donors against the particles actually returned.
#
# Usage (where Julia exists):
#     julia --project=/path/to/DifferentialEvolutionMCMC.jl julia/record_tape.jl out_dir [n_iter]
# records Examples/Gaussian_Example.jl's model by default; `record_run(model, de, n_iter, dir)` records any model
# whose parameters are scalars or arrays of Float64.
using DifferentialEvolutionMCMC, Random, Distributions, StatsBase
const DEM = DifferentialEvolutionMCMC

n_elems(θ) = θ isa AbstractArray ? length(θ) : 1
flat(Θ) = Float64[x for θ in Θ for x in (θ isa AbstractArray ? vec(θ) : (θ,))]

mutable struct Tape
    G::Int; Np::Int; d::Int; B::Int; n_iter::Int
    mig_u::Vector{Float64}; mig_n::Vector{Int32}; mig_groups::Matrix{Int32}; mig_pick_u::Matrix{Float64}   # [G, n_iter] (C order [n_iter][G])
    kind::Matrix{UInt8}                      # [P, S]
    idx::Array{Int32, 3}                     # [3, P, S]   0-based slots in the group (resample: particle ids)
    idx_row::Array{Int32, 3}                 # [3, P, S]   resample: 0-based rows of de.samples
    gamma1::Matrix{Float64}; gamma2::Matrix{Float64}; u_acc::Matrix{Float64}       # [P, S]
    noise::Array{Float64, 3}; keep::Array{UInt8, 3}                                 # [d, P, S]
    # trace (what the parity tests compare with demcmc_get_trace): proposals, their weights, log_adj, accept flags
    prop_theta::Array{Float64, 3}; prop_weight::Matrix{Float64}; log_adj::Matrix{Float64}; accepted::Matrix{UInt8}
    state_theta::Array{Float64, 3}; state_id::Matrix{Int32}     # [d, P, n_iter], [P, n_iter]: the groups after every iteration (teacher forcing)
end

function Tape(G, Np, d, B, n_iter)
    P, S = G * Np, n_iter * B
    return Tape(G, Np, d, B, n_iter, zeros(n_iter), zeros(Int32, n_iter), fill(Int32(-1), G, n_iter), zeros(G, n_iter),
        zeros(UInt8, P, S), fill(Int32(-1), 3, P, S), fill(Int32(-1), 3, P, S), zeros(P, S), zeros(P, S), zeros(P, S),
        zeros(d, P, S), zeros(UInt8, d, P, S), zeros(d, P, S), zeros(P, S), zeros(P, S), zeros(UInt8, P, S),
        zeros(d, P, n_iter), zeros(Int32, P, n_iter))
end

# where the recorder is: sweep s (1-based), position p (1-based, group-major), set by the loops below
const CUR = Ref((s = 0, p = 0))
const TAPE = Ref{Tape}()
const EXPECT_DONORS = Ref{Vector{Any}}(Any[])

same_state(a, b) = a == b                                   # Xoshiro defines ==

# Particle + Distribution (utilities.jl:291-306): scalar elements draw rand(d), array elements rand(d, size)
function predict_noise!(rng, dist, Θ, out)
    k = 0
    for θ in Θ
        if θ isa AbstractArray
            v = rand(rng, dist, size(θ))
            out[(k + 1):(k + length(θ))] .= vec(v); k += length(θ)
        else
            out[k += 1] = rand(rng, dist)
        end
    end
end

# recombination! (crossover.jl:301-321): one rand() per element when κ != 1, none otherwise
function predict_keep!(rng, de, Θ, out)
    de.κ == 1.0 && return
    k = 0
    for θ in Θ, _ = 1:n_elems(θ)
        out[k += 1] = rand(rng) <= (1 - de.κ) ? 1 : 0
    end
end

slot_of(group, p) = findfirst(q -> q === p, group)

# donors: `sample(group_diff, n; replace = false)` on the current group (crossover.jl:138-140), or `resample`'s cells of
# de.samples[1:de.iter-1, 1, :] (crossover.jl:113-124)
function predict_donors!(rng, de, group, pool, n, tape, s, p, first_col)
    if de.sample === DEM.resample
        cells = StatsBase.sample(rng, CartesianIndices(@view de.samples[1:(de.iter - 1), 1, :]), n; replace = false)
        for (q, c) in enumerate(cells)
            tape.idx_row[first_col + q - 1, p, s] = c[1] - 1
            tape.idx[first_col + q - 1, p, s] = c[2] - 1
        end
        EXPECT_DONORS[] = Any[]
    else
        picked = StatsBase.sample(rng, pool, n; replace = false)
        for (q, pt) in enumerate(picked)
            tape.idx[first_col + q - 1, p, s] = slot_of(group, pt) - 1
        end
        EXPECT_DONORS[] = Any[picked...]
    end
end

# one crossover!(model, de, group, pt[, block]) call (crossover.jl:30-47, 80-99)
function record_crossover!(model, de, group, pt, block, tape, s, p)
    rng = copy(Random.default_rng())
    d = tape.d
    u_snooker = rand(rng)                                                  # crossover.jl:31 -- drawn even when θsnooker == 0
    if u_snooker <= de.θsnooker                                            # snooker_update! (crossover.jl:239-257)
        tape.kind[p, s] = 1
        predict_donors!(rng, de, group, group, 3, tape, s, p, 1)          # (z, m, n) from the WHOLE group, target included
        tape.gamma1[p, s] = rand(rng, Uniform(1.2, 2.2))
        predict_noise!(rng, Uniform(-de.ϵ, de.ϵ), pt.Θ, @view tape.noise[:, p, s])
        predict_keep!(rng, de, pt.Θ, @view tape.keep[:, p, s])
    else
        tape.kind[p, s] = 0
        gp = de.generate_proposal
        if gp === DEM.random_gamma                                         # crossover.jl:154-172
            w = map(x -> x.weight, group)                                  # select_base (crossover.jl:282-289)
            θ = exp.(w) / sum(exp.(w))
            θ = any(isnan, θ) ? w : θ
            pb = StatsBase.sample(rng, group, Weights(θ))
            tape.idx[1, p, s] = slot_of(group, pb) - 1
        end
        predict_donors!(rng, de, group, setdiff(group, [pt]), 2, tape, s, p, 2)
        if gp === DEM.random_gamma
            tape.gamma1[p, s] = rand(rng, Uniform(0.5, 1))
            tape.gamma2[p, s] = de.iter > de.burnin ? 0.0 : rand(rng, Uniform(0.5, 1))
        elseif gp === DEM.fixed_gamma
            tape.gamma1[p, s] = 2.38
        elseif gp === DEM.variable_gamma
            tape.gamma1[p, s] = 2.38 / sqrt(2 * sum(length.(pt.Θ)))
        else
            error("record_tape: generate_proposal must be random_gamma, fixed_gamma or variable_gamma")
        end
        predict_noise!(rng, Uniform(-de.ϵ, de.ϵ), pt.Θ, @view tape.noise[:, p, s])
        predict_keep!(rng, de, pt.Θ, @view tape.keep[:, p, s])
    end
    tape.u_acc[p, s] = rand(rng)                                           # accept (utilities.jl:55-58): always one rand()
    CUR[] = (s = s, p = p)
    block === nothing ? DEM.crossover!(model, de, group, pt) : DEM.crossover!(model, de, group, pt, block)
    @assert same_state(rng, copy(Random.default_rng())) "record_tape: predicted draws of sweep $s particle $p are out of step with the reference"
    # cross-check with what the reference actually used: u <= p must reproduce its accept decision
    @assert d == length(flat(pt.Θ))
end

# mutation!(model, de, group) (mutation.jl:13-25): per particle d normals, then accept's uniform
function record_mutation!(model, de, group, g, tape, s)
    rng = copy(Random.default_rng())
    for (j, pt) in enumerate(group)
        p = (g - 1) * tape.Np + j
        tape.kind[p, s] = 2
        predict_noise!(rng, Normal(0.0, de.σ), pt.Θ, @view tape.noise[:, p, s])
        tape.u_acc[p, s] = rand(rng)
    end
    CUR[] = (s = s, p = (g - 1) * tape.Np)                                # the hooks count the particles of the group from here
    MUT_COUNT[] = 0
    DEM.mutation!(model, de, group)
    @assert same_state(rng, copy(Random.default_rng())) "record_tape: predicted mutation draws of sweep $s group $g are out of step"
end
const MUT_COUNT = Ref(0)

# migration! (migration.jl:11-19): N = rand(2:G); ordered subset; per selected group one pick ∝ exp(-w) (NaN => findmin, no draw)
function record_migration!(de, groups, tape, it)
    rng = copy(Random.default_rng())
    N = rand(rng, 2:(de.n_groups))
    sub = StatsBase.sample(rng, groups, N, replace = false)
    tape.mig_n[it] = N
    for (i, g) in enumerate(sub)
        tape.mig_groups[i, it] = findfirst(q -> q === g, groups) - 1
        w = map(x -> x.weight, g)
        θ = exp.(-w) / sum(exp.(-w))
        if !any(isnan, θ)
            # StatsBase.sample(1:n, Weights(θ)) draws ONE uniform and walks the cumulative weights: record the uniform
            # itself (the device recomputes the pick from it and reports it back, demcmc_get_migration)
            r2 = copy(rng)
            tape.mig_pick_u[i, it] = rand(r2)
            StatsBase.sample(rng, 1:length(g), Weights(θ))
        end
    end
    DEM.migration!(de, groups)
    @assert same_state(rng, copy(Random.default_rng())) "record_tape: predicted migration draws of iteration $it are out of step"
end

# the hooks: the reference's own evaluate_fitness! / update_particle!, observed
function make_hooks(de, tape)
    ef, up, sm = de.evaluate_fitness!, de.update_particle!, de.sample
    function eval_hook(de_, model_, proposal)
        ef(de_, model_, proposal)
        c = CUR[]
        p = MUT_ACTIVE[] ? c.p + (MUT_COUNT[] += 1) : c.p                 # mutation!: the group's particles in order
        LASTP[] = p
        tape.prop_theta[:, p, c.s] .= flat(proposal.Θ)
        tape.prop_weight[p, c.s] = proposal.weight
        return nothing
    end
    function update_hook(de_, current, proposal, log_adj = 0.0)
        before = current.weight
        log_adj == 0.0 ? up(de_, current, proposal) : up(de_, current, proposal, log_adj)
        c = CUR[]; p = LASTP[]
        tape.log_adj[p, c.s] = log_adj
        tape.accepted[p, c.s] = current.accept[de_.iter] ? 1 : 0
        # the predicted uniform must reproduce the reference's decision
        u = tape.u_acc[p, c.s]
        @assert (u <= min(1.0, exp(proposal.weight - before + log_adj))) == current.accept[de_.iter] "record_tape: accept uniform of sweep $(c.s) particle $p does not reproduce the decision"
        return nothing
    end
    function sample_hook(de_, pool, n, replace)
        out = sm(de_, pool, n, replace)
        exp_ = EXPECT_DONORS[]
        isempty(exp_) || @assert all(a === b for (a, b) in zip(out, exp_)) "record_tape: predicted donors differ from the reference's"
        return out
    end
    return eval_hook, update_hook, sample_hook
end
const LASTP = Ref(0)
const MUT_ACTIVE = Ref(false)

"""
    record_run(model, de, n_iter, dir) -> chains

The reference's `_sample` loop (main.jl:22-42) with `stepfun = step!` (main.jl:84-89), recorded.
"""
function record_run(model, de, n_iter, dir)
    groups = DEM.sample_init(model, de, n_iter)
    Θ1 = groups[1][1].Θ
    d = length(flat(Θ1))
    blocked0 = de.blocking_on(de)
    B = blocked0 ? length(de.blocks) : 1
    tape = Tape(de.n_groups, de.Np, d, B, n_iter)
    TAPE[] = tape
    theta0 = reduce(hcat, (flat(p.Θ) for g in groups for p in g))        # [d, P] == C [P][d]
    weight0 = Float64[p.weight for g in groups for p in g]
    init_rows = de.n_initial > 0 ? Float64[flat(de.samples[i, :, p])[k] for k = 1:d, p = 1:(de.n_groups * de.Np), i = 1:(de.n_initial)] : zeros(0)
    eh, uh, sh = make_hooks(de, tape)
    de.evaluate_fitness! = eh; de.update_particle! = uh; de.sample = de.sample === DEM.resample ? de.sample : sh
    for it = 1:n_iter
        de.iter = it + de.n_initial                                        # main.jl:34
        # step! (main.jl:84-89)
        tape.mig_u[it] = rand(copy(Random.default_rng()))
        if rand() <= de.α
            record_migration!(de, groups, tape, it)
        end
        blocked = de.blocking_on(de)
        @assert blocked == blocked0 "record_tape: blocking_on must be constant over a recorded run"
        for (g, group) in enumerate(groups)                                # update! (main.jl:161-167): map over groups, serial
            for b = 1:B                                                    # block_update! (main.jl:174-179)
                s = (it - 1) * B + b
                block = blocked ? de.blocks[b] : nothing
                if rand() <= de.β                                          # mutate_or_crossover! (main.jl:199-207)
                    MUT_ACTIVE[] = true
                    record_mutation!(model, de, group, g, tape, s)         # ignores the block (main.jl:205)
                    MUT_ACTIVE[] = false
                else
                    for (j, pt) in enumerate(group)                        # crossover!(model, de, group[, block]) (crossover.jl:12-17, 61-66)
                        record_crossover!(model, de, group, pt, block, tape, s, (g - 1) * de.Np + j)
                    end
                end
            end
        end
        DEM.store_samples!(de, groups)
        for (c, p) in enumerate(vcat(groups...))
            tape.state_theta[:, c, it] .= flat(p.Θ); tape.state_id[c, it] = p.id - 1
        end
    end
    write_tape(dir, tape, theta0, weight0, init_rows, de)
    return DEM.bundle_samples(model, de, groups, n_iter)
end

# raw little-endian arrays + a manifest; Julia's column-major [a, b, c] is C's [c][b][a], which is the layout of
# demcmc_tape (tests/tape_io.py reads it back)
function write_tape(dir, t::Tape, theta0, weight0, init_rows, de)
    mkpath(dir)
    arrays = ("mig_u" => t.mig_u, "mig_n" => t.mig_n, "mig_groups" => t.mig_groups, "mig_pick_u" => t.mig_pick_u, "kind" => t.kind,
        "idx" => t.idx, "idx_row" => t.idx_row, "gamma1" => t.gamma1, "gamma2" => t.gamma2, "u_acc" => t.u_acc, "noise" => t.noise,
        "keep" => t.keep, "prop_theta" => t.prop_theta, "prop_weight" => t.prop_weight, "log_adj" => t.log_adj, "accepted" => t.accepted,
        "state_theta" => t.state_theta, "state_id" => t.state_id, "theta0" => theta0, "weight0" => weight0, "init_rows" => init_rows)
    open(joinpath(dir, "manifest.txt"), "w") do io
        println(io, "format demcmc_tape_v1")
        println(io, "G $(t.G) Np $(t.Np) d $(t.d) B $(t.B) n_iter $(t.n_iter) n_initial $(de.n_initial) burnin $(de.burnin)")
        println(io, "alpha $(de.α) beta $(de.β) eps $(de.ϵ) sigma $(de.σ) kappa $(de.κ) theta_snooker $(de.θsnooker)")
        for (name, a) in arrays
            println(io, "array $name $(eltype(a)) $(join(reverse(size(a)), ' '))")      # C-order shape
            write(joinpath(dir, name * ".bin"), a)
        end
    end
end

if abspath(PROGRAM_FILE) == @__FILE__
    out = length(ARGS) >= 1 ? ARGS[1] : "tape_out"
    n_iter = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 50
    Random.seed!(50514)                                                    # Examples/Gaussian_Example.jl
    data = rand(Normal(0.0, 1.0), 50)
    prior_loglike(μ, σ) = logpdf(Normal(0, 1), μ) + logpdf(truncated(Cauchy(0, 1), 0, Inf), σ)
    sample_prior() = [rand(Normal(0, 1)), rand(truncated(Cauchy(0, 1), 0, Inf))]
    loglike(data, μ, σ) = sum(logpdf.(Normal(μ, σ), data))
    model = DEModel(; sample_prior, prior_loglike, loglike, data, names = (:μ, :σ))
    de = DE(; sample_prior, bounds = ((-Inf, Inf), (0.0, Inf)), burnin = 20, Np = 6, θsnooker = 0.1)
    record_run(model, de, n_iter, out)
    write(joinpath(out, "data_x.bin"), data)
    println("tape written to $out: replay it with tests/tape_io.py (load_tape) + Handle.replay")
end
