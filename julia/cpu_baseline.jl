# cpu_baseline.jl -- the reference's own CPU path on BASELINE.json's configs, timed the way bench.py times the
# B200 path: particle-updates/s = n_groups * Np * n_iter / wall time of sample(model, de, MCMCThreads(), n_iter)
# (src/main.jl:62-71 -> pstep! -> p_update!: one task per group, src/main.jl:135-148), plus ESS/s.
#
# STATUS: NOT EXECUTED.  There is no Julia toolchain in the build image or on the GPU box (SURVEY.md 0.2), so the
# "reference, Julia threads" column of BASELINE.md section 3 is empty and bench.py's reference arm times the C
# restatement (oracle/) instead.  Run this where Julia exists:
#
#     julia -t auto --project=/path/to/DifferentialEvolutionMCMC.jl julia/cpu_baseline.jl [c1|c2|c3|c4] [n_iter]
#
# and paste the JSON line into BASELINE.md.  The useful thread count is min(n_groups, Threads.nthreads()).
# Data and model definitions follow Examples/*.jl with the sizes of BASELINE.json (SURVEY.md 8d); seeds are the
# scripts' own numbers, but Julia's RNG stream differs from numpy's, so only the SHAPES match bench.py's data.
using DifferentialEvolutionMCMC, Random, Distributions, LinearAlgebra, Statistics
using MCMCChains: ess_rhat

config = length(ARGS) >= 1 ? ARGS[1] : "c2"

function gaussian_case()                       # Examples/Gaussian_Example.jl
    Random.seed!(50514)
    data = rand(Normal(0.0, 1.0), 50)
    prior_loglike(μ, σ) = logpdf(Normal(0, 1), μ) + logpdf(truncated(Cauchy(0, 1), 0, Inf), σ)
    sample_prior() = [rand(Normal(0, 1)), rand(truncated(Cauchy(0, 1), 0, Inf))]
    loglike(data, μ, σ) = sum(logpdf.(Normal(μ, σ), data))
    model = DEModel(; sample_prior, prior_loglike, loglike, data, names = (:μ, :σ))
    de = DE(; sample_prior, bounds = ((-Inf, Inf), (0.0, Inf)), burnin = 1000, Np = 6)
    return model, de, 2000
end

function mvn_case(; d = 50, n_obs = 100_000, n_groups = 4, Np = 256)      # Examples/Multivariate_Guassian_Example.jl
    Random.seed!(50514)
    μs = rand(Normal(0.0, 1.0), d)
    data = rand(MvNormal(μs, 1.0 * I), n_obs)                            # d x n_obs
    prior_loglike(μ, σ) = sum(logpdf.(Normal(0, 1), μ)) + logpdf(truncated(Cauchy(0, 1), 0, Inf), σ)
    sample_prior() = [rand(Normal(0, 1), d), rand(truncated(Cauchy(0, 1), 0, Inf))]
    loglike(data, μ, σ) = sum(logpdf(MvNormal(μ, σ^2 * I), data))
    model = DEModel(; sample_prior, prior_loglike, loglike, data, names = (:μ, :σ))
    de = DE(; sample_prior, bounds = ((-Inf, Inf), (0.0, Inf)), burnin = 0, Np, n_groups, θsnooker = 0.1)
    return model, de, 20
end

function hier_case(; S = 1000, n = 50, n_groups = 16, Np = 512)           # Examples/Hierarchical_Example.jl
    Random.seed!(9528)
    β0 = rand(Normal(0, 1), S)
    data = [rand(Normal(1.0 + β0[s], 0.5), n) for s = 1:S]
    function prior_loglike(μβ0, σβ0, β0, σ)
        return logpdf(Normal(1, 1), μβ0) + logpdf(truncated(Cauchy(0, 1), 0, Inf), σβ0) + sum(logpdf.(Normal(0, σβ0), β0)) +
               logpdf(truncated(Cauchy(0, 1), 0, Inf), σ)
    end
    sample_prior() = (σβ0 = rand(truncated(Cauchy(0, 1), 0, Inf)); Any[rand(Normal(1, 1)), σβ0, rand(Normal(0, σβ0), S), rand(truncated(Cauchy(0, 1), 0, Inf))])
    loglike(data, μβ0, σβ0, β0, σ) = sum(sum(logpdf.(Normal(μβ0 + β0[s], σ), data[s])) for s = 1:length(data))
    model = DEModel(; sample_prior, prior_loglike, loglike, data, names = (:μβ0, :σβ0, :β0, :σ))
    blocks = [[true, true, fill(false, S), true], [false, false, fill(true, S), false]]
    de = DE(; sample_prior, bounds = ((-Inf, Inf), (0.0, Inf), (-Inf, Inf), (0.0, Inf)), burnin = 0, Np, n_groups, blocking_on = x -> true, blocks)
    return model, de, 4
end

model, de, default_iter = config == "c1" ? gaussian_case() : config == "c2" ? mvn_case() : config == "c4" ? hier_case() :
                          config == "c5shard" ? mvn_case(; d = 100, n_groups = 8, Np = 4096) : error("c1 | c2 | c4 | c5shard (c3 needs SequentialSamplingModels: see Examples/Run_LBA.jl)")
n_iter = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : default_iter
sample(model, de, MCMCThreads(), 2)                                       # compile
t = @elapsed chains = sample(model, de, MCMCThreads(), n_iter; progress = false)
B = de.blocking_on(de) ? length(de.blocks) : 1
updates = de.n_groups * de.Np * n_iter * B
ess = n_iter - de.burnin >= 100 ? minimum(skipmissing(ess_rhat(chains).nt.ess)) : missing
println("""{"impl": "reference (Julia $(VERSION), $(Threads.nthreads()) threads, useful $(min(de.n_groups, Threads.nthreads())))", "config": "$config", """ *
        """"particle_updates_per_s": $(updates / t), "seconds": $t, "n_iter": $n_iter, "ess_per_s": $(ismissing(ess) ? "null" : ess / t)}""")
