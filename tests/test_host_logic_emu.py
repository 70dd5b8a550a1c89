"""Host logic of the library (planner levels, row bookkeeping, tape sharding, migration cycle, ABI
error paths, the Python mirror of DE/DEModel/sample) exercised on CPU through the host-only test
double of the device backend (tests/emu), and checked against the oracle.  The CUDA kernels
themselves are checked by the -m gpu tests."""
import numpy as np
import pytest

import common
from common import ALL_MODELS, D, O, compare_run, forced_run, hier_blocks, make_case, rel_err

pytestmark = pytest.mark.usefixtures("emu")

VARIANTS = {
    "default": dict(),
    "snooker_kappa": dict(theta_snooker=0.3, kappa=0.8),
    "fixed_gamma": dict(proposal="fixed_gamma"),
    "variable_gamma": dict(proposal="variable_gamma", theta_snooker=0.1),
    "more_mutation": dict(beta=0.5, alpha=0.4),
}


def check(r, out, rtol=1e-12):
    assert np.array_equal(out["accept"], r["accept"])
    assert np.array_equal(out["trace"]["accepted"], r["trace"]["accepted"])
    assert rel_err(out["trace"]["prop_theta"], r["trace"]["prop_theta"]) <= rtol
    assert rel_err(out["trace"]["prop_weight"], r["trace"]["prop_weight"]) <= rtol
    assert rel_err(out["trace"]["log_adj"], r["trace"]["log_adj"]) <= 1e-9
    assert rel_err(out["samples"], r["samples"]) <= rtol
    assert rel_err(out["lp"], r["lp"]) <= rtol
    assert np.array_equal(out["state"][2], r["final_id"])
    assert rel_err(out["state"][0], r["final_theta"]) <= rtol
    assert np.array_equal(out["mig"], r["tape"]["mig_slots"])


@pytest.mark.parametrize("mode", ["replay", "native"])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_gaussian_variants(mode, variant):
    case = make_case("gaussian", np.random.default_rng(11))
    r, out = compare_run(case, 4, 6, 120, mode, burnin=60, **VARIANTS[variant])
    check(r, out)


@pytest.mark.parametrize("mode", ["replay", "native"])
@pytest.mark.parametrize("model", ALL_MODELS)
def test_all_models(mode, model):
    case = make_case(model, np.random.default_rng(12))
    r, out = compare_run(case, 3, 5, 40, mode, burnin=20, theta_snooker=0.15, alpha=0.3)
    check(r, out)


@pytest.mark.parametrize("mode", ["replay", "native"])
def test_blocking(mode):
    case = make_case("hier_normal", np.random.default_rng(13))
    r, out = compare_run(case, 2, 8, 40, mode, burnin=20, blocks=hier_blocks(9), theta_snooker=0.2, alpha=0.3)
    check(r, out)
    # a mutation sweep ignores the block mask (main.jl:205): some mutation proposal moves a masked element
    kind = r["tape"]["kind"]
    assert (kind == O.KIND_MUTATION).any()


def test_levels_reproduce_sequential_sweep():
    """The level schedule must give the in-place sweep's result for many particles per group."""
    case = make_case("gaussian", np.random.default_rng(14))
    r, out = compare_run(case, 2, 64, 12, "replay", burnin=6, theta_snooker=0.2)
    check(r, out)
    assert out["counters"]["levels"] > 12 * 3      # several levels per sweep


def test_sharded_handles_match_single(emu):
    """Two handles holding half of the groups each, with alpha = 0 (no exchange), reproduce the
    single-handle run: the draw map is keyed by global positions."""
    case = make_case("gaussian", np.random.default_rng(15))
    G, Np, n = 4, 6, 30
    theta0 = case.theta0(np.random.default_rng(1), G * Np)
    full = case.handle(G, Np, seed=3, burnin=10, alpha=0.0)
    full.set_state(theta0); full.run(n)
    ref = full.history_by_slot()
    full.close()
    for half in range(2):
        h = case.handle(G, Np, seed=3, burnin=10, alpha=0.0, group_begin=2 * half, group_count=2)
        h.set_state(theta0[half * 12:(half + 1) * 12]); h.run(n)
        th, w, ids, acc = h.history_by_slot()
        h.close()
        assert np.array_equal(th, ref[0][:, half * 12:(half + 1) * 12])
        assert np.array_equal(acc, ref[3][:, half * 12:(half + 1) * 12])
        assert np.array_equal(ids, ref[2][:, half * 12:(half + 1) * 12])


def test_teacher_forced_stepping():
    """replay one iteration at a time with set_state in between (how long GPU runs are checked)."""
    case = make_case("lnr", np.random.default_rng(16))
    G, Np, n = 3, 5, 10
    theta0 = case.theta0(np.random.default_rng(2), G * Np)
    cfg = case.oracle_config(G, Np, seed=9, burnin=5, theta_snooker=0.2)
    r = O.run(cfg, case.oracle_model(), theta0, n)
    h = case.handle(G, Np, seed=9, burnin=5, theta_snooker=0.2, trace=True)
    h.set_state(theta0)
    tape = r["tape"]
    for it in range(n):
        one = {k: (v[it:it + 1] if v is not None else None) for k, v in tape.items()}
        if it > 0:
            h.set_state(r["trace"]["state_theta"][it - 1], r["trace"]["state_id"][it - 1])
        h.replay(one, 1)
        th, w, ids = h.get_state()
        assert np.array_equal(ids, r["trace"]["state_id"][it])
        assert rel_err(th, r["trace"]["state_theta"][it]) <= 1e-12
        assert rel_err(w, r["trace"]["state_weight"][it]) <= 1e-12
    h.close()


def test_abi_error_paths():
    case = make_case("gaussian", np.random.default_rng(17))
    with pytest.raises(D._ffi.DemcmcError, match="Np must be >= 3"):
        D.Handle(2, 2, 2, case.lo, case.hi)
    with pytest.raises(D._ffi.DemcmcError, match="resample needs n_initial"):
        D.Handle(2, 4, 2, case.lo, case.hi, resample=True)
    # resample on a shard of the groups reads every particle id's history: the replicated copy needs the communicator
    hs = D.Handle(2, 4, 2, case.lo, case.hi, resample=True, n_initial=3, group_begin=0, group_count=1)
    hs.set_model("gaussian", case.prior, x=case.data["x"])
    with pytest.raises(D._ffi.DemcmcError, match="demcmc_comm_init must come first"):
        hs.set_history(np.ones((3, 8, 2)))
    hs.close()
    h = D.Handle(2, 4, 2, case.lo, case.hi, n_initial=2)
    h.set_model("gaussian", case.prior, x=case.data["x"])
    with pytest.raises(D._ffi.DemcmcError, match="null theta"):
        h.set_state(None)
    h.set_state(np.ones((8, 2)))
    with pytest.raises(D._ffi.DemcmcError, match="set_history"):
        h.run(1)
    h.close()
    h = D.Handle(2, 4, 2, case.lo, case.hi)
    with pytest.raises(D._ffi.DemcmcError, match="set_model"):
        h.set_state(np.zeros((8, 2)))
    with pytest.raises(D._ffi.DemcmcError, match="expects d"):
        h.set_model("mvnormal", case.prior, x=np.zeros((5, 4)))
    with pytest.raises(D._ffi.DemcmcError, match="no registered kernel"):
        h.set_model(17, case.prior, x=np.zeros(5))
    h.set_model("gaussian", case.prior, x=case.data["x"])
    with pytest.raises(D._ffi.DemcmcError, match="set_state"):
        h.run(1)
    h.close()


def test_eval_matches_oracle():
    for model in ALL_MODELS:
        case = make_case(model, np.random.default_rng(18))
        th = case.theta0(np.random.default_rng(3), 9)
        th[0, -1] = -1.0                     # out of bounds for every model's last parameter
        h = case.handle(3, 3)
        ll, pr = h.eval(th)
        h.close()
        m, cfg = case.oracle_model(), case.oracle_config(3, 3)
        for i in range(9):
            post = O.posterior(cfg, m, th[i])
            if np.isfinite(post):
                assert np.isclose(ll[i], O.loglike(m, th[i]), rtol=1e-12)
                assert np.isclose(pr[i], O.prior_loglike(m, th[i]), rtol=1e-12)
        assert pr[0] == -np.inf


# ---- the Python mirror of the reference's API --------------------------------------------------
def test_sample_api_gaussian_posterior():
    """test/gaussian_tests.jl restated: Normal(mu,sigma), 50 obs, DE(burnin=1500, Np=6), 3000 it;
    the NUTS comparison is replaced by 2-D quadrature of the same posterior."""
    rng = np.random.default_rng(973536)
    data = rng.normal(0.0, 1.0, 50)
    model = D.DEModel(sample_prior=lambda: [rng.normal(0, 10), abs(rng.standard_cauchy())],
                      prior_loglike=D.GPUPrior(D.Normal(0, 10), D.HalfCauchy(0, 1)),
                      loglike=D.GPULoglike("gaussian", data), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), burnin=1500, Np=6, seed=1)
    chains = D.sample(model, de, 3000)
    assert len(chains) == 1500 and chains.names == ["μ", "σ", "acceptance", "lp"]
    mu = np.linspace(-1.5, 1.5, 401); sg = np.linspace(0.5, 2.2, 401)
    M, S = np.meshgrid(mu, sg, indexing="ij")
    from scipy import stats
    lp = -50 * np.log(S) - ((data[None, None, :] - M[..., None]) ** 2).sum(-1) / (2 * S ** 2) + stats.norm(0, 10).logpdf(M) + stats.halfcauchy.logpdf(S)
    w = np.exp(lp - lp.max()); w /= w.sum()
    mean = np.array([(w * M).sum(), (w * S).sum()])
    sd = np.sqrt(np.array([(w * M ** 2).sum(), (w * S ** 2).sum()]) - mean ** 2)
    assert np.allclose(chains.mean(), mean, atol=0.02)
    assert np.allclose(chains.std(), sd, atol=0.02)
    assert np.all(np.abs(chains.rhat() - 1.0) < 0.05)
    # discard_burnin = false keeps every iteration (test/utility_tests.jl:32-39)
    de2 = D.DE(sample_prior=model.sample_prior, bounds=de.bounds, burnin=100, Np=4, discard_burnin=False, seed=2)
    assert len(D.sample(model, de2, D.MCMCThreads(), 250)) == 250


def test_sample_api_lnr_posterior(emu):                 # test/lognormal_race_tests.jl at its own size
    common.lnr_posterior_check()


def test_sample_api_blocking_posterior():               # test/blocking_tests.jl at its own size
    common.blocking_posterior_check()


def test_sample_api_blocking_on_function():
    """DE(blocking_on = de -> ..., blocks = ...): the wrapper evaluates the function for every iteration with
    de.iter = iter + n_initial (src/main.jl:34,137,162) and hands the schedule to the library -- the chains
    equal those of a handle given the same schedule explicitly."""
    def make():
        rng = np.random.default_rng(12)
        data = np.random.default_rng(13).normal(0.4, 1.3, 40)
        model = D.DEModel(sample_prior=lambda: [rng.normal(0, 3), abs(rng.standard_cauchy()) + 0.1],
                          prior_loglike=D.GPUPrior(D.Normal(0, 10), D.HalfCauchy(0, 1)),
                          loglike=D.GPULoglike("gaussian", data), names=("μ", "σ"))
        return model, data
    model, data = make()
    seen = []
    on = lambda de: (seen.append(de.iter), de.iter % 3 == 0)[1]
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), burnin=10, Np=5, n_groups=2, seed=4,
              blocking_on=on, blocks=[[True, False], [False, True]], discard_burnin=False)
    chains = D.sample(model, de, 30)
    assert seen[:30] == list(range(1, 31))
    model2, _ = make()
    theta0 = np.array([[*model2.sample_prior()] for _ in range(11)])[1:]        # build_handle draws one for the shapes first
    h = D.Handle(2, 5, 2, [-np.inf, 0.0], [np.inf, np.inf], burnin=10, seed=4, blocks=np.array([[1, 0], [0, 1]], dtype=np.uint8),
                 blocking_schedule=[1 if it % 3 == 0 else 0 for it in range(1, 31)])
    h.set_model("gaussian", [("normal", 0, 10), ("halfcauchy", 0, 1)], x=data)
    h.set_state(theta0)
    h.run(30)
    assert h.counters()["sweeps"] == 30 + 10
    ref = h.chains(0, 30)
    h.close()
    assert np.array_equal(chains.value, ref.transpose(2, 1, 0))


def test_closure_raises_instead_of_cpu_fallback():
    model = D.DEModel(sample_prior=lambda: [0.0, 1.0], prior_loglike=D.GPUPrior(D.Normal(), D.HalfCauchy()),
                      loglike=lambda data, mu, sigma: 0.0, names=("μ", "σ"), data=np.zeros(3))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-1, 1), (0, 1)), Np=4)
    with pytest.raises(TypeError, match="no CPU fallback"):
        D.sample(model, de, 10)
    model2 = D.DEModel(sample_prior=lambda: [0.0, 1.0], prior_loglike=lambda mu, sigma: 0.0,
                       loglike=D.GPULoglike("gaussian", np.zeros(3)), names=("μ", "σ"))
    with pytest.raises(TypeError, match="GPUPrior"):
        D.sample(model2, de, 10)
    with pytest.raises(TypeError, match="custom host function"):
        D.DE(sample_prior=model.sample_prior, bounds=((-1, 1),), Np=4, sample=lambda *a: None)
    with pytest.raises(ValueError, match="n_initial"):
        D.DE(sample_prior=model.sample_prior, bounds=((-1, 1),), Np=4, sample=D.resample)


def test_names_blocks_and_bounds_flattening():
    from demcmc_b200.api import _expand_block, _flat_names, _flatten
    theta = [np.array([[1.0, 2.0], [3.0, 4.0]]), 5.0, np.array([6.0, 7.0])]
    assert _flatten(theta) == [1.0, 3.0, 2.0, 4.0, 5.0, 6.0, 7.0]               # column-major (main.jl:236-238)
    shapes = [np.shape(v) for v in theta]
    assert _flat_names(("A", "s", "v"), shapes) == ["A[1,1]", "A[2,1]", "A[1,2]", "A[2,2]", "s", "v[1]", "v[2]"]
    assert _expand_block([[[True, False], [False, True]], False, True], shapes) == [True, False, False, True, False, True, True]


def test_overlapped_chunks_do_not_change_the_result():
    """Consecutive sweeps share launches (planner.h); the chain must not depend on the chunk length."""
    case = make_case("gaussian", np.random.default_rng(19))
    theta0 = case.theta0(np.random.default_rng(4), 2 * 48)
    outs = []
    for chunk in (1, 3, 16):
        h = case.handle(2, 48, seed=21, burnin=5, theta_snooker=0.2, alpha=0.05)
        h.set_max_chunk(chunk)
        h.set_lanes(1)
        h.set_state(theta0)
        h.run(40)
        outs.append((h.samples(), h.accept(), h.lp(), h.counters()["levels"]))
        h.close()
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
    assert outs[2][3] < 0.8 * outs[0][3]       # fewer, fatter levels
    # and it still equals the sequential in-place sweep of the oracle
    cfg = case.oracle_config(2, 48, seed=21, burnin=5, theta_snooker=0.2, alpha=0.05, base_snapshot=1)
    r = O.run(cfg, case.oracle_model(), theta0, 40, record=False, trace=False)
    assert np.array_equal(outs[2][0], r["samples"]) and np.array_equal(outs[2][1], r["accept"])


def test_device_bundle_matches_bundle_samples():
    """demcmc_get_chains = bundle_samples (main.jl:222-250) including its by-position quirk: after
    migrations the acceptance/lp columns of chain c belong to the particle at final position c."""
    from demcmc_b200.api import DE, DEModel, GPULoglike, GPUPrior, HalfCauchy, Normal, bundle_samples
    case = make_case("gaussian", np.random.default_rng(23))
    theta0 = case.theta0(np.random.default_rng(5), 3 * 7)
    h = case.handle(3, 7, seed=5, burnin=10, alpha=0.6, theta_snooker=0.1)
    h.set_state(theta0)
    h.run(30)
    ids = h.get_state()[2]
    assert not np.array_equal(ids, np.arange(21))                      # migrations moved particles
    model = DEModel(sample_prior=lambda: [0.0, 1.0], prior_loglike=GPUPrior(Normal(), HalfCauchy()),
                    loglike=GPULoglike("gaussian", np.zeros(3)), names=("mu", "sigma"))
    de = DE(sample_prior=model.sample_prior, bounds=((-1, 1), (0, 1)), n_groups=3, Np=7, burnin=10)
    ref = bundle_samples(model, de, h.samples(), h.accept(), h.lp(), ids, [(), ()], 30)
    got = h.chains(10, 20)
    assert got.shape == (21, 4, 20)
    assert np.array_equal(got.transpose(2, 1, 0), ref.value)
    assert np.array_equal(h.chains(0, 30)[:, :2, :], h.samples())
    assert h.chains(30, 0).shape == (21, 4, 0)
    with pytest.raises(D._ffi.DemcmcError):
        h.chains(25, 10)
    h.close()


@pytest.mark.parametrize("mode", ["replay", "native"])
def test_blocking_on_as_a_function_of_the_iteration(mode):
    """blocking_on(de) is evaluated every iteration (main.jl:137,162): a schedule that switches block updating
    on and off -- block_update! over the blocks in some iterations, update! with all parameters in the others,
    consecutive unblocked iterations overlapped in one chunk."""
    case = make_case("hier_normal", np.random.default_rng(91))
    sched = [1, 0, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 0, 1]
    r, out = forced_run(case, 2, 8, len(sched), mode, burnin=6, blocks=hier_blocks(9), alpha=0.3, blocking_schedule=sched)
    check(r, out)
    # unforced, as one call: same chain as iteration by iteration in native mode
    if mode == "native":
        theta0 = case.theta0(np.random.default_rng(0), 16)
        h = case.handle(2, 8, seed=5, burnin=6, blocks=hier_blocks(9), alpha=0.3, blocking_schedule=sched)
        h.set_state(theta0)
        h.run(len(sched))
        assert h.counters()["sweeps"] == sum(2 if s else 1 for s in sched)
        h.close()
    # unforced and short: runs of unblocked iterations share a chunk (overlapped sweeps with the blocks' sweep stride)
    sched2 = [0, 0, 0, 1, 0, 0, 0, 0]
    r, out = compare_run(case, 2, 8, len(sched2), mode, burnin=3, blocks=hier_blocks(9), alpha=0.3, blocking_schedule=sched2)
    assert np.array_equal(out["accept"], r["accept"])
    assert rel_err(out["samples"], r["samples"]) < 1e-9


def test_checkpoint_resume_is_exact():
    """run(a) + get_state + a NEW handle (set_state, set_weights, set_iteration) + run(b) = run(a + b): the
    Philox counters, the burn-in switch and the migration schedule continue."""
    for model, kw in (("mvnormal", dict(theta_snooker=0.2, alpha=0.3, burnin=14)), ("gaussian", dict(alpha=0.4, burnin=3, kappa=0.9))):
        case = make_case(model, np.random.default_rng(83))
        G, Np, a, b = 3, 8, 9, 11
        theta0 = case.theta0(np.random.default_rng(4), G * Np)
        h = case.handle(G, Np, seed=21, **kw)
        h.set_state(theta0)
        h.run(a + b)
        full = (h.history_by_slot(a, b), h.get_state())
        h.close()
        h1 = case.handle(G, Np, seed=21, **kw)
        h1.set_state(theta0)
        h1.run(a)
        th, w, ids = h1.get_state()
        h1.close()
        h2 = case.handle(G, Np, seed=21, **kw)
        h2.set_state(th, ids)
        h2.set_weights(w)
        h2.set_iteration(a)
        h2.run(b)
        res = (h2.history_by_slot(0, b), h2.get_state())
        h2.close()
        for x, y in zip(full[0] + full[1], res[0] + res[1]):
            assert np.array_equal(x, y)


def test_device_moments_match_the_chains():
    """demcmc_get_moments: pooled mean / variance per parameter over a row range, computed by the
    backend from the stored rows, against numpy on the downloaded history."""
    case = make_case("gaussian", np.random.default_rng(29))
    h = case.handle(3, 7, seed=6, burnin=5, alpha=0.4)
    h.set_state(case.theta0(np.random.default_rng(8), 21))
    h.run(40)
    th = h.history_by_slot(10, 30)[0]                                   # [30][21][d]
    cnt, mean, var = h.moments(10, 30)
    flat = th.reshape(-1, th.shape[2])
    assert cnt == flat.shape[0]
    assert np.allclose(mean, flat.mean(axis=0), rtol=1e-12, atol=1e-14)
    assert np.allclose(var, flat.var(axis=0, ddof=1), rtol=1e-10, atol=1e-14)
    assert h.moments(40, 0)[0] == 0
    with pytest.raises(D._ffi.DemcmcError):
        h.moments(35, 10)
    h.close()


# ---- de.sample = resample (DE-MCz donors from the history, crossover.jl:113-124) and n_initial ----
@pytest.mark.parametrize("mode", ["replay", "native"])
@pytest.mark.parametrize("kw", [dict(theta_snooker=0.3), dict(proposal="fixed_gamma", kappa=0.7, theta_snooker=0.1),
                                dict(alpha=0.6, theta_snooker=0.2)])
def test_resample_from_history(mode, kw):
    case = make_case("gaussian", np.random.default_rng(31))
    r, out = compare_run(case, 3, 6, 40, mode, burnin=15, n_initial=5, resample=True, **kw)
    check(r, out)
    # the first n_initial rows of de.samples are the prior rows, and the chain started from row 1
    assert np.array_equal(out["samples"][:, :, :5], r["samples"][:, :, :5])
    assert (r["tape"]["idx_row"][r["tape"]["kind"] == O.KIND_SNOOKER] >= 0).all()


@pytest.mark.parametrize("mode", ["replay", "native"])
def test_resample_blocking_hierarchical(mode):
    """Examples/Hierarchical_Example.jl: blocks + sample = resample + n_initial."""
    case = make_case("hier_normal", np.random.default_rng(32))
    r, out = compare_run(case, 2, 8, 30, mode, burnin=10, n_initial=4, resample=True, blocks=hier_blocks(9), alpha=0.3)
    check(r, out)


def test_n_initial_without_resample():
    """n_initial > 0 with the default donors: prior rows are stored, bundle_samples' window is not
    shifted by n_initial (main.jl:226-234)."""
    case = make_case("gaussian", np.random.default_rng(33))
    r, out = compare_run(case, 2, 5, 25, "native", burnin=5, n_initial=3)
    check(r, out)
    assert out["samples"].shape[2] == 28


def test_sample_api_mvn_resample():
    """test/multivariate_normal_tests.jl restated at its own size: MvNormal(mu, sigma^2 I), 30 means, 100
    observations, DE(sample = resample, n_initial = 124, Np = 3, n_groups = 1, theta_snooker = 0.1),
    50 000 iterations, and the reference's own assertions (:62-69)."""
    common.mvn_resample_check(n_iter=50_000, burnin=5000, sd_atol=0.01)


def test_lanes_do_not_change_the_result():
    """Independent sets of groups run as concurrent kernel chains (demcmc_set_lanes); the chain must
    not depend on how many there are, with or without a tape, blocking, migrations."""
    case = make_case("hier_normal", np.random.default_rng(41))
    theta0 = case.theta0(np.random.default_rng(6), 3 * 10)
    outs = []
    for lanes in (1, 2):
        h = case.handle(3, 10, seed=8, burnin=6, theta_snooker=0.2, alpha=0.4, blocks=hier_blocks(9))
        h.set_lanes(lanes)
        h.set_state(theta0)
        h.run(25)
        outs.append((h.samples(), h.accept(), h.lp(), h.get_state()[2]))
        h.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


# ---- the optimize path: maximize! / minimize! + evaluate_fun! (optimize.jl, utilities.jl:113-120,212-226) ----
@pytest.mark.parametrize("mode", ["replay", "native"])
@pytest.mark.parametrize("model,update", [("rastrigin", "minimize"), ("gaussian", "maximize"), ("mvnormal", "maximize")])
def test_optimize_updates(mode, model, update):
    case = make_case(model, np.random.default_rng(51))
    r, out = compare_run(case, 2, 6, 60, mode, burnin=20, update=update, fitness="fun", alpha=0.3)
    check(r, out)
    # maximize!/minimize! never write Particle.accept / Particle.lp
    assert not out["accept"].any() and not out["lp"].any()
    w = out["state"][1]
    assert np.all(np.isfinite(w))


def test_optimize_api():
    common.optimize_checks()
    case = make_case("gaussian", np.random.default_rng(52))
    with pytest.raises(D._ffi.DemcmcError, match="MethodError"):
        D.Handle(2, 4, 2, case.lo, case.hi, update="maximize", fitness="fun", theta_snooker=0.1)


# ---- the cross term on operands whose answer is not zero (tests/xdot_common.py) -----------------------------
@pytest.mark.parametrize("kind,n,k", [("mvnormal", 130, 50), ("mvnormal", 65, 100), ("hier_normal", 50, 120)])
def test_cross_term_with_a_given_centre(emu, kind, n, k):
    import xdot_common as X
    err, ll, case, th = X.eval_error(kind, n, k, 19, seed=k)
    assert err <= X.TOL
    m = O.Model(kind, case["d"], case["prior"], x=case["x"])
    assert common.rel_err(ll, np.array([O.loglike(m, t) for t in th])) <= 1e-11
    kw = dict(blocks=common.hier_blocks(k)) if kind == "hier_normal" else dict(theta_snooker=0.2)
    err, ctr, n_checked = X.run_error(kind, n, k, 2, 9, 4, seed=k, **kw)
    assert n_checked > 0 and err <= X.TOL


def test_cross_term_mutation_is_caught_by_the_cpu_suite(emu):
    """a test double whose cross term used the wrong dimension's mean x 3 + 17 passed all 93 CPU tests.
    DEMCMC_TEST_CORRUPT makes the double do exactly that; the parity check must now fail."""
    import os, subprocess, sys
    script = os.path.join(os.path.dirname(__file__), "xdot_common.py")
    env = dict(os.environ)
    env.pop("DEMCMC_TEST_CORRUPT", None)
    assert subprocess.run([sys.executable, script, common.EMU_LIB], env=env, capture_output=True).returncode == 0
    env["DEMCMC_TEST_CORRUPT"] = "1"
    assert subprocess.run([sys.executable, script, common.EMU_LIB], env=env, capture_output=True).returncode == 3


# ---- thinning (demcmc_config.store_every) and the multi-device handle (demcmc_config.n_devices) ---------------
def _full_outputs(h):
    return dict(samples=h.samples(), accept=h.accept(), lp=h.lp(), chains=h.chains(), state=h.get_state(), mom=h.moments(),
                slot=h.history_by_slot(), mig=h.migration_slots(), updates=h.counters()["particle_updates"])


def _same(a, b):
    if isinstance(a, (tuple, list)):
        return all(_same(x, y) for x, y in zip(a, b))
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("model,kw", [("mvnormal", dict(theta_snooker=0.2, alpha=0.4)), ("gaussian", dict(kappa=0.8)),
                                      ("hier_normal", dict(blocks=True, alpha=0.3))])
def test_thinned_run_keeps_every_kth_row_of_the_full_run(emu, model, kw):
    """store_every = k: the chain is the same chain; de.samples holds iterations k, 2k, ... only"""
    rng = np.random.default_rng(11)
    case = make_case(model, rng)
    kw = dict(kw)
    if kw.pop("blocks", False):
        kw["blocks"] = hier_blocks(9)
    G, Np, n_iter, k = 4, 7, 23, 3
    th0 = case.theta0(rng, G * Np)
    outs = []
    for every in (1, k):
        with case.handle(G, Np, seed=3, burnin=5, store_every=every, **kw) as h:
            h.set_state(th0)
            h.run(10)
            h.run(n_iter - 10)                     # two calls: the history grows in place
            assert h.n_rows == n_iter // every
            outs.append(_full_outputs(h))
    full, thin = outs
    keep = np.arange(k - 1, n_iter, k)
    assert np.array_equal(thin["samples"], full["samples"][:, :, keep])
    assert np.array_equal(thin["accept"], full["accept"][:, keep]) and np.array_equal(thin["lp"], full["lp"][:, keep])
    assert _same(thin["state"], full["state"]) and thin["updates"] == full["updates"]
    assert np.array_equal(thin["slot"][0], full["slot"][0][keep]) and np.array_equal(thin["slot"][2], full["slot"][2][keep])
    # bundle_samples of the kept rows: parameter columns are the kept draws of each id
    d = case.d
    assert np.array_equal(thin["chains"][:, :d, :], full["chains"][:, :d, :][:, :, keep])


def test_thinned_resample_runs_on_the_stored_rows(emu):
    rng = np.random.default_rng(4)
    case = make_case("mvnormal", rng)
    G, Np, n0 = 2, 5, 6
    rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)])
    with case.handle(G, Np, seed=2, burnin=4, n_initial=n0, resample=True, store_every=4, theta_snooker=0.2) as h:
        h.set_history(rows)
        h.set_state(None)
        h.run(17)
        assert h.n_rows == n0 + 4
        s = h.samples()
        assert np.array_equal(s[:, :, :n0], rows.transpose(1, 2, 0)) and np.isfinite(s).all()


@pytest.mark.parametrize("n_dev", [2, 4])
@pytest.mark.parametrize("model,kw", [("mvnormal", dict(theta_snooker=0.2, alpha=0.6)), ("lnr", dict(alpha=0.5)),
                                      ("hier_normal", dict(blocks=True, alpha=0.5, store_every=2)),
                                      ("mvnormal", dict(resample=True, n_initial=5, theta_snooker=0.2, alpha=0.6)),
                                      ("hier_normal", dict(blocks=True, resample=True, n_initial=4, alpha=0.5))])
def test_multi_device_handle_is_the_single_device_chain(emu, model, kw, n_dev):
    """cfg.n_devices = N: one handle, one process, N devices (host threads of the test double here), migration through
    the mailboxes -- every output identical to the single-device handle's, bit for bit"""
    rng = np.random.default_rng(21)
    case = make_case(model, rng)
    kw = dict(kw)
    if kw.pop("blocks", False):
        kw["blocks"] = hier_blocks(9)
    G, Np, n_iter = 4, 6, 25
    th0 = case.theta0(rng, G * Np)
    rows = np.stack([case.theta0(rng, G * Np) for _ in range(kw["n_initial"])]) if kw.get("resample") else None
    outs = []
    for devices in (None, list(range(n_dev))):
        with case.handle(G, Np, seed=8, burnin=6, trace=True, devices=devices, **kw) as h:
            if rows is not None:                           # sample = resample: every device keeps a replicated history
                h.set_history(rows)
            h.set_state(th0)
            h.run(n_iter)
            o = _full_outputs(h)
            o["trace"] = h.trace()
            o["eval"] = h.eval(th0[:5])
            outs.append(o)
    one, many = outs
    assert (one["mig"] >= 0).any(), "no migration happened: the test would not cross devices"
    for key in ("samples", "accept", "lp", "chains", "state", "slot", "mig", "updates", "eval"):
        assert _same(one[key], many[key]), key
    for key in one["trace"]:
        assert np.array_equal(one["trace"][key], many["trace"][key], equal_nan=True), key
    assert one["mom"][0] == many["mom"][0] and rel_err(many["mom"][1], one["mom"][1]) < 1e-13 and rel_err(many["mom"][2], one["mom"][2]) < 1e-12


def test_multi_device_handle_replays_the_oracle_and_checks_its_arguments(emu):
    rng = np.random.default_rng(2)
    case = make_case("mvnormal", rng)
    G, Np, n_iter = 4, 5, 8
    th0 = case.theta0(rng, G * Np)
    cfg = case.oracle_config(G, Np, seed=5, burnin=3, alpha=0.5, theta_snooker=0.2)
    r = O.run(cfg, case.oracle_model(), th0, n_iter)
    with case.handle(G, Np, seed=5, burnin=3, alpha=0.5, theta_snooker=0.2, devices=[0, 1]) as h:
        h.set_state(th0)
        h.replay(r["tape"], n_iter)
        assert np.array_equal(h.accept(), r["accept"])
        assert rel_err(h.samples(), r["samples"]) < 1e-12
        with pytest.raises(D._ffi.DemcmcError):
            h.comm_init(bytes(128), 0, 2)
    for bad in (dict(devices=[0, 0]), dict(devices=[0, 1, 2])):
        with pytest.raises(D._ffi.DemcmcError):
            case.handle(G, Np, **bad)


def test_julia_tape_format_round_trip(emu, tmp_path):
    """the on-disk tape julia/record_tape.jl writes (never produced here: no Julia), written from an oracle run by
    tests/tape_io.py and replayed through the library after a round trip through the files"""
    import tape_io
    rng = np.random.default_rng(31)
    case = make_case("gaussian", rng)
    G, Np, n_iter = 4, 6, 12
    th0 = case.theta0(rng, G * Np)
    kw = dict(burnin=5, theta_snooker=0.2, kappa=0.8)
    r = O.run(case.oracle_config(G, Np, seed=5, **kw), case.oracle_model(), th0, n_iter)
    arrays = dict(r["tape"])
    arrays.update(theta0=th0, accepted=r["trace"]["accepted"] if "accepted" in r["trace"] else None)
    tape_io.save_tape(str(tmp_path), dict(G=G, Np=Np, d=case.d, B=1, n_iter=n_iter, n_initial=0, burnin=5, alpha=0.1, beta=0.1, eps=0.001,
                                          sigma=0.05, kappa=0.8, theta_snooker=0.2), arrays)
    meta, tape, extra = tape_io.load_tape(str(tmp_path))
    assert meta["G"] == G and meta["kappa"] == 0.8
    for k in tape_io.TAPE_FIELDS:
        if r["tape"].get(k) is not None and tape.get(k) is not None:
            assert np.array_equal(np.asarray(r["tape"][k]).reshape(-1), np.asarray(tape[k]).reshape(-1)), k
    with case.handle(G, Np, seed=5, **kw) as h:
        h.set_state(extra["theta0"])
        h.replay(tape, n_iter)
        assert np.array_equal(h.accept(), r["accept"])


def _check_device_diagnostics(devices=None, G=4, Np=6, n_iter=90, store_every=1):
    """demcmc_get_diagnostics against the host estimators (diagnostics.py without the rank normalisation)"""
    from demcmc_b200 import diagnostics as Dg
    rng = np.random.default_rng(12)
    case = make_case("gaussian", rng)
    th0 = case.theta0(rng, G * Np)
    with case.handle(G, Np, seed=4, burnin=0, alpha=0.4, devices=devices, store_every=store_every) as h:
        h.set_state(th0)
        h.run(n_iter)
        row0 = 7
        n = h.n_rows - row0
        rhat, ess = h.diagnostics(row0, n)
        s = h.samples()[:, :, row0:]                       # [P][d][n]: chain = particle id
    for k in range(case.d):
        x = Dg._split(s[:, k, :].T)                        # (draws, chains) -> split halves
        assert abs(rhat[k] - Dg._rhat(x)) < 1e-9 * max(1.0, abs(rhat[k])), (k, rhat[k], Dg._rhat(x))
        assert abs(ess[k] - Dg._ess(x)) < 1e-6 * ess[k], (k, ess[k], Dg._ess(x))


def test_device_diagnostics_match_the_host_estimators(emu):
    _check_device_diagnostics()
    _check_device_diagnostics(n_iter=91)                   # odd number of rows: the middle draw is dropped
    _check_device_diagnostics(devices=[0, 1])              # shards merged
    _check_device_diagnostics(store_every=2, n_iter=120)


def test_device_diagnostics_in_batches_of_lags(emu, monkeypatch):
    """the autocovariances arrive in batches of lags, a further batch only while a Geyer sequence is still positive: batches
    of 2, 6 and 16 lags on chains of 41-60 draws give what one batch of all lags gives"""
    for step in ("2", "6", "16", "3"):                     # an odd step is rounded down to whole pairs
        monkeypatch.setenv("DEMCMC_DIAG_LAGS", step)
        _check_device_diagnostics()
        _check_device_diagnostics(n_iter=127, G=2, Np=5)


# ---- the full-covariance MVN kernel and the vector-parameter Gaussian example --------------------------
def _check_mvn_full():
    from scipy import stats
    rng = np.random.default_rng(8)
    case = make_case("mvnormal_full", rng)
    th = case.theta0(rng, 9)
    # the oracle's restatement against scipy's multivariate normal
    m = case.oracle_model()
    for t in th:
        ref = stats.multivariate_normal(mean=t[:6], cov=t[6] ** 2 * case.data["cov"]).logpdf(case.data["x"]).sum()
        assert abs(O.loglike(m, t) - ref) <= 1e-11 * abs(ref)
    with case.handle(1, 9) as h:
        ll, pr = h.eval(th)
    assert rel_err(ll, np.array([O.loglike(m, t) for t in th])) <= 1e-12
    for mode in ("replay", "native"):
        r, out = forced_run(case, 2, 7, 6, mode, burnin=2, theta_snooker=0.2, alpha=0.4)
        assert np.array_equal(out["accept"], r["accept"])
        assert rel_err(out["samples"], r["samples"]) <= 1e-12 and rel_err(out["lp"], r["lp"]) <= 1e-12


def test_full_covariance_mvn_kernel(emu):
    _check_mvn_full()
    rng = np.random.default_rng(3)
    case = make_case("mvnormal_full", rng)
    bad = dict(case.data)
    bad["cov"] = -np.eye(6)
    with pytest.raises(D._ffi.DemcmcError):
        D.Handle(1, 4, 7, case.lo, case.hi).set_model("mvnormal_full", case.prior, **bad)


def _check_vector_gaussian_example():
    """Examples/Guassian_Example_Vector.jl: the Gaussian model whose loglike(data, θ...) destructures μ, σ -- whether the
    two parameters are named separately or come as ONE 2-vector parameter, the registered "gaussian" kernel is its
    kernel and the chains are the same chains"""
    x = np.random.default_rng(50514).normal(0, 1, 50)
    outs = []
    for vector in (False, True):
        rng = np.random.default_rng(9)
        draw = lambda: [rng.normal(0, 1), abs(rng.standard_cauchy())]      # noqa: E731
        if vector:
            model = D.DEModel(sample_prior=lambda: [np.array(draw())], prior_loglike=D.GPUPrior(D.Flat()), loglike=D.GPULoglike("gaussian", x), names=("θ",))
            de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf),), burnin=100, Np=6, seed=3)
        else:
            model = D.DEModel(sample_prior=draw, prior_loglike=D.GPUPrior(D.Flat(), D.Flat()), loglike=D.GPULoglike("gaussian", x), names=("μ", "σ"))
            de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (-np.inf, np.inf)), burnin=100, Np=6, seed=3)
        ch = D.sample(model, de, D.MCMCThreads(), 300)
        outs.append(ch)
    assert outs[1].names[:2] == ["θ[1]", "θ[2]"] and outs[0].names[:2] == ["μ", "σ"]
    assert np.array_equal(outs[0].value, outs[1].value, equal_nan=True)
    # and with the example's own priors and bounds it recovers the posterior (mean of μ near the sample mean)
    rng = np.random.default_rng(1)
    model = D.GPUDEModel(sample_prior=lambda: [rng.normal(0, 1), abs(rng.standard_cauchy())], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                         loglike=D.GPULoglike("gaussian", x), names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), burnin=1000, Np=6, seed=4)
    ch = D.sample(model, de, D.MCMCThreads(), 2000)
    assert abs(ch.mean()[0] - x.mean() * 50 / 51) < 0.05 and abs(ch.mean()[1] - x.std()) < 0.1


def test_vector_parameter_gaussian_example(emu):
    _check_vector_gaussian_example()


@pytest.mark.parametrize("kw", [dict(alpha=0.3), dict(alpha=0.0, proposal="fixed_gamma", kappa=0.8), dict(alpha=0.2, store_every=3)])
def test_overlapped_block_sweeps_are_a_schedule_only(emu, kw):
    """blocking_on: the block sweeps of consecutive iterations share a chunk (their dependency levels overlap); one sweep
    per chunk (max_chunk = 1) must give the same chain bit for bit -- native draws and replayed ones"""
    rng = np.random.default_rng(14)
    case = make_case("hier_normal", rng)
    G, Np, n_iter = 3, 8, 21
    th0 = case.theta0(rng, G * Np)
    outs = []
    for chunk in (1, 16, 5):
        with case.handle(G, Np, seed=6, burnin=4, blocks=hier_blocks(9), proposal=kw.get("proposal", "random_gamma"),
                         **{k: v for k, v in kw.items() if k != "proposal"}) as h:
            h.set_max_chunk(chunk)
            h.set_state(th0)
            h.run(9)
            h.run(n_iter - 9)
            outs.append((h.samples(), h.accept(), h.lp(), h.get_state(), h.counters()["levels"]))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
        assert all(np.array_equal(a, b) for a, b in zip(o[3], outs[0][3]))
    assert outs[1][4] < outs[0][4]                         # fewer, fuller levels
    # replay through overlapped block sweeps against the oracle
    r, out = compare_run(case, 2, 6, 7, "replay", seed=3, burnin=3, blocks=hier_blocks(9), proposal="fixed_gamma")
    assert np.array_equal(out["accept"], r["accept"])


def test_background_model_binding_reports_its_errors(emu):
    """sample() binds the model on a host thread while the initial particles are drawn: an error of that call must
    surface in the caller, and the handle must still be released"""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(30, 4))
    bad = D.GPULoglike("mvnormal_full", x, cov=-np.eye(4))          # not positive definite: demcmc_set_model fails
    model = D.DEModel(sample_prior=lambda: [rng.normal(size=4), 1.0], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                      loglike=bad, names=("μ", "σ"))
    de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=2, Np=4, burnin=0, seed=1)
    with pytest.raises(D._ffi.DemcmcError, match="positive definite"):
        D.sample(model, de, 5)
    good = D.DEModel(sample_prior=model.sample_prior, prior_loglike=model.prior_loglike, loglike=D.GPULoglike("mvnormal_full", x, cov=np.eye(4)), names=("μ", "σ"))
    assert len(D.sample(good, de, 5)) == 5


@pytest.mark.parametrize("model,kw", [("mvnormal", dict(theta_snooker=0.25, alpha=0.4, burnin=6)), ("gaussian", dict(kappa=0.8)),
                                      ("mvnormal", dict(resample=True, theta_snooker=0.2, burnin=3))])
def test_plan_records_equal_the_draws_made_in_place(emu, model, kw, monkeypatch):
    """DEMCMC_PLAN=1: kind, donors, gammas and the accept uniform of a whole chunk are drawn by one launch (PlanRec) and
    read back by the level kernels; the chain is the one of the in-place draws, bit for bit"""
    rng = np.random.default_rng(31)
    case = make_case(model, rng)
    kw = dict(kw)
    G, Np, n0 = 3, 6, 5
    rows = np.stack([case.theta0(rng, G * Np) for _ in range(n0)]) if kw.get("resample") else None
    if rows is not None:
        kw["n_initial"] = n0
    th0 = case.theta0(rng, G * Np)
    outs = []
    for plan in ("0", "1"):
        monkeypatch.setenv("DEMCMC_PLAN", plan)
        with case.handle(G, Np, seed=9, **kw) as h:
            if rows is not None:
                h.set_history(rows)
                h.set_state(None)
            else:
                h.set_state(th0)
            h.run(23)
            outs.append((h.samples(), h.accept(), h.lp(), h.counters()["kernel_launches"]))
    assert outs[1][3] > outs[0][3]                            # the plan launches happened
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert np.array_equal(a, b, equal_nan=True)
