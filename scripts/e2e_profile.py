"""Times the pieces of the public sample() call on the GPU box (diagnostic)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import demcmc_b200 as D
from demcmc_b200.api import build_handle, bundle_samples, _flatten
x, prior, lo, hi, theta0 = bench.workload(4)
rng = np.random.default_rng(7)
model = D.DEModel(sample_prior=lambda: [rng.normal(size=50), abs(rng.standard_cauchy())], prior_loglike=D.GPUPrior(D.Normal(0, 1), D.HalfCauchy(0, 1)),
                  loglike=D.GPULoglike("mvnormal", x), names=("μ", "σ"))
de = D.DE(sample_prior=model.sample_prior, bounds=((-np.inf, np.inf), (0.0, np.inf)), n_groups=4, Np=256, burnin=0, θsnooker=0.1, seed=11)
for rep in range(3):
    t = [time.perf_counter()]
    h, shapes, d = build_handle(model, de); t.append(time.perf_counter())
    th0 = np.array([_flatten(model.sample_prior()) for _ in range(1024)]); t.append(time.perf_counter())
    h.set_state(th0); t.append(time.perf_counter())
    h.run(100); t.append(time.perf_counter())
    s = h.samples(); t.append(time.perf_counter())
    a = h.accept(); l = h.lp(); t.append(time.perf_counter())
    _, _, ids = h.get_state(); ch = bundle_samples(model, de, s, a, l, ids, shapes, 100); t.append(time.perf_counter())
    h.close(); t.append(time.perf_counter())
    names = ["build_handle(create+set_model)", "sample_prior x1024", "set_state", "run(100)", "samples()", "accept+lp", "bundle", "close"]
    print(rep, {n: round((b - a) * 1e3, 2) for n, a, b in zip(names, t[:-1], t[1:])}, "device_ms", round(h.counters()["device_ms"] if False else 0, 2))
