"""Second-pass kernels of round 2 under compute-sanitizer (memcheck): the one-pass wide proposal (blocks, snooker, kappa, DE-MCz
donors, an odd vector length), the 128 x 12 accept, plan records, two lanes, the short-stream k_xdot with two dimension splits
per CTA, and the single-CTA chunk kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["DEMCMC_PERSIST"] = "0"                     # the level-by-level path: that is where these kernels are
import numpy as np
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
from common import make_case, hier_blocks
rng = np.random.default_rng(11)
case = make_case("hier_normal", rng, n_obs=12, n_subjects=297)                 # d = 300
for kw in (dict(blocks=hier_blocks(297), theta_snooker=0.2, alpha=0.4, burnin=3), dict(kappa=0.7, theta_snooker=0.15, burnin=6),
           dict(resample=True, n_initial=4, theta_snooker=0.2, burnin=2)):
    with case.handle(4, 9, seed=4, **kw) as h:
        if kw.get("resample"):
            h.set_history(np.stack([case.theta0(rng, 36) for _ in range(4)])); h.set_state(None)
        else:
            h.set_state(case.theta0(rng, 36))
        h.run(10)
        c = h.counters()
        print("hier d=300", sorted(kw), "launches", c["kernel_launches"], "accept", float(h.accept().mean()), "finite", bool(np.isfinite(h.samples()).all()))
case = make_case("mvnormal", rng, n_obs=130, n_dim=300)                         # three observation tiles, six dimension splits
with case.handle(2, 10, seed=2, theta_snooker=0.1, burnin=2) as h:
    h.set_state(case.theta0(rng, 20)); h.run(8)
    print("mvnormal d=301 accept", float(h.accept().mean()))
for model in ("gaussian", "lnr", "lba", "binomial"):
    case = make_case(model, rng)
    with case.handle(4, 6, seed=5, burnin=6, theta_snooker=0.2) as h:
        h.set_state(case.theta0(rng, 24)); h.run(30)
        print(model, "single-CTA chunks: launches", h.counters()["kernel_launches"], "accept", float(h.accept().mean()))
