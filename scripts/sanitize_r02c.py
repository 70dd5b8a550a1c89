"""Last-pass additions of round 2 under compute-sanitizer (memcheck): k_diag_partial with its split chain in opted-in shared
memory (more than 8192 stored rows) and with lag0 > 0 (batches of lags), and -- on a box with two GPUs -- the peer-access
history gather of `sample = resample` on a multi-device handle.
    compute-sanitizer --tool memcheck python scripts/sanitize_r02c.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
from common import make_case
rng = np.random.default_rng(3)
case = make_case("gaussian", rng)
for lags, n_iter in ((None, 9000), ("8", 200)):
    if lags: os.environ["DEMCMC_DIAG_LAGS"] = lags
    with case.handle(2, 4, seed=4, burnin=0, alpha=0.4) as h:
        h.set_state(case.theta0(rng, 8)); h.run(n_iter)
        rhat, ess = h.diagnostics(3, h.n_rows - 3)
        print("diagnostics over", h.n_rows - 3, "rows, lag batches of", lags or 4096, ": rhat", rhat, "ess", ess)
os.environ.pop("DEMCMC_DIAG_LAGS", None)
if D._ffi.lib().demcmc_device_count() >= 2:
    case = make_case("mvnormal", rng)
    rows = np.stack([case.theta0(rng, 24) for _ in range(5)])
    outs = []
    for devices in (None, [0, 1]):
        with case.handle(4, 6, seed=8, burnin=4, alpha=0.6, theta_snooker=0.2, resample=True, n_initial=5, devices=devices) as h:
            h.set_history(rows); h.set_state(None); h.run(20)
            outs.append(h.samples())
    print("resample on a multi-device handle: identical to one device:", bool(np.array_equal(outs[0], outs[1])))
else:
    print("one GPU: the multi-device gather is not exercised")
