"""Soak of the wide-parameter path (configs[3] shape, smaller population, level-by-level launches forced): a long native run
with two lanes, plan records, the one-pass proposal and the short-stream k_xdot against the same run with one lane, in-place
draws, the three-pass proposal and the streaming k_xdot -- accept decisions, ids and states must be identical (the states are
copies of proposals whose arithmetic is the same; only the mean-square sums differ in order, which would show as a flipped
decision); and the default configuration twice: bit-identical.  Every variant runs in its own process (some switches are read
once per process)."""
import hashlib, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300


def child():
    import numpy as np
    import bench_configs as BC
    import demcmc_b200 as D
    D._ffi.use_library(D._ffi.DEFAULT_LIB)
    c = BC.config("c4")
    G, Np = 8, 128
    theta0 = c["theta0"](np.random.default_rng(1), G * Np)
    with D.Handle(G, Np, c["d"], c["lo"], c["hi"], seed=20261017, **c["kw"]) as h:
        h.set_model(c["kind"], c["prior"], **c["data"])
        h.set_state(theta0)
        t0 = time.perf_counter(); h.run(n_iter); dt = time.perf_counter() - t0
        th, w, ids = h.get_state()
        acc = h.accept()
        ctr = h.counters()
        assert ctr["persistent_chunks"] == 0
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
        print(json.dumps({"acc": sha(acc), "ids": sha(ids), "state": sha(th), "mean_abs": float(np.abs(th).mean()), "launches": ctr["kernel_launches"], "levels": ctr["levels"], "s": round(dt, 3)}))


if "--child" in sys.argv:
    child()
    sys.exit(0)
old = {"DEMCMC_LANES": "1", "DEMCMC_PLAN": "0", "DEMCMC_WIDE_SHAPE": "0", "DEMCMC_XD_SHORT": "0"}
out = {}
for name, env in (("default", {}), ("default again", {}), ("one lane, in-place draws, three-pass proposal, streaming k_xdot", old)):
    e = {k: v for k, v in os.environ.items() if k not in old}
    e.update(env, DEMCMC_PERSIST="0")                       # (a population this small would otherwise run in the persistent kernel)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), str(n_iter), "--child"], capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    print(name, out[name])
a, b, c3 = out.values()
assert a["acc"] == b["acc"] and a["state"] == b["state"], "the default configuration is not reproducible"
assert a["launches"] != c3["launches"], "the switches did not take"
same = a["acc"] == c3["acc"] and a["ids"] == c3["ids"] and a["state"] == c3["state"]
print("default twice: bit-identical; against the kernels they replace: accept decisions, ids and states", "identical" if same else "DIFFERENT")
assert same
