"""Timeline of the persistent chunk kernel from in-kernel %globaltimer stamps (no events, no profiler).
Usage on the GPU box:  [DEMCMC_LANES=2] python scripts/pk_timeline.py [out.csv] [n_iter]"""
import os, sys, csv
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pk_timeline.csv"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 80
os.environ["DEMCMC_PK_TIMELINE"] = "8000"
os.environ["DEMCMC_PK_TIMELINE_FILE"] = out
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import demcmc_b200 as D
D._ffi.use_library(D._ffi.DEFAULT_LIB)
x, prior, lo, hi, theta0 = bench.workload(4)
h = D.Handle(4, 256, 51, lo, hi, burnin=0, theta_snooker=0.1, seed=20261017)
h.set_model("mvnormal", prior, x=x)
h.set_state(theta0)
h.run(20)
c0 = h.counters()
h.run(n_iter)
c = h.counters()
print("lanes", os.environ.get("DEMCMC_LANES", "1"), "updates/s %.0f" % (1024 * n_iter / (c["device_ms"] * 1e-3)), "ms/iter %.4f" % (c["device_ms"] / n_iter),
      "persistent chunks", c["persistent_chunks"] - c0["persistent_chunks"], "launches", c["kernel_launches"] - c0["kernel_launches"])
h.close()
R = [{k: float(v) for k, v in r.items()} for r in csv.DictReader(open(out))]
chunks = sorted(set(r["chunk"] for r in R))
R = [r for r in R if r["chunk"] >= chunks[min(3, len(chunks) - 1)]]
def mean(f, rows=R): return float(np.mean([f(r) for r in rows]))
print("levels", len(R), "mean n %.1f" % mean(lambda r: r["n"]))
print("propose: first warp at level -> dependency ready %.2f us; ready -> last proposal staged %.2f us" % (
    mean(lambda r: r["propose_dep_ready"] - r["propose_first_start"]), mean(lambda r: r["propose_last_end"] - r["propose_dep_ready"])))
print("last proposal staged -> first DMMA warp released %.2f us, -> last released %.2f us" % (
    mean(lambda r: r["xdot_first_ready"] - r["propose_last_end"]), mean(lambda r: r["xdot_last_ready"] - r["propose_last_end"])))
print("DMMA: first released -> last item end %.2f us" % mean(lambda r: r["xdot_last_end"] - r["xdot_first_ready"]))
print("last item end -> last accept end %.2f us (first accept released %.2f us before the last item end)" % (
    mean(lambda r: r["accept_last_end"] - r["xdot_last_end"]), mean(lambda r: r["xdot_last_end"] - r["accept_first_ready"])))
per = []
for ch in sorted(set(r["chunk"] for r in R)):
    rows = [r for r in R if r["chunk"] == ch]
    for a, b in zip(rows[:-1], rows[1:]):
        per.append(b["xdot_last_end"] - a["xdot_last_end"])
print("level period (last item end to last item end) %.2f us" % float(np.mean(per)))
print("per DMMA warp and item: wait for proposals %.2f us, item (B fragments + DMMA loop + flush) %.2f us, of which B fragments + FIRST observation tile %.2f us" % (
    mean(lambda r: r["wait_us_per_item"]), mean(lambda r: r["work_us_per_item"]), mean(lambda r: r["arrive_us_per_item"])))
print("per proposal: prologue + pending accepts + dependency wait %.2f us, body %.2f us, staging %.2f us, arrive %.2f us; per accept: body %.2f us, arrive %.2f us" % (
    mean(lambda r: r["prop_pre_us"]), mean(lambda r: r["prop_body_us"]), mean(lambda r: r["prop_stage_us"]), mean(lambda r: r["prop_arrive_us"]),
    mean(lambda r: r["acc_body_us"]), mean(lambda r: r["acc_arrive_us"])))
print("ideal DMMA time of a level (13 k-steps, octet padding) %.2f us" % mean(lambda r: np.ceil(r["n"] / 8) * 1563 * 8 * 13 * 16 / (148 * 4) / 1965.0))
