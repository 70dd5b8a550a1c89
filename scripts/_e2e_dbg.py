import sys, os, cProfile, pstats, io
sys.path.insert(0, os.getcwd())
import bench, demcmc_b200 as D
orig = D.sample
def prof_sample(*a, **k):
    pr = cProfile.Profile(); pr.enable()
    r = orig(*a, **k)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:3000], file=sys.stderr)
    return r
D.sample = prof_sample
sys.argv = ["bench.py", "--no-cpu"]
bench.main()
