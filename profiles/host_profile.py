"""Developer helper (run under gpurun): cProfile of the host side of threshold.from_cv on the config-5 tables."""
import cProfile, os, pstats, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
from biscuit_b200 import threshold as T
from oracle import synth
dfs = synth.cv_tables(k=10, n_slides=100, tiles_per_slide=2000, seed0=0)
pats = {}
for d in dfs:
    pats.update(synth.patients_map(d))
T.from_cv([d.copy() for d in dfs], patients=pats)
cp = [d.copy() for d in dfs]
pr = cProfile.Profile()
pr.enable()
T.from_cv(cp, patients=pats)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(32)
