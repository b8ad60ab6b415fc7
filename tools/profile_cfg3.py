#!/usr/bin/env python
"""Developer tool (GPU box): cProfile of bench.selection_rate (cfg3, 256 replicates in lock step): where the wall time
outside the solver goes."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cProfile.run("out = bench.selection_rate(0, n_rep=n_rep)", "/tmp/prof.out")
print(out)
p = pstats.Stats("/tmp/prof.out")
p.sort_stats("tottime").print_stats(35)
p.sort_stats("cumulative").print_stats(45)
