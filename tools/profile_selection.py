#!/usr/bin/env python
"""Developer tool (GPU box): cProfile of tools/selection_bench.py (host side of the lock-step backward selection)."""
import cProfile, pstats, runpy, sys
args = sys.argv[1:] or ["256", "50000", "32", "8", "0.1"]
sys.argv = ["tools/selection_bench.py"] + args
cProfile.run("runpy.run_path(sys.argv[0], run_name='__main__')", "/tmp/prof.out")
p = pstats.Stats("/tmp/prof.out")
p.sort_stats("cumulative").print_stats(45)
