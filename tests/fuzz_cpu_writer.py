#!/usr/bin/env python
"""The randomised writer / statistics tests of tests/test_oracle_vs_reference.py over other seeds than the committed ones (not collected
by pytest; needs oracle/_ref).  test_print_block_on_random_records differs from the reference only where a random block begins at
position 1 or 2 (the stale-window behaviour DESIGN.md section 4 describes).
   usage: python tests/fuzz_cpu_writer.py <test name> [first_seed] [n_seeds]"""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bindings import Oracle, Reference
import tests.test_oracle_vs_reference as T
oracle, ref = Oracle(), Reference(calc_threads=2)
which, first, count = sys.argv[1], (int(sys.argv[2]) if len(sys.argv) > 2 else 1000), (int(sys.argv[3]) if len(sys.argv) > 3 else 50)
fn = getattr(T, which)
bad = []
for seed in range(first, first + count):
    try:
        fn(oracle, ref, seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:80]))
print(which, "seeds %d..%d" % (first, first + count - 1), "failures:", bad[:5], len(bad))
