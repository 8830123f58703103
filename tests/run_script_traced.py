"""Test harness (not product): run a reference script UNMODIFIED (`runpy`, __main__) with this repo's `compat/` packages first on
sys.path and the module tree in trace mode (plain torch ops on CPU, as the dependency tracer uses) — exercises the scripts'
import surface, checkpoint I/O and pruning plumbing in the CPU-only build container.  The kernels are covered by the -m gpu tests."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "diff-pruning_b200", "compat"), ROOT]

if __name__ == "__main__":
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(2, os.path.dirname(os.path.abspath(script)))   # `import utils` of the scripts — AFTER compat/, which shadows the vendored diffusers
    import diff_pruning_b200 as dp
    with dp.trace_mode():
        runpy.run_path(script, run_name="__main__")
