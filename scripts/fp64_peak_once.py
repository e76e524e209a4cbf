"""Launch the two FP64-pipe microbenchmarks once (for an ncu pass that reads the pipe-utilisation metrics at the measured
peak: what ncu's fp64 percentages read when the pipe is saturated)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_fidelity_gpflow_b200 import _lib
h = _lib.Handle(0)
print("DFMA", h.fp64_peak(0, 4000) / 1e12, "DMMA", h.fp64_peak(1, 4000) / 1e12, "MIX", h.fp64_peak(2, 4000) / 1e12)
