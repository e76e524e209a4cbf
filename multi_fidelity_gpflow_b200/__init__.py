"""B200-native Kennedy-O'Hagan linear multi-fidelity GP core (CUDA sm_100a behind a C-ABI).

Host-side mirror of the reference's interface (mfgpflow/linear.py, singlebin_svgp.py,
linear_svgp.py); the arithmetic runs in libmfgp.so.  No CPU fallback.
"""
__version__ = "0.1.0"
