"""Synthetic workload generators (checkasm-style inputs, synthetic frames) shared
by the tests and bench.py.  Host-side numpy only; no DSP arithmetic lives here."""
