"""Helper of tests/test_multigpu.py: runs the snippet in RB200_GLOO_CODE under torchrun (gloo, CPU)."""
import os
exec(os.environ["RB200_GLOO_CODE"])
