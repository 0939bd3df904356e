import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_terminal_summary(terminalreporter):
    """Tally of the tolerance policy of tests/_parity.py: how many cost assertions held at the plain 1e-5 / 1e-10 rule and
    how many needed the float32 noise allowance; how many LM iteration-count comparisons were identical and how many used
    the borderline-zeta rule (shown with `pytest -rA` or any run that reaches the summary)."""
    par = sys.modules.get("_parity")
    if par is None:
        return
    a = par.AUDIT
    tr = terminalreporter
    tr.write_sep("-", "parity tolerance audit (tests/_parity.py)")
    tr.write_line("cost assertions: %d at the plain rule, %d via NOISE_FACTOR" % (a["cost_plain"], a["cost_via_noise_factor"]))
    tr.write_line("LM PCG-count comparisons: %d identical, %d via ZETA_MARGIN" % (a["lm_counts_identical"], a["lm_counts_via_zeta_margin"]))
    for t in a["hatch_users"]:
        tr.write_line("  allowance used by: " + t)
