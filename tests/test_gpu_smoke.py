"""The driver's round-end smoke() must stay green: run it as part of the GPU suite (it asserts that the named C-ABI
entry points were launched by a train step and checks the encoder features against the oracle)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.gpu
def test_graft_entry_smoke_runs():
    import __graft_entry__ as g
    g.smoke()
