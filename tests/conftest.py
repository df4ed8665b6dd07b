import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Say whether the unmodified compiled reference (oracle/_ref) took part in this run: its parity tests skip
    without it (GDR_REQUIRE_REF=1 turns those skips into failures)."""
    from oracle import ref_api

    skipped = [r for r in terminalreporter.stats.get("skipped", []) if "oracle/_ref" in str(getattr(r, "longrepr", ""))]
    state = "present" if ref_api.available() else "ABSENT"
    terminalreporter.write_line(f"reference parity: oracle/_ref {state}; {len(skipped)} reference-parity test(s) skipped")
