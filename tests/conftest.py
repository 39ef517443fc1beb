import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


import os


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite (-m "not gpu") is dominated by the kernel interpreter: spread it over the cores with pytest-xdist when
    it is installed and the caller did not choose a worker count.  The GPU suite stays in one process."""
    try:
        import xdist  # noqa: F401
    except Exception:
        return None
    expr = (getattr(config.option, "markexpr", "") or "").replace(" ", "")
    if expr == "notgpu" and getattr(config.option, "numprocesses", None) is None and "PYTEST_XDIST_WORKER" not in os.environ:
        config.option.numprocesses = max(1, min(6, (os.cpu_count() or 2) - 1))
    return None
