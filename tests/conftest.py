import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def readme_data(oracle):
    """README.md:112-115 — set.seed(78845); X <- rnorm(n); y <- 3 + 2 X + rnorm(n, sd=4),
    regenerated with the restated R RNG (oracle/r_rng.c)."""
    R = oracle.RRng
    R.set_seed(78845)
    n = 1000
    X = R.rnorm(n)
    y = 3.0 + 2.0 * X + R.rnorm(n, 0.0, 4.0)
    return dict(n=n, X=X, y=y, sd_y=R.sd(y))


def pytest_terminal_summary(terminalreporter):
    """Worst norm-wise and element-wise errors the parity helper saw (tests/gpu_util.py)."""
    try:
        from gpu_util import PARITY_LOG
    except Exception:
        return
    if PARITY_LOG:
        worst_n = max(PARITY_LOG, key=lambda r: r[2])
        worst_e = max(PARITY_LOG, key=lambda r: r[3])
        terminalreporter.write_line(
            f"parity: {len(PARITY_LOG)} comparisons; worst norm-wise {worst_n[2]:.2e} ({worst_n[0]} {worst_n[1]}); "
            f"worst element-wise {worst_e[3]:.2e} ({worst_e[0]} {worst_e[1]})")
