import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def rel_close(a, b, rtol=1e-3, eps_scale=1e-6):
    """Parity metric of SURVEY.md section 8c: |a-b| <= rtol*max(|a|,|b|,eps), eps = eps_scale*mean|ref|... returns the
    fraction of entries that pass and the worst relative error."""
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    eps = max(eps_scale * float(np.abs(b).mean()), 1e-30)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), eps)
    err = np.abs(a - b) / den
    return float((err <= rtol).mean()), float(err.max())


@pytest.fixture(scope="session")
def relclose():
    return rel_close
