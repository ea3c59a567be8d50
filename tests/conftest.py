import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libottocov.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from otto_recommender_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def engine(built_lib):
    """One Engine on cuda:0 for the whole GPU test session (creation fails loudly without a GPU)."""
    from otto_recommender_b200 import Engine
    eng = Engine(device=0)
    yield eng
    eng.close()


def small_events(seed: int, n_sessions: int = 300, n_aids: int = 50, max_len: int = 40, span: int = 200_000,
                 shuffle: bool = True, dup_frac: float = 0.05):
    """Dense little event sets: few aids (many repeated pairs), timestamps straddling the 12 h / 24 h
    window edges, duplicates, optional shuffling.  Returns int32/int8 numpy columns."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len + 1, n_sessions)
    sess = np.repeat(np.arange(n_sessions) * 7 + 3, lens)        # non-contiguous ids
    n = len(sess)
    start = np.repeat(rng.integers(1_660_000_000, 1_660_000_000 + 5 * 86400, n_sessions), lens)
    gaps = rng.choice([0, 1, 30, 600, 43200, 43201, 86400, 86401, 100_000], n) * rng.integers(0, 2, n)
    ts = start + rng.integers(0, span, n) // 4 + gaps
    aid = rng.integers(0, n_aids, n)
    typ = rng.choice([0, 0, 0, 0, 1, 1, 2], n)
    nd = int(n * dup_frac)
    if nd:
        pick = rng.integers(0, n, nd)
        sess, ts, aid, typ = (np.concatenate([x, x[pick]]) for x in (sess, ts, aid, typ))
    order = rng.permutation(len(sess)) if shuffle else np.lexsort((ts, sess))
    return (sess[order].astype(np.int32), aid[order].astype(np.int32), ts[order].astype(np.int32),
            typ[order].astype(np.int8))
