"""CPU suite for the popularity stage: the two restatements of
model/count_popularity.py:53-85 against the hand-verified vector and against each other, plus the host-side
session -> cluster join of the drop-in."""
import json
import os

import numpy as np
import pytest

from oracle import popularity_oracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pop_events.json")


def _as_rows(r):
    return np.stack([r["cluster"], r["aid"]] + [r[c] for c in po.RANK_COLUMNS], 1).tolist()


def random_pop_events(seed, n=5000, n_aids=80, n_clusters=7):
    rng = np.random.default_rng(seed)
    aid = np.minimum((rng.pareto(1.2, n) * 4).astype(np.int64), n_aids - 1).astype(np.int32)     # skewed: many ties at 1..3
    cluster = rng.integers(-1, n_clusters, n).astype(np.int32)
    ts = (1_660_000_000 + rng.integers(0, 28 * 86400, n)).astype(np.int32)
    type_ = rng.choice([0, 0, 0, 0, 1, 2], n).astype(np.int8)
    return cluster, aid, ts, type_


@pytest.mark.parametrize("fn", [po.popularity_ranks_frame, po.popularity_ranks_loops])
def test_golden_popularity(fn):
    g = json.load(open(GOLD))
    ev = np.array(g["events"])
    r = fn(ev[:, 0], ev[:, 1], ev[:, 2], ev[:, 3], keep_top_k=g["expected"]["keep_top_k"])
    assert _as_rows(r) == g["expected"]["rows"]
    assert all(r[c].dtype == np.int16 for c in po.RANK_COLUMNS)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("k", [1, 5, 20, 2000])
def test_restatements_agree(seed, k):
    c, a, t, y = random_pop_events(seed)
    r1 = po.popularity_ranks_frame(c, a, t, y, keep_top_k=k)
    r2 = po.popularity_ranks_loops(c, a, t, y, keep_top_k=k)
    assert _as_rows(r1) == _as_rows(r2)
    if k == 2000:        # nothing filtered: ranks clip at 999 and every (cluster, aid) group is present
        assert len(r1["aid"]) == len({(int(x), int(z)) for x, z in zip(c, a)})
        assert max(int(r1[col].max()) for col in po.RANK_COLUMNS) <= 999


def test_rank_properties():
    c, a, t, y = random_pop_events(7, n=20000, n_aids=3000, n_clusters=2)
    r = po.popularity_ranks_frame(c, a, t, y, keep_top_k=10 ** 6)
    for cl in np.unique(r["cluster"]):
        m = r["cluster"] == cl
        for col in po.RANK_COLUMNS:
            ranks = np.sort(r[col][m].astype(np.int64))
            want = np.minimum(np.arange(1, m.sum() + 1), 999)
            assert np.array_equal(ranks, want)            # an ordinal rank: a permutation of 1..n, clipped


def test_join_clusters_host():
    from otto_recommender_b200.count_popularity import join_clusters
    session = np.array([5, 5, 9, 2, 7, 100], np.int32)
    got = join_clusters(session, np.array([9, 2, 5], np.int32), np.array([3, 0, 41], np.int32))
    assert got.tolist() == [41, 41, 3, 0, -1, -1]
    assert join_clusters(session, np.zeros(0, np.int32), np.zeros(0, np.int32)).tolist() == [-1] * 6
