"""GPU parity for the popularity stage (ottocov_count_popularity, csrc/popularity.cu) against the CPU
restatements of model/count_popularity.py:53-85: integer counts and ranks, bit-exact."""
import json
import os

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq
import pytest
import torch

from oracle import popularity_oracle as po
from otto_recommender_b200 import OttocovError
from otto_recommender_b200.synth import SynthSpec, generate_numpy
from test_popularity_cpu import GOLD, _as_rows, random_pop_events

pytestmark = pytest.mark.gpu


def test_golden_popularity_gpu(engine):
    g = json.load(open(GOLD))
    ev = np.array(g["events"])
    r = engine.count_popularity(ev[:, 0], ev[:, 1], ev[:, 2], ev[:, 3], ts_recent=po.ts_recent_of(ev[:, 2]),
                                keep_top_k=g["expected"]["keep_top_k"])
    assert _as_rows(r) == g["expected"]["rows"]
    assert all(r[c].dtype == np.int16 for c in po.RANK_COLUMNS)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("k", [1, 5, 20, 2000])
def test_popularity_vs_oracle(engine, seed, k):
    c, a, t, y = random_pop_events(seed)
    want = po.popularity_ranks_frame(c, a, t, y, keep_top_k=k)
    got = engine.count_popularity(c, a, t, y, ts_recent=po.ts_recent_of(t), keep_top_k=k)
    assert _as_rows(got) == _as_rows(want)


def test_popularity_edges(engine):
    z = np.zeros(0, np.int32)
    r = engine.count_popularity(z, z, z, z.astype(np.int8), ts_recent=0, keep_top_k=20)
    assert len(r["aid"]) == 0
    # one event; all in cluster -1; every event recent / none recent
    for rec in (-2 ** 31, 2 ** 31 - 1):
        c = np.full(50, -1); a = np.arange(50) % 7; t = np.arange(50); y = np.arange(50) % 3
        want = po.popularity_ranks_frame(c, a, t, y, keep_top_k=3, ts_recent=rec)
        got = engine.count_popularity(c, a, t, y, ts_recent=rec, keep_top_k=3)
        assert _as_rows(got) == _as_rows(want)
    # more than 999 aids in a cluster: ranks clip at 999, ties by aid
    a = np.arange(3000); c = np.zeros(3000); t = np.zeros(3000); y = np.zeros(3000)
    got = engine.count_popularity(c, a, t, y, ts_recent=-5, keep_top_k=10 ** 6)
    assert got["rank_clicks"].tolist() == np.minimum(np.arange(1, 3001), 999).tolist()
    # device-resident inputs
    c, a, t, y = random_pop_events(11)
    dev = [torch.from_numpy(x).cuda() for x in (c, a, t, y)]
    got = engine.count_popularity(*dev, ts_recent=po.ts_recent_of(t), keep_top_k=5)
    assert _as_rows(got) == _as_rows(po.popularity_ranks_frame(c, a, t, y, keep_top_k=5))
    # schema violations are loud
    bad = y.copy(); bad[0] = 3
    with pytest.raises(OttocovError) as e:
        engine.count_popularity(c, a, t, bad, ts_recent=0)
    assert e.value.code == -3
    with pytest.raises(OttocovError):
        engine.count_popularity(np.full_like(c, -2), a, t, y, ts_recent=0)


def test_popularity_otto_shape(engine):
    """150 k synthetic OTTO-shaped sessions, general popularity (one cluster) and 50 pseudo-clusters."""
    d = generate_numpy(SynthSpec(n_sessions=150_000, seed=3))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    rng = np.random.default_rng(0)
    cl_of_session = rng.integers(-1, 50, int(s.max()) + 1).astype(np.int32)
    for cl in (np.zeros(len(s), np.int32), cl_of_session[s]):
        want = po.popularity_ranks_frame(cl, a, t, y, keep_top_k=20)
        got = engine.count_popularity(cl, a, t, y, ts_recent=po.ts_recent_of(t), keep_top_k=20)
        for k in want:
            assert np.array_equal(got[k], want[k]), k


def test_popularity_dropin_cli(engine, tmp_path, monkeypatch):
    """python -m model.count_popularity on a synthetic data directory == the restatement, file for file."""
    from otto_recommender_b200 import count_popularity as cp
    d = generate_numpy(SynthSpec(n_sessions=20_000, seed=9))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    root = tmp_path / "data"
    cut = len(s) // 2
    for sub, sl in (("train_sessions", slice(0, cut)), ("test_sessions", slice(cut, None))):
        os.makedirs(root / "x-parquet" / sub)
        pq.write_table(pa.table({"session": s[sl], "aid": a[sl], "ts": t[sl], "type": y[sl]}),
                       root / "x-parquet" / sub / "0000000_0100000.parquet")
    rng = np.random.default_rng(1)
    ses = np.unique(s)
    have = ses[rng.random(len(ses)) < 0.9]                       # 10 % of the sessions have no cluster
    os.makedirs(root / "x-sessions-clusters")
    pq.write_table(pa.table({"session": have.astype(np.int32), "cluster": rng.integers(0, 50, len(have)).astype(np.int16)}),
                   root / "x-sessions-clusters" / "sessions-clusters-50.parquet")
    cp.run(str(root / "x-parquet"), str(root / "x-sessions-clusters"), str(root / "x-counts-popularity"), keep_top_k=20,
           engine=engine)
    cl50 = cp.join_clusters(s, have.astype(np.int32), pq.read_table(root / "x-sessions-clusters" / "sessions-clusters-50.parquet")["cluster"].to_numpy().astype(np.int32))
    for n, cl in ((1, np.zeros(len(s), np.int32)), (50, cl50)):
        df = pd.read_parquet(root / "x-counts-popularity" / f"aid_clusters_{n}_count_ranks.parquet")
        want = po.popularity_ranks_frame(cl, a, t, y, keep_top_k=20)
        assert list(df.columns) == ["aid", f"cl{n}"] + [f"{c}_cl{n}" for c in po.RANK_COLUMNS]
        assert np.array_equal(df["aid"].to_numpy(), want["aid"]) and np.array_equal(df[f"cl{n}"].to_numpy(), want["cluster"])
        for c in po.RANK_COLUMNS:
            assert df[f"{c}_cl{n}"].dtype == np.int16 and np.array_equal(df[f"{c}_cl{n}"].to_numpy(), want[c])
    sc = pd.read_parquet(root / "x-counts-popularity" / "sessions_clusters.parquet")
    assert np.array_equal(sc["session"].to_numpy(), ses) and list(sc.columns) == ["session", "cl1", "cl50"]
    first = np.unique(s, return_index=True)[1]
    assert np.array_equal(sc["cl50"].to_numpy(), cl50[first]) and bool((sc["cl1"] == 0).all())
