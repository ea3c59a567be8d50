"""Drop-in for the reference stage ``model/count_popularity.py`` on a B200.

Same command line (``python -m model.count_popularity --data_split_alias A --keep_top_k K``; reference :14-20),
same inputs (``{DIR_DATA}/{alias}-parquet/{train,test}_sessions/*.parquet`` and, for every n > 1 in
``N_CLUSTERS_TO_JOIN``, ``{alias}-sessions-clusters/sessions-clusters-{n}.parquet`` written by the reference's
kmeans stage, :26,46) and same outputs (``{alias}-counts-popularity/aid_clusters_{n}_count_ranks.parquet`` with
``aid, cl{n}, rank_{clicks,carts,orders}[_7d]_cl{n}`` (ranks Int16) and ``sessions_clusters.parquet``, :85-89).

The group-by, the six ordinal ranks per cluster and the top-k filter run on the GPU (``ottocov_count_popularity``,
csrc/popularity.cu); the session -> cluster left join (:47-51) is a sorted lookup on the host.  Deviations from the
reference: ``--keep_top_k`` is parsed as int (the reference compares a str when the flag is given); ties of the
ordinal rank are broken by aid ascending (the reference leaves them to its hash group-by); output rows are ordered
by (cluster, aid).  There is no CPU fallback.
"""
from __future__ import annotations

import argparse
import glob
import json
import logging
import os
from typing import Dict, Optional, Sequence

import numpy as np
import pandas as pd
import pyarrow.parquet as pq

from .config import DEFAULT_CONFIG
from .engine import Engine

log = logging.getLogger("count_popularity")

KEEP_TOP_K = 20                      # reference config.py:31
N_CLUSTERS_TO_JOIN = (1, 50)         # reference config.py:196
SEVEN_DAYS = 7 * 24 * 60 * 60


def join_clusters(session: np.ndarray, cl_sessions: np.ndarray, cl_values: np.ndarray) -> np.ndarray:
    """df_sessions.join(df_clusters, on='session', how='left').fill_null(-1) (reference :47-51), per event."""
    order = np.argsort(cl_sessions, kind="stable")
    ks, vs = cl_sessions[order], cl_values[order]
    pos = np.searchsorted(ks, session)
    pos_c = np.minimum(pos, max(len(ks) - 1, 0))
    hit = (pos < len(ks)) & (ks[pos_c] == session) if len(ks) else np.zeros(len(session), bool)
    return np.where(hit, vs[pos_c] if len(ks) else 0, -1).astype(np.int32)


def count_popularity(engine: Engine, session, aid, ts, type_, clusters: Dict[int, np.ndarray],
                     keep_top_k: int = KEEP_TOP_K) -> Dict[int, pd.DataFrame]:
    """clusters: n_clusters -> per-EVENT cluster id (int, -1 = none).  Returns n_clusters -> the frame the
    reference writes to aid_clusters_{n}_count_ranks.parquet (:76-85)."""
    ts = np.asarray(ts)
    ts_7d = int(ts.max()) - SEVEN_DAYS if len(ts) else 0                      # :53-54
    out = {}
    for n_clusters, cl in clusters.items():
        r = engine.count_popularity(cl, aid, ts, type_, ts_recent=ts_7d, keep_top_k=keep_top_k)
        cl_dtype = np.int8 if n_clusters == 1 else np.int16                   # :41 (lit Int8), kmeans_sessions.py:169
        cols = {"aid": r["aid"], f"cl{n_clusters}": r["cluster"].astype(cl_dtype)}
        for name in Engine.POPULARITY_RANK_COLUMNS:
            cols[f"{name}_cl{n_clusters}"] = r[name]
        out[n_clusters] = pd.DataFrame(cols)
    return out


def run(dir_sessions: str, dir_sessions_clusters: str, dir_out: str, keep_top_k: int = KEEP_TOP_K,
        n_clusters_to_join: Sequence[int] = N_CLUSTERS_TO_JOIN, engine: Optional[Engine] = None) -> None:
    os.makedirs(dir_out, exist_ok=True)
    files = sorted(glob.glob(f"{dir_sessions}/train_sessions/*.parquet") + glob.glob(f"{dir_sessions}/test_sessions/*.parquet"))
    if not files:
        raise FileNotFoundError(f"no session parquet files under {dir_sessions}")
    parts = [pq.read_table(f, columns=["session", "aid", "ts", "type"]) for f in files]           # :30-34
    session = np.concatenate([p["session"].to_numpy() for p in parts]).astype(np.int32)
    aid = np.concatenate([p["aid"].to_numpy() for p in parts]).astype(np.int32)
    ts = np.concatenate([p["ts"].to_numpy() for p in parts]).astype(np.int32)
    type_ = np.concatenate([p["type"].to_numpy() for p in parts]).astype(np.int8)
    log.debug(f"Loaded {len(session):,} events")
    clusters: Dict[int, np.ndarray] = {}
    for n in n_clusters_to_join:                                                                  # :39-49
        if n == 1:
            clusters[1] = np.zeros(len(session), np.int32)
        else:
            t = pq.read_table(f"{dir_sessions_clusters}/sessions-clusters-{n}.parquet")
            clusters[n] = join_clusters(session, t["session"].to_numpy().astype(np.int32), t["cluster"].to_numpy().astype(np.int32))
    own = engine is None
    eng = engine or Engine()
    try:
        frames = count_popularity(eng, session, aid, ts, type_, clusters, keep_top_k)
    finally:
        if own:
            eng.close()
    for n, df in frames.items():
        df.to_parquet(f"{dir_out}/aid_clusters_{n}_count_ranks.parquet", index=False)             # :85
        log.debug(f"Saved clusters with top {keep_top_k} aids by each type/horizon, n_clusters={n}: {len(df):,} rows")
    # session -> clusters, unique, sorted by session (:87-89)
    first = np.unique(session, return_index=True)
    cols = {"session": first[0]}
    for n in n_clusters_to_join:
        cols[f"cl{n}"] = clusters[n][first[1]].astype(np.int8 if n == 1 else np.int16)
    pd.DataFrame(cols).to_parquet(f"{dir_out}/sessions_clusters.parquet", index=False)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--data_split_alias", default="train-test")
    parser.add_argument("--keep_top_k", type=int, default=KEEP_TOP_K)
    args = parser.parse_args(argv)
    log.info("Running count_popularity with parameters: \n" + json.dumps(vars(args), indent=2))
    d = DEFAULT_CONFIG.DIR_DATA
    run(f"{d}/{args.data_split_alias}-parquet", f"{d}/{args.data_split_alias}-sessions-clusters",
        f"{d}/{args.data_split_alias}-counts-popularity", keep_top_k=args.keep_top_k)


if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s - %(name)s - %(levelname)s - %(message)s", level=logging.DEBUG)
    main()
