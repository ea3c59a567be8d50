"""CPU restatement of the reference's popularity counting.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/model/count_popularity.py:53-85 for one clustering, twice and independently:

  * ``popularity_ranks_frame``  dataframe-shaped (pandas group-by / sort / cumcount), following the reference
                                line by line: time_max / ts_7d (:53-54), the six conditional sums (:61-70), the
                                ordinal rank per cluster clipped to 999 as Int16 (:73-75), the min-rank filter (:81);
  * ``popularity_ranks_loops``  plain Python dictionaries and sorts, sharing no code with the first.

PARITY UNPINNED by the reference (it runs inside polars, ships no tests or fixtures for this stage).  Pinned
instead by the hand-verified vector tests/golden/pop_events.json and by the agreement of the two restatements.
The reference's ``rank('ordinal')`` breaks ties by the row order its hash group-by happened to leave; the
canonical rule used by both restatements and by the engine is count descending, then aid ascending.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg (--workload popularity) may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

RANK_COLUMNS = ("rank_clicks", "rank_carts", "rank_orders", "rank_clicks_7d", "rank_carts_7d", "rank_orders_7d")
SEVEN_DAYS = 7 * 24 * 60 * 60


def ts_recent_of(ts) -> int:
    """count_popularity.py:53-54: time_max - 7 days; an event is 'recent' iff ts > that."""
    return int(np.max(ts)) - SEVEN_DAYS


def popularity_ranks_frame(cluster, aid, ts, type_, keep_top_k: int = 20, ts_recent: Optional[int] = None
                           ) -> Dict[str, np.ndarray]:
    import pandas as pd
    if len(aid) == 0:
        return {"aid": np.zeros(0, np.int32), "cluster": np.zeros(0, np.int32),
                **{c: np.zeros(0, np.int16) for c in RANK_COLUMNS}}
    if ts_recent is None:
        ts_recent = ts_recent_of(ts)
    df = pd.DataFrame({"cl": np.asarray(cluster, np.int64), "aid": np.asarray(aid, np.int64),
                       "ts": np.asarray(ts, np.int64), "type": np.asarray(type_, np.int64)})
    rec = df["ts"] > ts_recent
    for t, nm in enumerate(("clicks", "carts", "orders")):
        df[f"n_{nm}"] = (df["type"] == t).astype(np.int64)
        df[f"n_{nm}_7d"] = ((df["type"] == t) & rec).astype(np.int64)
    ncols = ["n_clicks", "n_carts", "n_orders", "n_clicks_7d", "n_carts_7d", "n_orders_7d"]
    agg = df.groupby(["cl", "aid"], sort=True)[ncols].sum().reset_index()
    for col in ncols:
        o = agg.sort_values(["cl", col, "aid"], ascending=[True, False, True], kind="stable")
        r = o.groupby("cl", sort=False).cumcount() + 1
        agg.loc[o.index, col.replace("n_", "rank_")] = np.minimum(r.to_numpy(), 999)
    rcols = [c.replace("n_", "rank_") for c in ncols]
    keep = agg[rcols].min(axis=1) <= keep_top_k
    agg = agg[keep].sort_values(["cl", "aid"])
    out = {"aid": agg["aid"].to_numpy(np.int32), "cluster": agg["cl"].to_numpy(np.int32)}
    for c in rcols:
        out[c] = agg[c].to_numpy().astype(np.int16)
    return out


def popularity_ranks_loops(cluster, aid, ts, type_, keep_top_k: int = 20, ts_recent: Optional[int] = None
                           ) -> Dict[str, np.ndarray]:
    if ts_recent is None and len(aid):
        ts_recent = ts_recent_of(ts)
    counts: Dict[tuple, list] = {}
    for c, a, t, y in zip(cluster, aid, ts, type_):
        row = counts.setdefault((int(c), int(a)), [0] * 6)
        row[int(y)] += 1
        if int(t) > ts_recent:
            row[3 + int(y)] += 1
    ranks: Dict[tuple, list] = {k: [0] * 6 for k in counts}
    clusters = sorted({k[0] for k in counts})
    for cl in clusters:
        members = [k for k in counts if k[0] == cl]
        for col in range(6):
            members.sort(key=lambda k: (-counts[k][col], k[1]))
            for pos, k in enumerate(members):
                ranks[k][col] = min(pos + 1, 999)
    kept = sorted(k for k in counts if min(ranks[k]) <= keep_top_k)
    out = {"aid": np.array([k[1] for k in kept], np.int32), "cluster": np.array([k[0] for k in kept], np.int32)}
    for i, name in enumerate(RANK_COLUMNS):
        out[name] = np.array([ranks[k][i] for k in kept], np.int16)
    return out
