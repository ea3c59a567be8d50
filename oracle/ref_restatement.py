"""CPU restatement of the reference's co-event counting path.  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: it restates, step for step, what
``/root/reference/model/count_co_events.py`` and the per-aid top-N in
``/root/reference/model/retrieve.py`` compute, with pyarrow (Acero hash join +
hash group-by, multi-threaded) standing in for polars.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product package never does.

PARITY UNPINNED: the reference executes inside the third-party ``polars`` engine
(unpinned in requirements.txt:19, not vendored, not installable here) and ships no
tests, fixtures or golden vectors for this path.  This oracle is pinned instead by
(1) the hand-verified known-answer vectors of SURVEY.md App. B.1
(``tests/golden/b1_events.json``), (2) an independent plain-C per-session double
loop (``oracle/cov_oracle.c``) and (3) the algebraic invariants of App. B.2.

Reference semantics restated (file:line into /root/reference):
  * event dedup ``df.unique()``                       model/count_co_events.py:92
  * self join on session, 10 000 sessions per slice    model/count_co_events.py:17-19, 41-57
  * drop the identical-event rows                      model/count_co_events.py:23-27
  * time_to_next, +-24 h filter (inclusive)            model/count_co_events.py:30-36, config.py:41-42
  * five filters + group-by count                      model/count_co_events.py:60-77, config.py:43-49, 81-88
  * merge / thresholds / sort desc / head / Int32      model/count_co_events.py:103-181, config.py:52-64
  * per-aid ordinal rank desc, keep rank <= first_n    model/retrieve.py:41-51, config.py:90-96
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

# ---- constants: config.py:41-96 of the reference --------------------------------------------
MIN_TIME_TO_NEXT = -24 * 60 * 60
MAX_TIME_TO_NEXT = 24 * 60 * 60
MAP_MAX_TIME_TO_NEXT = {
    "click_to_click": 12 * 60 * 60,
    "click_to_cart_or_buy": MAX_TIME_TO_NEXT,
    "cart_to_cart": MAX_TIME_TO_NEXT,
    "cart_to_buy": MAX_TIME_TO_NEXT,
    "buy_to_buy": MAX_TIME_TO_NEXT,
}
OPTIM_ROWS_GROUPBY = 100_000_000
MAX_ROWS_GROUPBY = 300_000_000
MIN_COUNT_TO_SAVE = {
    "click_to_click": 10,
    "click_to_cart_or_buy": 5,
    "cart_to_cart": 2,
    "cart_to_buy": 2,
    "buy_to_buy": 2,
}
MIN_COUNT_IN_PART = {"click_to_click": 2, "click_to_cart_or_buy": 2}
MAX_PAIRS_TO_SAVE = 300_000_000
CO_EVENTS_TO_COUNT = list(MAP_MAX_TIME_TO_NEXT)
MAP_NAME_COUNT_TYPE = {
    "click_to_click": (0, [0]),
    "click_to_cart_or_buy": (0, [1, 2]),
    "cart_to_cart": (1, [1]),
    "cart_to_buy": (1, [2]),
    "buy_to_buy": (2, [2]),
}
RETRIEVAL_FIRST_N = {
    "click_to_click": 10,
    "click_to_cart_or_buy": 10,
    "cart_to_cart": 20,
    "cart_to_buy": 20,
    "buy_to_buy": 20,
}

EVENT_COLS = ["session", "aid", "ts", "type"]


def events_table(session, aid, ts, type_) -> pa.Table:
    """Columns with the dtypes etl/jsonl_to_parquet.py:23-29 writes."""
    return pa.table(
        {
            "session": pa.array(np.asarray(session), pa.int32()),
            "aid": pa.array(np.asarray(aid), pa.int32()),
            "ts": pa.array(np.asarray(ts), pa.int32()),
            "type": pa.array(np.asarray(type_), pa.int8()),
        }
    )


def unique_events(df: pa.Table) -> pa.Table:
    """``df.unique()`` over all four columns (count_co_events.py:92)."""
    return df.group_by(EVENT_COLS, use_threads=True).aggregate([])


def self_merge(df_part: pa.Table) -> pa.Table:
    """count_co_events.py:17-38 -- join, drop identical event, time filter."""
    m = df_part.join(df_part, keys="session", right_suffix="_next", join_type="inner")
    same = pc.and_(
        pc.and_(pc.equal(m["aid"], m["aid_next"]), pc.equal(m["ts"], m["ts_next"])),
        pc.equal(m["type"], m["type_next"]),
    )
    m = m.filter(pc.invert(same))
    # int32 - int32 stays int32 in polars; widen here so a pathological ts range cannot wrap
    dt = pc.subtract(pc.cast(m["ts_next"], pa.int64()), pc.cast(m["ts"], pa.int64()))
    m = m.append_column("time_to_next", dt)
    keep = pc.and_(
        pc.greater_equal(m["time_to_next"], MIN_TIME_TO_NEXT),
        pc.less_equal(m["time_to_next"], MAX_TIME_TO_NEXT),
    )
    return m.filter(keep)


def self_merge_big_df(df: pa.Table, n_sessions_in_part: int = 10_000) -> pa.Table:
    """count_co_events.py:41-57 -- slices of 10 000 distinct sessions (memory chunking only)."""
    sessions = pc.unique(df["session"])
    n_sessions = len(sessions)
    n_parts = math.ceil(n_sessions / n_sessions_in_part)
    merged = []
    for i_part in range(n_parts):
        lo = i_part * n_sessions_in_part
        hi = min(lo + n_sessions_in_part, n_sessions)
        part = df.filter(pc.is_in(df["session"], value_set=sessions.slice(lo, hi - lo)))
        merged.append(self_merge(part))
    if not merged:
        return self_merge(df)
    return pa.concat_tables(merged)


def count_co_events(df_merged: pa.Table) -> Dict[str, pa.Table]:
    """count_co_events.py:60-77 -- five (type, type_next, |dt|) filters + group-by count."""
    out = {}
    for name, (type_this, types_next) in MAP_NAME_COUNT_TYPE.items():
        mask = pc.and_(
            pc.and_(
                pc.equal(df_merged["type"], type_this),
                pc.is_in(df_merged["type_next"], value_set=pa.array(types_next, pa.int8())),
            ),
            pc.less_equal(pc.abs(df_merged["time_to_next"]), MAP_MAX_TIME_TO_NEXT[name]),
        )
        g = (
            df_merged.filter(mask)
            .group_by(["aid", "aid_next"], use_threads=True)
            .aggregate([("aid_next", "count")])
        )
        out[name] = pa.table(
            {
                "aid": g["aid"],
                "aid_next": g["aid_next"],
                "count": pc.cast(g["aid_next_count"], pa.uint32()),
            }
        )
    return out


def count_part(df: pa.Table, n_sessions_in_part: int = 10_000) -> Dict[str, pa.Table]:
    """Body of the per-file loop, count_co_events.py:91-94."""
    df = unique_events(df)
    return count_co_events(self_merge_big_df(df, n_sessions_in_part))


def merge_counts(
    name: str,
    tables: Iterable[pa.Table],
    *,
    exact: bool = False,
    min_count_to_save: Optional[int] = None,
    rows_trigger_min_in_part: int = 100_000_000,
    max_rows_groupby: int = MAX_ROWS_GROUPBY,
    optim_rows_groupby: int = OPTIM_ROWS_GROUPBY,
    max_pairs_to_save: int = MAX_PAIRS_TO_SAVE,
) -> pa.Table:
    """concat_files_w_stats (count_co_events.py:103-181) on in-memory tables.

    ``exact=True`` switches the two row-count-triggered lossy steps off (they never trigger
    for N <= 1e8 anyway).  Rows come back ordered by (count desc, aid asc, aid_next asc): the
    reference leaves the order among equal counts unspecified, this is the canonical choice.
    """
    df = pa.concat_tables([t.cast(pa.schema([("aid", pa.int32()), ("aid_next", pa.int32()),
                                             ("count", pa.int64())])) for t in tables])
    assert df.column_names == ["aid", "aid_next", "count"]          # :128
    n = df.num_rows
    if not exact and "click_to" in name and n > rows_trigger_min_in_part:   # :131-132 (100_000_000 there)
        df = df.filter(pc.greater_equal(df["count"], MIN_COUNT_IN_PART.get(name, 1)))
    if not exact and df.num_rows > max_rows_groupby:                 # :135-166
        n = df.num_rows
        rows_part = optim_rows_groupby
        n_parts = math.ceil(n / rows_part)
        max_rows_part = int(max_rows_groupby / n * rows_part)
        rows_part = math.ceil(n / n_parts)
        parts = []
        for i in range(n_parts):
            p = _sum_by_pair(df.slice(i * rows_part, rows_part))
            p = p.filter(pc.greater_equal(p["count"], MIN_COUNT_IN_PART.get(name, 1)))
            p = _sort_canonical(p).slice(0, max_rows_part)
            parts.append(p)
        df = pa.concat_tables(parts)
    df = _sum_by_pair(df)                                            # :168
    thr = MIN_COUNT_TO_SAVE.get(name, 1) if min_count_to_save is None else min_count_to_save
    df = df.filter(pc.greater_equal(df["count"], thr))              # :172
    df = _sort_canonical(df).slice(0, max_pairs_to_save)             # :173-174
    return pa.table({"aid": df["aid"], "aid_next": df["aid_next"],
                     "count": pc.cast(df["count"], pa.int32())})    # :175


def _sum_by_pair(df: pa.Table) -> pa.Table:
    g = df.group_by(["aid", "aid_next"], use_threads=True).aggregate([("count", "sum")])
    return pa.table({"aid": g["aid"], "aid_next": g["aid_next"], "count": g["count_sum"]})


def _sort_canonical(df: pa.Table) -> pa.Table:
    return df.sort_by([("count", "descending"), ("aid", "ascending"), ("aid_next", "ascending")])


def top_n_per_aid(df_count: pa.Table, first_n: int) -> pa.Table:
    """retrieve.py:41-51 -- ordinal rank of count (descending) over aid, keep rank <= first_n.

    Canonical tie rule: (count desc, aid_next asc).  Returns columns
    aid, aid_next, count, rank (1-based, int16) ordered by (aid, rank).
    """
    df = df_count.sort_by([("aid", "ascending"), ("count", "descending"), ("aid_next", "ascending")])
    aid = df["aid"].to_numpy()
    n = len(aid)
    if n == 0:
        rank = np.zeros(0, np.int64)
    else:
        start = np.r_[True, aid[1:] != aid[:-1]]
        seg_start = np.maximum.accumulate(np.where(start, np.arange(n), 0))
        rank = np.arange(n) - seg_start + 1
    keep = rank <= first_n
    df = df.filter(pa.array(keep))
    return df.append_column("rank", pa.array(rank[keep].astype(np.int16)))


def count_features(df_count: pa.Table, count_type: str, first_n: Optional[int] = None) -> Dict[str, np.ndarray]:
    """get_df_count_for_co_event_type (retrieve.py:18-63) on an in-memory count table given in FILE order.

    count_pop  = int16(clip_max((count - min) / (quantile_0.9999 - min), 1) * 10_000)      :33-35
    perc_pop   = int16(row_nr / n * 10_000), row_nr from 1 in file order                    :36-38
    rank       = ordinal rank of count, descending, over aid (ties keep file order)         :41-44
    count_rel  = int8(count / max(count over aid) * 100)                                    :45-49
    keep rank <= first_n; rows come back ordered by aid, then rank.
    """
    first_n = RETRIEVAL_FIRST_N[count_type] if first_n is None else first_n
    aid = df_count["aid"].to_numpy(); nxt = df_count["aid_next"].to_numpy()
    cnt = df_count["count"].to_numpy().astype(np.int64)
    n = len(aid)
    cmin = cnt.min() if n else 0
    q = np.quantile(cnt, 0.9999, method="nearest") if n else 0
    denom = max(float(q - cmin), 1e-12)
    count_pop = (np.minimum((cnt - cmin) / denom, 1.0) * 10_000).astype(np.int16)
    perc_pop = (np.arange(1, n + 1) / max(n, 1) * 10_000).astype(np.int16)
    order = np.argsort(aid, kind="stable")                       # sort(['aid']) keeps file order inside an aid
    a, b, c = aid[order], nxt[order], cnt[order]
    start = np.r_[True, a[1:] != a[:-1]] if n else np.zeros(0, bool)
    seg = np.maximum.accumulate(np.where(start, np.arange(n), 0)) if n else np.zeros(0, np.int64)
    # ordinal rank, descending: position after a stable sort by -count inside the aid
    o2 = np.lexsort((np.arange(n), -c, a))
    rank = np.empty(n, np.int64); rank[o2] = np.arange(n) - seg[o2] + 1 if n else 0
    seg_id = np.cumsum(start) - 1 if n else np.zeros(0, np.int64)
    mx = np.maximum.reduceat(c, np.flatnonzero(start))[seg_id] if n else np.zeros(0, np.int64)
    keep = rank <= first_n
    o3 = np.lexsort((rank[keep], a[keep]))
    out = {"aid": a[keep][o3], "aid_next": b[keep][o3], f"{count_type}_count": c[keep][o3],
           f"{count_type}_count_pop": count_pop[order][keep][o3], f"{count_type}_perc_pop": perc_pop[order][keep][o3],
           f"{count_type}_rank": rank[keep][o3].astype(np.int16),
           f"{count_type}_count_rel": (c[keep] / mx[keep] * 100).astype(np.int8)[o3]}
    return out


# ---- convenience wrappers used by tests and bench ------------------------------------------
def table_to_dict(t: pa.Table) -> Dict[tuple, int]:
    a, b, c = (t[k].to_numpy() for k in ("aid", "aid_next", "count"))
    return {(int(x), int(y)): int(z) for x, y, z in zip(a, b, c)}


def count_events_all_names(session, aid, ts, type_, n_sessions_in_part: int = 10_000
                           ) -> Dict[str, pa.Table]:
    return count_part(events_table(session, aid, ts, type_), n_sessions_in_part)


def pipeline_single_population(session, aid, ts, type_, names: Optional[List[str]] = None,
                               part_sessions: int = 100_000, exact: bool = True,
                               min_count: Optional[Dict[str, int]] = None) -> Dict[str, pa.Table]:
    """Phase 1 (per 100k-session part) + phase 2 for ONE population, in memory."""
    names = names or CO_EVENTS_TO_COUNT
    session = np.asarray(session)
    order = np.argsort(session, kind="stable")
    session, aid, ts, type_ = (np.asarray(x)[order] for x in (session, aid, ts, type_))
    uniq = np.unique(session)
    parts: Dict[str, List[pa.Table]] = {n: [] for n in names}
    for lo in range(0, len(uniq), part_sessions):
        hi = min(lo + part_sessions, len(uniq))
        a = np.searchsorted(session, uniq[lo], "left")
        b = np.searchsorted(session, uniq[hi - 1], "right")
        res = count_events_all_names(session[a:b], aid[a:b], ts[a:b], type_[a:b])
        for n in names:
            parts[n].append(res[n])
    out = {}
    for n in names:
        mc = None if min_count is None else min_count.get(n)
        out[n] = merge_counts(n, parts[n] or [_empty_counts()], exact=exact, min_count_to_save=mc)
    return out


def _empty_counts() -> pa.Table:
    return pa.table({"aid": pa.array([], pa.int32()), "aid_next": pa.array([], pa.int32()),
                     "count": pa.array([], pa.uint32())})
