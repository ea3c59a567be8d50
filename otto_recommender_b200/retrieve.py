"""Per-aid top-N of a co-event count table + the derived count features of the consumer.

Mirror of ``get_df_count_for_co_event_type`` (reference model/retrieve.py:18-63): same arguments,
same output columns.  The segmented top-N (retrieve.py:41-47), the population-level features
(:33-38) and ``count_rel`` (:49) all run on the GPU (``ottocov_count_features``, csrc/features.cu).

Tie rule: the reference ranks with ``rank('ordinal', reverse=True).over('aid')`` on a frame sorted
by aid only, so rows with equal count keep file order, which is itself unspecified upstream.  Here
ties are ordered by aid_next ascending.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import pandas as pd
import pyarrow.parquet as pq

from .config import DEFAULT_CONFIG
from .count_co_events import get_engine


def topn_long(ax, nv, ay, ac):
    """(aid_x[A], n_valid[A], aid_y[A,k], cnt[A,k]) -> long-format (aid, aid_next, count, rank)."""
    A, k = ay.shape
    mask = np.arange(k)[None, :] < nv[:, None]
    aid = np.repeat(ax, nv)
    rank = (np.broadcast_to(np.arange(1, k + 1)[None, :], (A, k)))[mask]
    return aid, ay[mask], ac[mask], rank.astype(np.int16)


def get_df_count_for_co_event_type(count_type: str, dir_counts: str, first_n: Optional[int] = None,
                                   config=DEFAULT_CONFIG) -> pd.DataFrame:
    if first_n is None:
        first_n = config.RETRIEVAL_FIRST_N_CO_COUNTS[count_type]
    t = pq.read_table(f"{dir_counts}/{count_type}.parquet", columns=["aid", "aid_next", "count"])
    aid, aid_next, count = (t[c].to_numpy() for c in ("aid", "aid_next", "count"))
    cols = [f"{count_type}_count", f"{count_type}_count_pop", f"{count_type}_perc_pop", f"{count_type}_rank",
            f"{count_type}_count_rel"]
    if len(aid) == 0:
        return pd.DataFrame({"aid": np.zeros(0, np.int32), "aid_next": np.zeros(0, np.int32),
                             **{c: np.zeros(0, np.int32) for c in cols}})
    # everything below the file read runs on the GPU (ottocov_count_features): the population statistics
    # (retrieve.py:33-38), the segmented top-N (:41-47) and count_rel (:45-49)
    f = get_engine().count_features(aid, aid_next, count, first_n)
    return pd.DataFrame({"aid": f["aid"], "aid_next": f["aid_next"], cols[0]: f["count"], cols[1]: f["count_pop"],
                         cols[2]: f["perc_pop"], cols[3]: f["rank"], cols[4]: f["count_rel"]})


def get_pairs_co_event_type(df_aids, df_count, type: int = 0) -> pd.DataFrame:
    """Mirror of retrieve.py:75-91: unique aids of `df_aids` joined with the (top-N) count rows on aid.
    `df_count` is the frame returned by get_df_count_for_co_event_type (or any frame with aid, aid_next);
    `type` is unused, exactly as in the reference (its type filter is commented out).  The join runs as a
    GPU lookup over the per-aid top-K rows."""
    aids = np.unique(np.asarray(df_aids["aid"], dtype=np.int32))
    a = np.asarray(df_count["aid"], dtype=np.int32)
    b = np.asarray(df_count["aid_next"], dtype=np.int32)
    if len(a) == 0 or len(aids) == 0:
        return pd.DataFrame({"aid": np.zeros(0, np.int32), "aid_next": np.zeros(0, np.int32)})
    eng = get_engine()
    # rows per aid in df_count are already the kept top-N rows: rank them by position to rebuild the matrix
    tab = eng.table_from_arrays(a, b, np.ones(len(a), np.uint32))
    k = int(min(32, max(1, np.bincount(np.unique(a, return_inverse=True)[1]).max())))
    eng.topk(tab, k)
    tab.free()
    nv, ay, _ = eng.topk_lookup(aids)
    mask = np.arange(ay.shape[1])[None, :] < nv[:, None]
    return pd.DataFrame({"aid": np.repeat(aids, nv), "aid_next": ay[mask]})
