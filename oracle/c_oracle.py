"""ctypes front for oracle/cov_oracle.c (plain-C brute force).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (no tests / polars absent); pinned by tests/golden/ and by
agreement with oracle/ref_restatement.py.  Semantics: model/count_co_events.py:17-77,92 and
config.py:41-49,81-88 of the reference; top-N: model/retrieve.py:41-51.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcovoracle.so")
_lib = None

NAMES = {
    # name: (type_this, next_mask, window_s)   config.py:43-49, 81-88
    "click_to_click": (0, 0b001, 12 * 3600),
    "click_to_cart_or_buy": (0, 0b110, 24 * 3600),
    "cart_to_cart": (1, 0b010, 24 * 3600),
    "cart_to_buy": (1, 0b100, 24 * 3600),
    "buy_to_buy": (2, 0b100, 24 * 3600),
}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cov_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        P = ctypes.POINTER
        lib.cov_oracle_count.restype = ctypes.c_int64
        lib.cov_oracle_count.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int64,
            P(ctypes.c_void_p), P(ctypes.c_void_p), P(ctypes.c_void_p),
            P(ctypes.c_int64), P(ctypes.c_int64)]
        lib.cov_oracle_count_range.restype = ctypes.c_int64
        lib.cov_oracle_count_range.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            P(ctypes.c_void_p), P(ctypes.c_void_p), P(ctypes.c_void_p),
            P(ctypes.c_int64), P(ctypes.c_int64)]
        lib.cov_oracle_score.restype = ctypes.c_int64
        lib.cov_oracle_score.argtypes = [
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            P(ctypes.c_void_p), P(ctypes.c_void_p), P(ctypes.c_void_p), P(ctypes.c_void_p)]
        lib.cov_oracle_free.argtypes = [ctypes.c_void_p]
        lib.cov_oracle_free.restype = None
        _lib = lib
    return _lib


def count(session, aid, ts, type_, type_this: int, next_mask: int, window: int,
          dt_min: int = -86400, dt_max: int = 86400) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int, int]:
    """-> (aid, aid_next, count) sorted by (aid, aid_next), n_emitted_pairs, n_events_after_dedup.
    dt_min / dt_max: config.MIN_TIME_TO_NEXT / MAX_TIME_TO_NEXT (the pre-filter of count_co_events.py:33-36)."""
    lib = _load()
    s = np.ascontiguousarray(session, np.int32)
    a = np.ascontiguousarray(aid, np.int32)
    t = np.ascontiguousarray(ts, np.int32)
    y = np.ascontiguousarray(type_, np.int8)
    n = len(s)
    pa_, pb_, pc_ = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    emitted, nded = ctypes.c_int64(), ctypes.c_int64()
    u = lib.cov_oracle_count_range(n, s.ctypes.data, a.ctypes.data, t.ctypes.data, y.ctypes.data,
                                   type_this, next_mask, window, dt_min, dt_max,
                                   ctypes.byref(pa_), ctypes.byref(pb_), ctypes.byref(pc_),
                                   ctypes.byref(emitted), ctypes.byref(nded))
    if u < 0:
        raise MemoryError("cov_oracle_count failed")
    try:
        def take(p, ct, dt):
            if u == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), shape=(u,)).astype(dt, copy=True)
        oa, ob, oc = take(pa_, ctypes.c_int32, np.int32), take(pb_, ctypes.c_int32, np.int32), \
            take(pc_, ctypes.c_uint32, np.uint32)
    finally:
        for p in (pa_, pb_, pc_):
            lib.cov_oracle_free(p)
    return oa, ob, oc, int(emitted.value), int(nded.value)


def score(session, aid, ts, type_, type_this: int, next_mask: int, window: int, dt_min: int = -86400, dt_max: int = 86400):
    """EXTENSION oracle (no reference counterpart): time-decay weighted scores in float64,
    w = max(0.10, 1 - |dt| / window) per counted pair.  -> (aid, aid_next, score f64, count u32) sorted by (aid, aid_next)."""
    lib = _load()
    s = np.ascontiguousarray(session, np.int32); a = np.ascontiguousarray(aid, np.int32)
    t = np.ascontiguousarray(ts, np.int32); y = np.ascontiguousarray(type_, np.int8)
    pa_, pb_, ps_, pc_ = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    u = lib.cov_oracle_score(len(s), s.ctypes.data, a.ctypes.data, t.ctypes.data, y.ctypes.data, type_this, next_mask, window,
                             dt_min, dt_max, ctypes.byref(pa_), ctypes.byref(pb_), ctypes.byref(ps_), ctypes.byref(pc_))
    if u < 0:
        raise MemoryError("cov_oracle_score failed")
    try:
        def take(p, ct, dt):
            if u == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), shape=(u,)).astype(dt, copy=True)
        out = (take(pa_, ctypes.c_int32, np.int32), take(pb_, ctypes.c_int32, np.int32), take(ps_, ctypes.c_double, np.float64),
               take(pc_, ctypes.c_uint32, np.uint32))
    finally:
        for p in (pa_, pb_, ps_, pc_):
            lib.cov_oracle_free(p)
    return out


def score_name(session, aid, ts, type_, name: str):
    th, mask, w = NAMES[name]
    return score(session, aid, ts, type_, th, mask, w)


def count_name(session, aid, ts, type_, name: str):
    th, mask, w = NAMES[name]
    return count(session, aid, ts, type_, th, mask, w)


def merge_tables(tables, min_count: int = 1):
    """Sum counts by (aid, aid_next) over several (aid, aid_next, count) triples, keep
    count >= min_count (count_co_events.py:168-172).  Returns arrays sorted by (aid, aid_next)."""
    if not tables:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.int64)
    a = np.concatenate([np.asarray(t[0], np.int64) for t in tables])
    b = np.concatenate([np.asarray(t[1], np.int64) for t in tables])
    c = np.concatenate([np.asarray(t[2], np.int64) for t in tables])
    key = (a << 32) | b
    uk, inv = np.unique(key, return_inverse=True)
    s = np.zeros(len(uk), np.int64)
    np.add.at(s, inv, c)
    keep = s >= min_count
    uk, s = uk[keep], s[keep]
    return (uk >> 32).astype(np.int32), (uk & 0xFFFFFFFF).astype(np.int32), s


def sort_count_desc(aid, aid_next, cnt):
    """Canonical file order: (count desc, aid asc, aid_next asc) -- count_co_events.py:173."""
    order = np.lexsort((aid_next, aid, -np.asarray(cnt, np.int64)))
    return aid[order], aid_next[order], np.asarray(cnt)[order]


def top_n(aid, aid_next, cnt, n: int):
    """retrieve.py:41-51 with the canonical tie rule (count desc, aid_next asc).
    Returns (aid, aid_next, count, rank) ordered by (aid, rank)."""
    aid = np.asarray(aid); aid_next = np.asarray(aid_next); cnt = np.asarray(cnt, np.int64)
    order = np.lexsort((aid_next, -cnt, aid))
    aid, aid_next, cnt = aid[order], aid_next[order], cnt[order]
    m = len(aid)
    if m == 0:
        return aid, aid_next, cnt, np.zeros(0, np.int64)
    start = np.r_[True, aid[1:] != aid[:-1]]
    seg = np.maximum.accumulate(np.where(start, np.arange(m), 0))
    rank = np.arange(m) - seg + 1
    k = rank <= n
    return aid[k], aid_next[k], cnt[k], rank[k]
