"""Host-side mirror of the C ABI: one Engine per GPU, Tables of (aid, aid_next, count).

The reference keeps these frames in polars (model/count_co_events.py); here they live in HBM and
only leave it through ``Table.fetch`` / ``Engine.topk``.  numpy arrays are passed as host
pointers, torch CUDA tensors as device pointers (no copies, no torch types cross the ABI).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import OttocovError
from .config import DEFAULT_CONFIG, CoEventConfig


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _as_col(x, np_dtype, torch_dtype_name):
    """-> (pointer, where, keepalive).  A column wider than its target type is range-checked before it is narrowed:
    the C side only ever sees int32 / int8 values, so raw OTTO millisecond timestamps (~1.66e12) or 64-bit session
    ids would otherwise wrap silently.  `ts` must be in SECONDS, as etl/jsonl_to_parquet.py:28 writes it."""
    if _is_torch(x):
        import torch
        dt = getattr(torch, torch_dtype_name)
        if x.dtype == dt and x.is_contiguous():                 # the common case: nothing to convert, nothing to check
            return x.data_ptr(), (_lib.DEVICE if x.is_cuda else _lib.HOST), x
    elif isinstance(x, np.ndarray) and x.dtype == np_dtype and x.flags.c_contiguous:
        return x.ctypes.data, _lib.HOST, x
    info = np.iinfo(np_dtype)
    if _is_torch(x):
        import torch
        dt = getattr(torch, torch_dtype_name)
        if x.dtype != dt and x.numel() and not (np_dtype == np.uint64 and x.dtype == torch.int64):
            if x.dtype.is_floating_point and not bool((x == x.floor()).all()):
                raise ValueError(f"column of dtype {x.dtype} with non-integral values where {torch_dtype_name} is expected")
            lo, hi = int(x.min()), int(x.max())
            if lo < info.min or hi > info.max:
                raise ValueError(f"column values [{lo}, {hi}] do not fit {torch_dtype_name} (timestamps must be seconds)")
        t = x.to(dt).contiguous()
        return t.data_ptr(), (_lib.DEVICE if t.is_cuda else _lib.HOST), t
    a = np.asarray(x)
    if a.dtype != np_dtype and a.size:
        if a.dtype.kind not in "iufb" or (a.dtype.kind == "f" and not np.all(a == np.floor(a))):
            raise ValueError(f"column of dtype {a.dtype} where {np.dtype(np_dtype).name} is expected")
        lo, hi = int(a.min()), int(a.max())
        if lo < info.min or hi > info.max:
            raise ValueError(f"column values [{lo}, {hi}] do not fit {np.dtype(np_dtype).name} (timestamps must be seconds)")
    a = np.ascontiguousarray(a, dtype=np_dtype)
    return a.ctypes.data, _lib.HOST, a


class Table:
    """A device-resident (aid, aid_next, count) table: rows sorted by (aid, aid_next), distinct."""

    def __init__(self, engine: "Engine", handle: int):
        self._e = engine
        self._h = ctypes.c_void_p(handle)

    @property
    def rows(self) -> int:
        n = ctypes.c_int64()
        self._e._check(self._e._lib.ottocov_table_rows(self._h, ctypes.byref(n)))
        return int(n.value)

    def total(self) -> int:
        s = ctypes.c_int64()
        self._e._check(self._e._lib.ottocov_table_total(self._e._ctx, self._h, ctypes.byref(s)))
        return int(s.value)

    def fetch(self, order: str = "key", head: int = -1, device: bool = False, pinned: bool = False):
        """-> (aid, aid_next, count) as int32 numpy arrays (or torch CUDA tensors if device=True).
        order='count_desc' is the file order of model/count_co_events.py:173-175.  pinned=True
        lands the rows in reusable page-locked buffers (views valid until the next pinned fetch)."""
        n = self.rows if head < 0 else min(head, self.rows)
        o = _lib.ORDER_COUNT_DESC if order == "count_desc" else _lib.ORDER_KEY
        n_out = ctypes.c_int64()
        if device:
            import torch
            dev = torch.device("cuda", self._e.device)
            a, b, c = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3))
            pa, pb, pc, where = a.data_ptr(), b.data_ptr(), c.data_ptr(), _lib.DEVICE
        else:
            if pinned:
                a, b, c = (self._e._pinned(f"fetch{i}", n) for i in range(3))
            else:
                a, b, c = (np.empty(n, np.int32) for _ in range(3))
            pa, pb, pc, where = a.ctypes.data, b.ctypes.data, c.ctypes.data, _lib.HOST
        self._e._sync_stream()
        self._e._check(self._e._lib.ottocov_table_fetch(self._e._ctx, self._h, o, head, pa, pb, pc, n, where,
                                                        ctypes.byref(n_out)))
        return a, b, c

    def to_dict(self) -> Dict[Tuple[int, int], int]:
        a, b, c = self.fetch()
        return {(int(x), int(y)): int(z) for x, y, z in zip(a, b, c)}

    def device_ptrs(self) -> Tuple[int, int]:
        k, c = ctypes.c_void_p(), ctypes.c_void_p()
        self._e._check(self._e._lib.ottocov_table_device_ptrs(self._h, ctypes.byref(k), ctypes.byref(c)))
        return int(k.value or 0), int(c.value or 0)

    def free(self):
        if self._h is not None and self._e._ctx is not None:
            self._e._lib.ottocov_table_free(self._e._ctx, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class WeightedTable:
    """EXTENSION (no reference counterpart): device-resident (aid, aid_next, score, count) rows of
    ``Engine.count_weighted`` -- time-decay weighted co-event scores, w = max(0.10, 1 - |dt| / window) per pair."""

    def __init__(self, engine: "Engine", handle: int):
        self._e = engine
        self._h = ctypes.c_void_p(handle)

    @property
    def rows(self) -> int:
        n = ctypes.c_int64()
        self._e._check(self._e._lib.ottocov_wtable_rows(self._h, ctypes.byref(n)))
        return int(n.value)

    def fetch(self):
        """-> aid i32, aid_next i32, score f64, count i32; rows in (aid, aid_next) order."""
        n = self.rows
        a, b, c = (np.empty(n, np.int32) for _ in range(3))
        s = np.empty(n, np.float64)
        got = ctypes.c_int64()
        self._e._sync_stream()
        self._e._check(self._e._lib.ottocov_wtable_fetch(self._e._ctx, self._h, a.ctypes.data, b.ctypes.data, s.ctypes.data,
                                                         c.ctypes.data, n, _lib.HOST, ctypes.byref(got)))
        return a, b, s, c

    def topk(self, k: int = 20):
        """Per-aid top-k by (score desc, aid_next asc), long format ordered by (aid, rank): aid, aid_next, score, rank."""
        n = self.rows
        a, b, r = (np.empty(n, np.int32) for _ in range(3))
        s = np.empty(n, np.float64)
        got = ctypes.c_int64()
        self._e._sync_stream()
        self._e._check(self._e._lib.ottocov_wtable_topk(self._e._ctx, self._h, int(k), a.ctypes.data, b.ctypes.data, s.ctypes.data,
                                                        r.ctypes.data, n, _lib.HOST, ctypes.byref(got)))
        m = int(got.value)
        return a[:m], b[:m], s[:m], r[:m]

    def topk_rows(self, k: int = 20) -> int:
        """Runs the per-aid top-k on the device without copying it out; -> rows it has."""
        got = ctypes.c_int64()
        self._e._sync_stream()
        self._e._check(self._e._lib.ottocov_wtable_topk(self._e._ctx, self._h, int(k), None, None, None, None, 0, _lib.HOST,
                                                        ctypes.byref(got)))
        return int(got.value)

    def free(self):
        if self._h is not None and self._e._ctx is not None:
            self._e._lib.ottocov_wtable_free(self._e._ctx, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One co-visitation counting context on one GPU (``ottocov_ctx``)."""

    def __init__(self, device: int = 0, config: CoEventConfig = DEFAULT_CONFIG, profiling: bool = False):
        self._lib = _lib.load_library()          # raises if the CUDA library is not built
        self.config = config
        self.device = int(device)
        ctx = ctypes.c_void_p()
        rc = self._lib.ottocov_create(self.device, ctypes.byref(ctx))
        if rc != 0:
            msg = self._lib.ottocov_last_error(None)
            raise OttocovError(rc, msg.decode() if msg else "ottocov_create failed")
        self._ctx = ctx
        self._stream = None
        self.load_generation = 0          # bumped by every load_events call
        if profiling:
            self.set_profiling(True)

    # ---- plumbing ----------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.ottocov_last_error(self._ctx)
            raise OttocovError(rc, msg.decode() if msg else "")

    def _sync_stream(self):
        """Run on torch's current stream when torch has CUDA initialised (so torch CUDA events and
        tensors produced by torch ops order correctly with our kernels)."""
        try:
            import sys
            torch = sys.modules.get("torch")
            if torch is None or not torch.cuda.is_initialized():
                return
            s = torch.cuda.current_stream(self.device).cuda_stream
        except Exception:
            return
        if s != self._stream:
            self._check(self._lib.ottocov_set_stream(self._ctx, ctypes.c_void_p(s)))
            self._stream = s

    def _pinned(self, tag: str, n: int):
        """Reusable page-locked int32 staging buffer (numpy view); valid until the next call with `tag`."""
        import torch
        cache = self.__dict__.setdefault("_pin_cache", {})
        buf = cache.get(tag)
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=torch.int32).pin_memory()
            cache[tag] = buf
        return buf[:n].numpy()

    def close(self):
        if self._ctx is not None:
            self._lib.ottocov_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self._lib.ottocov_synchronize(self._ctx))

    def trim(self):
        """Give cached device blocks back to the driver."""
        self._check(self._lib.ottocov_trim(self._ctx))

    def memory_info(self) -> Dict[str, int]:
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.ottocov_memory_info(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"live_bytes": a.value, "cached_bytes": b.value, "peak_bytes": c.value}

    def set_profiling(self, on, families=None):
        """on=True: CUDA-event timing of every kernel family; families=[...] restricts it (event
        records between back-to-back launches cost a few microseconds of GPU idle each)."""
        mask = int(bool(on))
        if on and families:
            names = [self._lib.ottocov_kernel_family_name(i).decode() for i in range(_lib.K_FAMILIES)]
            mask = 0
            for f in families:
                mask |= 1 << names.index(f)
            if mask == 1:                 # family 0 alone would read as "all": add a harmless second bit
                mask |= 1 << names.index("misc")
        self._check(self._lib.ottocov_set_profiling(self._ctx, mask))

    def kernel_stats(self, reset: bool = False) -> Dict[str, Dict[str, float]]:
        arr = (_lib.KernelStat * _lib.K_FAMILIES)()
        self._check(self._lib.ottocov_kernel_stats(self._ctx, arr, int(reset)))
        out = {}
        for i in range(_lib.K_FAMILIES):
            name = self._lib.ottocov_kernel_family_name(i).decode()
            out[name] = {"launches": int(arr[i].launches), "ms": float(arr[i].ms),
                         "algo_bytes": float(arr[i].algo_bytes)}
        return out

    # ---- (1) loader ----------------------------------------------------------------------------------
    def load_events(self, session, aid, ts, type_) -> Dict[str, int]:
        """pl.read_parquet(file).unique() + grouping by session (count_co_events.py:91-92, :19)."""
        self._sync_stream()
        ps, ws, ks = _as_col(session, np.int32, "int32")
        pa, wa, ka = _as_col(aid, np.int32, "int32")
        pt, wt, kt = _as_col(ts, np.int32, "int32")
        py, wy, ky = _as_col(type_, np.int8, "int8")
        if len({ws, wa, wt, wy}) != 1:
            raise ValueError("all four columns must live on the same side (host or device)")
        n = len(ks)
        if not (len(ka) == len(kt) == len(ky) == n):
            raise ValueError("columns differ in length")
        self.load_generation += 1
        self._check(self._lib.ottocov_load_events(self._ctx, ps, pa, pt, py, n, ws))
        return self.events_info()

    def events_info(self) -> Dict[str, int]:
        info = _lib.EventsInfo()
        self._check(self._lib.ottocov_get_events_info(self._ctx, ctypes.byref(info)))
        d = {k: getattr(info, k) for k, _ in _lib.EventsInfo._fields_ if k != "n_by_type"}
        d["n_by_type"] = [int(x) for x in info.n_by_type]
        return d

    # ---- (2)+(3) expansion + reduce-by-key ------------------------------------------------------------
    def count(self, name: Optional[str] = None, *, type_this: Optional[int] = None,
              next_types: Optional[Sequence[int]] = None, window: Optional[int] = None,
              pair_budget: Optional[int] = None, min_count: int = 1,
              symmetric: Optional[bool] = None, hashed: Optional[bool] = None) -> Table:
        """One iteration of count_co_events' loop (count_co_events.py:64-72) on the loaded events.
        min_count > 1 fuses filter(count >= min_count) (count_co_events.py:172) into the reduce.
        hashed: force (True) or forbid (False) the bucketed hash reduce; None = auto (min_count > 1).
        Both strategies return the same table."""
        spec = self._spec(name, type_this, next_types, window, pair_budget, min_count, symmetric, hashed)
        h = ctypes.c_void_p()
        self._sync_stream()
        self._check(self._lib.ottocov_count(self._ctx, ctypes.byref(spec), ctypes.byref(h)))
        return Table(self, h.value)

    def _spec(self, name, type_this, next_types, window, pair_budget, min_count, symmetric, hashed=None):
        if name is not None:
            th, mask, w = self.config.spec(name)
        else:
            th, mask, w = int(type_this), 0, int(window)
            for t in next_types:
                mask |= 1 << int(t)
        if window is not None:
            w = int(window)
        budget = self.config.PAIR_BUDGET if pair_budget is None else int(pair_budget)
        flags = 0 if symmetric is None else (2 if symmetric else 1)
        if hashed is not None:
            flags |= 8 if hashed else 4
        # the +-24 h pre-filter of self_merge (config.MIN/MAX_TIME_TO_NEXT, count_co_events.py:33-36)
        dt_min, dt_max = int(self.config.MIN_TIME_TO_NEXT), int(self.config.MAX_TIME_TO_NEXT)
        if (dt_min, dt_max) != (-86400, 86400):
            flags |= 16
        return _lib.Spec(th, mask, w, budget, max(int(min_count), 0), flags, dt_min, dt_max)

    def count_weighted(self, name: Optional[str] = None, *, type_this: Optional[int] = None,
                       next_types: Optional[Sequence[int]] = None, window: Optional[int] = None,
                       min_count: int = 1) -> WeightedTable:
        """EXTENSION (north_star config 4; the reference has no weighting): the pairs of `count`, each
        contributing w = max(0.10, 1 - |dt| / window) to score(aid, aid_next); rows keep the integer count and
        min_count thresholds it.  Exact fixed-point sums (24 fractional bits): within 3e-7 of a float64 evaluation."""
        spec = self._spec(name, type_this, next_types, window, None, min_count, False, None)
        h = ctypes.c_void_p()
        self._sync_stream()
        self._check(self._lib.ottocov_count_weighted(self._ctx, ctypes.byref(spec), ctypes.byref(h)))
        return WeightedTable(self, h.value)

    # ---- streamed ingest + count ------------------------------------------------------------------------------
    def count_parts(self, parts, names: Sequence[str], min_counts: Optional[Sequence[int]] = None) -> List[Table]:
        """The per-file loop of count_co_events_all_files + the merge of the part tables for one population
        (count_co_events.py:83-100, :112-168), with the host -> device copy of the parts overlapped with the counting.

        parts: list of (session, aid, ts, type) column tuples as the ETL wrote them (one tuple per parquet part; a
        session never spans two parts).  Host arrays (numpy, or torch CPU tensors -- pinned ones copy asynchronously).
        Returns one Table per name, identical to load_events(concatenation) + count(name, min_count=...)."""
        keep = []
        n_parts = len(parts)
        cols = [(ctypes.c_void_p * max(n_parts, 1))() for _ in range(4)]
        rows = (ctypes.c_int64 * max(n_parts, 1))()
        for i, p in enumerate(parts):
            s, a, t, y = p
            n = len(s)
            for j, (x, npd, tdn) in enumerate(((s, np.int32, "int32"), (a, np.int32, "int32"), (t, np.int32, "int32"),
                                               (y, np.int8, "int8"))):
                ptr, where, k = _as_col(x, npd, tdn)
                if where != _lib.HOST:
                    raise ValueError("count_parts takes host columns (the parts as read from parquet)")
                if len(k) != n:
                    raise ValueError("columns of a part differ in length")
                keep.append(k)
                cols[j][i] = ptr
            rows[i] = n
        if min_counts is None:
            min_counts = [self.config.MIN_COUNT_TO_SAVE.get(nm, 1) for nm in names]
        specs = (_lib.Spec * len(names))()
        for k, (nm, mc) in enumerate(zip(names, min_counts)):
            specs[k] = self._spec(nm, None, None, None, None, mc, None, None)
        out = (ctypes.c_void_p * len(names))()
        self._sync_stream()
        self.load_generation += 1
        self._check(self._lib.ottocov_count_parts(self._ctx, n_parts, cols[0], cols[1], cols[2], cols[3], rows, specs,
                                                  len(names), out))
        del keep
        return [Table(self, h) for h in out]

    @staticmethod
    def split_at_sessions(session, aid, ts, type_, n_parts: int):
        """One population in contiguous host columns ordered by session -> n_parts column-view tuples cut at session
        boundaries (what reading the ETL's parts would give).  O(n_parts log n) on the host, no copies."""
        s = session.numpy() if _is_torch(session) else np.asarray(session)
        n = len(s)
        cuts = [0]
        for i in range(1, n_parts):
            r = n * i // n_parts
            if r <= cuts[-1] or r >= n:
                continue
            r = int(np.searchsorted(s, s[r], side="left"))         # back to the first row of that session
            if r > cuts[-1]:
                cuts.append(r)
        cuts.append(n)
        return [tuple(x[cuts[i]:cuts[i + 1]] for x in (session, aid, ts, type_)) for i in range(len(cuts) - 1)]

    # ---- exchange-before-reduce building blocks (multi-GPU) ---------------------------------------------
    def expand_prepare(self, name: Optional[str] = None, *, type_this=None, next_types=None, window=None,
                       min_count: int = 1, symmetric: Optional[bool] = None) -> Tuple[int, bool]:
        """Window pass only: -> (number of keys expand_run will emit, keys are symmetric half pairs)."""
        spec = self._spec(name, type_this, next_types, window, None, min_count, symmetric)
        n, sym = ctypes.c_int64(), ctypes.c_int()
        self._sync_stream()
        self._check(self._lib.ottocov_expand_prepare(self._ctx, ctypes.byref(spec), ctypes.byref(n), ctypes.byref(sym)))
        return int(n.value), bool(sym.value)

    def expand_run(self, n_ranks: int, buf_a, buf_b=None):
        """Emit the prepared keys into caller tensors (int64, CUDA), grouped by destination rank when
        buf_b is given (else only stamped + counted: see push_keys).
        -> (tensor holding the keys, rows_per_dest)."""
        rows = (ctypes.c_int64 * n_ranks)()
        in_b = ctypes.c_int()
        self._sync_stream()
        self._check(self._lib.ottocov_expand_run(self._ctx, n_ranks, buf_a.data_ptr(),
                                                 buf_b.data_ptr() if buf_b is not None else None,
                                                 ctypes.byref(in_b), rows))
        return (buf_b if in_b.value else buf_a), [int(x) for x in rows]

    def push_keys(self, keys, n: int, dest_ptrs: Sequence[int]):
        """Fused partition + exchange: one distribution pass over stamped keys whose per-destination runs
        start at dest_ptrs[rank] (device byte addresses, typically peer memory)."""
        arr = (ctypes.c_uint64 * len(dest_ptrs))(*[int(p) for p in dest_ptrs])
        self._sync_stream()
        self._check(self._lib.ottocov_push_keys(self._ctx, keys.data_ptr() if n else None, int(n), len(dest_ptrs), arr))

    def reduce_pairs(self, keys, n: int, aid_bits: int, min_count: int = 1, symmetric: bool = False,
                     strip_dest: bool = False, hashed: Optional[bool] = None) -> Table:
        """Reduce-by-key of raw keys (an int64 CUDA tensor, used as scratch): sort + run-length count, or the
        bucketed hash reduce (auto when min_count > 1; `hashed` forces the choice)."""
        h = ctypes.c_void_p()
        self._sync_stream()
        mode = int(bool(strip_dest)) | (0 if hashed is None else (2 if hashed else 4))
        self._check(self._lib.ottocov_reduce_pairs(self._ctx, keys.data_ptr() if n else None, int(n), int(aid_bits),
                                                   max(int(min_count), 0), int(symmetric), mode,
                                                   ctypes.byref(h)))
        return Table(self, h.value)

    # ---- fused expansion + exchange (default multi-GPU path) ------------------------------------------------
    def make_xplan(self, n_ranks: int, aid_bits: int, max_local_keys: int, total_keys: int, stripe_cap: int = 0,
                   mirror_cap: int = 0) -> "_lib.XPlan":
        """Stripe capacities + layout of a rank's receive area; pure arithmetic, identical on every rank."""
        plan = _lib.XPlan()
        rc = self._lib.ottocov_xplan_make(int(n_ranks), int(aid_bits), int(max_local_keys), int(total_keys), int(stripe_cap),
                                          int(mirror_cap), ctypes.byref(plan))
        if rc != 0:
            msg = self._lib.ottocov_last_error(None)
            raise OttocovError(rc, msg.decode() if msg else "ottocov_xplan_make failed")
        return plan

    def expand_scatter(self, plan, rank: int, peer_bases: Sequence[int]):
        """After expand_prepare: expand this rank's keys straight into their owners' stripes (peer_bases[r] = device
        address of rank r's receive area) and publish counts / histograms / status to every rank.  Enqueue only."""
        arr = (ctypes.c_uint64 * len(peer_bases))(*[int(p) for p in peer_bases])
        self._sync_stream()
        self._check(self._lib.ottocov_expand_scatter(self._ctx, ctypes.byref(plan), int(rank), arr))

    def reduce_received(self, plan, recv_area: int, min_count: int = 1, symmetric: bool = False):
        """Remaining passes + hash reduce over the stripes this rank received.  -> (Table | None, need_cap): need_cap > 0
        means some rank overflowed a stripe (every rank reads the same value): grow the plan and repeat the step."""
        h, need = ctypes.c_void_p(), ctypes.c_int64()
        self._sync_stream()
        self._check(self._lib.ottocov_reduce_received(self._ctx, ctypes.byref(plan), int(recv_area), max(int(min_count), 0),
                                                      int(symmetric), ctypes.byref(h), ctypes.byref(need)))
        return (Table(self, h.value) if h.value else None), int(need.value)

    def mirror_push(self, plan, rank: int, half: Table, peer_bases: Sequence[int]):
        """Transposed off-diagonal rows of a half table -> the stripes of their owner ranks.  Enqueue only."""
        arr = (ctypes.c_uint64 * len(peer_bases))(*[int(p) for p in peer_bases])
        self._sync_stream()
        self._check(self._lib.ottocov_mirror_push(self._ctx, ctypes.byref(plan), int(rank), half._h, arr))

    def mirror_collect(self, plan, rank: int, half: Table, recv_area: int):
        """-> (full Table | None, need_rows): own half rows + the transposed rows received, sorted by key."""
        h, need = ctypes.c_void_p(), ctypes.c_int64()
        self._sync_stream()
        self._check(self._lib.ottocov_mirror_collect(self._ctx, ctypes.byref(plan), int(rank), half._h, int(recv_area),
                                                     ctypes.byref(h), ctypes.byref(need)))
        return (Table(self, h.value) if h.value else None), int(need.value)

    def mirror(self, table: Table, transpose_only: bool = False) -> Table:
        """Half table (rows aid <= aid_next) -> full symmetric table (or only the transposed rows)."""
        h = ctypes.c_void_p()
        self._sync_stream()
        self._check(self._lib.ottocov_table_mirror(self._ctx, table._h, int(transpose_only), ctypes.byref(h)))
        return Table(self, h.value)

    def count_info(self) -> Dict[str, int]:
        ci = _lib.CountInfo()
        self._check(self._lib.ottocov_get_count_info(self._ctx, ctypes.byref(ci)))
        return {k: int(getattr(ci, k)) for k, _ in _lib.CountInfo._fields_}

    # ---- tables ------------------------------------------------------------------------------------------
    def table_from_arrays(self, aid, aid_next, count) -> Table:
        self._sync_stream()
        pa, wa, ka = _as_col(aid, np.int32, "int32")
        pb, wb, kb = _as_col(aid_next, np.int32, "int32")
        if _is_torch(count):
            import torch
            kc = count.to(torch.int32).contiguous()
            pc, wc = kc.data_ptr(), (_lib.DEVICE if kc.is_cuda else _lib.HOST)
        else:
            kc = np.ascontiguousarray(count).astype(np.uint32, copy=False)
            kc = np.ascontiguousarray(kc)
            pc, wc = kc.ctypes.data, _lib.HOST
        if len({wa, wb, wc}) != 1:
            raise ValueError("all three columns must live on the same side")
        h = ctypes.c_void_p()
        self._check(self._lib.ottocov_table_from_arrays(self._ctx, pa, pb, pc, len(ka), wa, ctypes.byref(h)))
        return Table(self, h.value)

    def table_from_packed(self, keys, count, n: Optional[int] = None) -> Table:
        """keys: u64 (aid << 32 | aid_next) as int64 tensor/array; count: u32 as int32 tensor/array."""
        self._sync_stream()
        pk, wk, kk = _as_col(keys, np.uint64, "int64")
        pc, wc, kc = _as_col(count, np.uint32, "int32")
        if wk != wc:
            raise ValueError("keys and count must live on the same side")
        n = len(kk) if n is None else int(n)
        h = ctypes.c_void_p()
        self._check(self._lib.ottocov_table_from_packed(self._ctx, pk, pc, n, wk, ctypes.byref(h)))
        return Table(self, h.value)

    def merge(self, tables: Iterable[Table]) -> Table:
        """groupby(['aid','aid_next']).sum() over the concatenation (count_co_events.py:168)."""
        tabs = list(tables)
        arr = (ctypes.c_void_p * max(len(tabs), 1))(*[t._h for t in tabs])
        h = ctypes.c_void_p()
        self._sync_stream()
        self._check(self._lib.ottocov_table_merge(self._ctx, arr, len(tabs), ctypes.byref(h)))
        return Table(self, h.value)

    def filter(self, table: Table, min_count: int) -> Table:
        h = ctypes.c_void_p()
        self._sync_stream()
        self._check(self._lib.ottocov_table_filter(self._ctx, table._h, int(min_count), ctypes.byref(h)))
        return Table(self, h.value)

    # ---- (4) segmented top-K --------------------------------------------------------------------------------
    def topk(self, table: Table, k: Optional[int] = None, device: bool = False, pinned: bool = False, fetch: bool = True):
        """-> aid_x [A], n_valid [A], aid_y [A, k], cnt [A, k]  (retrieve.py:41-47; canonical ties).
        fetch=False leaves the result in the engine's device buffers (for topk_lookup) and returns the number of aids."""
        k = self.config.TOP_K if k is None else int(k)
        n = ctypes.c_int64()
        self._sync_stream()
        self._check(self._lib.ottocov_table_topk(self._ctx, table._h, k, ctypes.byref(n)))
        self._last_topk_k = k
        A = int(n.value)
        if not fetch:
            return A
        if device:
            import torch
            dev = torch.device("cuda", self.device)
            ax = torch.empty(A, dtype=torch.int32, device=dev); nv = torch.empty(A, dtype=torch.int32, device=dev)
            ay = torch.empty((A, k), dtype=torch.int32, device=dev); ac = torch.empty((A, k), dtype=torch.int32, device=dev)
            ptrs, where = (ax.data_ptr(), nv.data_ptr(), ay.data_ptr(), ac.data_ptr()), _lib.DEVICE
        else:
            if pinned:
                ax = self._pinned("tk_ax", A); nv = self._pinned("tk_nv", A)
                ay = self._pinned("tk_ay", A * k).reshape(A, k); ac = self._pinned("tk_ac", A * k).reshape(A, k)
            else:
                ax = np.empty(A, np.int32); nv = np.empty(A, np.int32)
                ay = np.empty((A, k), np.int32); ac = np.empty((A, k), np.int32)
            ptrs, where = (ax.ctypes.data, nv.ctypes.data, ay.ctypes.data, ac.ctypes.data), _lib.HOST
        self._check(self._lib.ottocov_topk_fetch(self._ctx, *ptrs, A, where))
        return ax, nv, ay, ac

    def topk_lookup(self, aids):
        """Top-K rows of the last topk() result for the given aids (numpy int32) -> n_valid [n], aid_y [n,k],
        cnt [n,k].  The consumer's candidate join (retrieve.py:75-91)."""
        a = np.ascontiguousarray(aids, np.int32)
        n = len(a)
        k = ctypes.c_int64()
        nv = np.empty(n, np.int32)
        kk = self._last_topk_k
        ay = np.empty((n, kk), np.int32); ac = np.empty((n, kk), np.int32)
        self._sync_stream()
        self._check(self._lib.ottocov_topk_lookup(self._ctx, a.ctypes.data, n, _lib.HOST, nv.ctypes.data, ay.ctypes.data,
                                                  ac.ctypes.data))
        return nv, ay, ac

    # ---- derived co-count features (model/retrieve.py:18-63) ------------------------------------------------------
    def count_features(self, aid, aid_next, count, first_n: int) -> Dict[str, np.ndarray]:
        """get_df_count_for_co_event_type on a count table given in FILE order: per-aid top-first_n (count desc,
        aid_next asc) with count_pop / perc_pop / rank / count_rel, all computed on the device.
        -> {'aid', 'aid_next', 'count', 'count_pop' (i16), 'perc_pop' (i16), 'rank' (i16), 'count_rel' (i8)}."""
        pa, wa, ka = _as_col(aid, np.int32, "int32")
        pb, wb, kb = _as_col(aid_next, np.int32, "int32")
        pc, wc, kc = _as_col(count, np.int32, "int32")
        if len({wa, wb, wc}) != 1:
            raise ValueError("all three columns must live on the same side (host or device)")
        n = len(ka)
        # numpy's method='nearest': the row of the ascending order closest to (n - 1) * q (retrieve.py:34 uses
        # polars' quantile(0.9999, 'nearest'))
        qrow = int(np.around((n - 1) * 0.9999)) if n else 0
        rows = ctypes.c_int64()
        self._sync_stream()
        self._check(self._lib.ottocov_count_features(self._ctx, pa, pb, pc, n, wa, int(first_n), qrow, ctypes.byref(rows)))
        m = int(rows.value)
        o = {"aid": np.empty(m, np.int32), "aid_next": np.empty(m, np.int32), "count": np.empty(m, np.int32),
             "count_pop": np.empty(m, np.int16), "perc_pop": np.empty(m, np.int16), "rank": np.empty(m, np.int16),
             "count_rel": np.empty(m, np.int8)}
        self._check(self._lib.ottocov_count_features_fetch(self._ctx, *(o[k].ctypes.data for k in
                                                           ("aid", "aid_next", "count", "count_pop", "perc_pop", "rank", "count_rel")),
                                                           m, _lib.HOST))
        return o

    # ---- popularity inside session clusters (model/count_popularity.py:56-85) ---------------------------------
    POPULARITY_RANK_COLUMNS = ("rank_clicks", "rank_carts", "rank_orders", "rank_clicks_7d", "rank_carts_7d",
                               "rank_orders_7d")

    def count_popularity(self, cluster, aid, ts, type_, ts_recent: int, keep_top_k: int = 20) -> Dict[str, np.ndarray]:
        """groupby([cluster, aid]) with six counts (type x {all time, ts > ts_recent}), ordinal rank per cluster
        (count desc, aid asc; clipped to 999) for each count, rows whose best rank is <= keep_top_k.
        `cluster` is the per-event cluster id (-1 = none).  -> {'aid', 'cluster', rank_* (int16)}, rows ordered
        by (cluster, aid)."""
        pc, wc, kc = _as_col(cluster, np.int32, "int32")
        pa, wa, ka = _as_col(aid, np.int32, "int32")
        pt, wt, kt = _as_col(ts, np.int32, "int32")
        py, wy, ky = _as_col(type_, np.int8, "int8")
        if len({wc, wa, wt, wy}) != 1:
            raise ValueError("all four columns must live on the same side (host or device)")
        n = int(kc.shape[0])
        rows = ctypes.c_int64()
        self._sync_stream()
        self._check(self._lib.ottocov_count_popularity(self._ctx, pc, pa, pt, py, n, wc, int(ts_recent), int(keep_top_k),
                                                       ctypes.byref(rows)))
        m = int(rows.value)
        o_aid = np.empty(m, np.int32); o_cl = np.empty(m, np.int32); o_rank = np.empty((6, m), np.int16)
        self._check(self._lib.ottocov_popularity_fetch(self._ctx, o_aid.ctypes.data, o_cl.ctypes.data, o_rank.ctypes.data,
                                                       m, _lib.HOST))
        out = {"aid": o_aid, "cluster": o_cl}
        for i, name in enumerate(self.POPULARITY_RANK_COLUMNS):
            out[name] = o_rank[i]
        return out

    # ---- multi-GPU support ----------------------------------------------------------------------------------
    def partition(self, table: Table, n_ranks: int, keys_out_ptr: int, count_out_ptr: int) -> List[int]:
        rows = (ctypes.c_int64 * n_ranks)()
        self._sync_stream()
        self._check(self._lib.ottocov_table_partition(self._ctx, table._h, n_ranks, keys_out_ptr, count_out_ptr, rows))
        return [int(x) for x in rows]

    def sort_u64(self, keys_ptr: int, vals_ptr: Optional[int], n: int, lo_bit: int = 0, hi_bit: int = 64):
        self._sync_stream()
        self._check(self._lib.ottocov_sort_u64(self._ctx, keys_ptr, vals_ptr, n, lo_bit, hi_bit))

    # ---- convenience: whole hot path for one co-event kind ------------------------------------------------------
    def count_topk(self, name: str, min_count: Optional[int] = None, k: Optional[int] = None):
        """expand -> reduce -> threshold -> top-K for one name; returns (filtered Table, topk tuple)."""
        thr = self.config.MIN_COUNT_TO_SAVE.get(name, 1) if min_count is None else min_count
        f = self.count(name, min_count=thr)
        return f, self.topk(f, k)
