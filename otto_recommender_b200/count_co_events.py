"""Drop-in for the reference stage ``model/count_co_events.py`` on a B200.

Same entry point and arguments (``python -m model.count_co_events --data_split_alias A --count 1
--merge 1 --merge_train_test 1``; reference :184-190), same directory layout (:194-198, :97-100,
:179) and the same parquet schemas (part files ``aid:int32, aid_next:int32, count:uint32``; merged
files ``count:int32`` ordered by count descending, written through pandas).  The file-level
functions keep the reference's names and signatures:

    count_co_events_all_files(dir_sessions, dir_stats, skip_if_exists=True)      reference :80-100
    concat_files_w_stats(name, dir_stats, files_stats=None)                      reference :103-181

The three frame-level helpers (``self_merge``, ``self_merge_big_df``, ``count_co_events``) exist in
the reference only to build and scan the n^2 joined frame; the engine never materialises it, so
``count_co_events`` here takes the *events* of one part and returns the five count frames, and the
two self-merge helpers are intentionally absent.

Deviations, all documented in SURVEY.md section 0:
  * the three phase flags are parsed as int (in the reference any value given on the command line
    is a str, so ``--count 1`` *disables* the phase, :187-189,201);
  * order among equal counts in the merged files is (aid, aid_next) ascending -- the reference
    leaves it to its hash group-by;
  * the >300 M-row sliced aggregation (:135-166) is reproduced only up to that tie order.

All counting, merging, thresholding and ordering runs on the GPU through libottocov.so; there is no
CPU fallback.
"""
from __future__ import annotations

import argparse
import glob
import logging
import math
import os
import time
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq

from .config import DEFAULT_CONFIG, CoEventConfig
from .engine import Engine, Table

log = logging.getLogger(os.path.basename(__file__))

config: CoEventConfig = DEFAULT_CONFIG
_engine: Optional[Engine] = None


def get_engine() -> Engine:
    """One lazily created engine on the current CUDA device (LOCAL_RANK if set)."""
    global _engine
    if _engine is None:
        _engine = Engine(device=int(os.environ.get("LOCAL_RANK", "0")), config=config)
    return _engine


def set_config(cfg: CoEventConfig):
    global config, _engine
    config = cfg
    if _engine is not None:
        _engine.config = cfg


# ---- part level ------------------------------------------------------------------------------------
def _read_events(file_parquet: str):
    t = pq.read_table(file_parquet, columns=["session", "aid", "ts", "type"])
    return (t["session"].to_numpy(), t["aid"].to_numpy(), t["ts"].to_numpy(), t["type"].to_numpy())


def _counts_frame(table: Table) -> pa.Table:
    a, b, c = table.fetch(order="key")
    return pa.table({"aid": pa.array(a, pa.int32()), "aid_next": pa.array(b, pa.int32()),
                     "count": pa.array(c.astype(np.uint32), pa.uint32())})


def count_co_events(df_events) -> Dict[str, pd.DataFrame]:
    """Five co-event count frames for the events of one part (reference :60-77 + :17-57 fused).

    ``df_events``: pandas DataFrame or pyarrow Table with columns session, aid, ts, type.
    """
    if isinstance(df_events, pa.Table):
        cols = [df_events[c].to_numpy() for c in ("session", "aid", "ts", "type")]
    else:
        cols = [np.asarray(df_events[c]) for c in ("session", "aid", "ts", "type")]
    eng = get_engine()
    info = eng.load_events(*cols)
    log.debug(f"compute_co_events(): input has {info['n_rows_in']:,} rows, {info['n_events']:,} unique events")
    out = {}
    for name in config.MAP_NAME_COUNT_TYPE:
        t = eng.count(name)
        out[name] = _counts_frame(t).to_pandas()
        t.free()
    return out


def count_co_events_all_files(dir_sessions, dir_stats, skip_if_exists=True):
    files_parquet = sorted(glob.glob(f"{dir_sessions}/*.parquet"))
    eng = get_engine()
    for file_parquet in files_parquet:
        stem = Path(file_parquet).stem
        all_exists = all(os.path.exists(f"{dir_stats}/{name_df}/{stem}.parquet")
                         for name_df in config.CO_EVENTS_TO_COUNT)
        if skip_if_exists and all_exists:
            log.debug(f"skipping {stem}.parquet, counts already exist")
            continue
        eng.load_events(*_read_events(file_parquet))          # read + unique (:91-92)
        for name_df in config.MAP_NAME_COUNT_TYPE:            # self-merge + 5 counts (:93-94)
            t = eng.count(name_df)
            file_name_out = f"{dir_stats}/{name_df}/{stem}.parquet"
            os.makedirs(os.path.dirname(file_name_out), exist_ok=True)
            pq.write_table(_counts_frame(t), file_name_out)   # (:97-100)
            t.free()


# ---- merge level -----------------------------------------------------------------------------------
def _read_counts(files: List[str]):
    tabs = [pq.read_table(f) for f in files]
    for t in tabs:
        assert t.column_names == ["aid", "aid_next", "count"], t.column_names     # reference :128
    if not tabs:
        z = np.zeros(0, np.int32)
        return z, z.copy(), np.zeros(0, np.uint32)
    t = pa.concat_tables([x.cast(pa.schema([("aid", pa.int32()), ("aid_next", pa.int32()),
                                            ("count", pa.int64())])) for x in tabs])
    return (t["aid"].to_numpy(), t["aid_next"].to_numpy(), t["count"].to_numpy().astype(np.uint32))


def concat_files_w_stats(name, dir_stats, files_stats=None):
    log.debug(f"merge and aggregate counts for {name}")
    eng = get_engine()
    file_tmp = f"{dir_stats}/tmp/{name}.parquet"     # cache written after a sliced aggregation (:106)
    loaded_from_cache = False
    if os.path.exists(file_tmp):
        log.debug(f"loading cached {file_tmp}")
        aid, aid_next, cnt = _read_counts([file_tmp])
        loaded_from_cache = True
    elif files_stats is not None:
        aid, aid_next, cnt = _read_counts(list(files_stats))
    else:
        aid, aid_next, cnt = _read_counts(sorted(glob.glob(f"{dir_stats}/{name}/*.parquet")))
    n_rows = len(aid)
    log.debug(f"loaded {n_rows:,} rows in total")

    min_in_part = config.MIN_COUNT_IN_PART.get(name, 1)
    lossy = not config.EXACT_MERGE and not loaded_from_cache
    # truncate small counts if table is big (:131-132)
    if lossy and "click_to" in name and n_rows > config.ROWS_TRIGGER_MIN_COUNT_IN_PART:
        keep = cnt >= min_in_part
        aid, aid_next, cnt = aid[keep], aid_next[keep], cnt[keep]
        n_rows = len(aid)

    if lossy and n_rows > config.MAX_ROWS_POLARS_GROUPBY:
        # aggregate by positional slices, truncate each (:135-166)
        rows_part = config.OPTIM_ROWS_POLARS_GROUPBY
        n_parts = math.ceil(n_rows / rows_part)
        max_rows_part = int(config.MAX_ROWS_POLARS_GROUPBY / n_rows * rows_part)
        rows_part = math.ceil(n_rows / n_parts)
        pa_, pb_, pc_ = [], [], []
        for i in range(n_parts):
            s = slice(i * rows_part, (i + 1) * rows_part)
            t = eng.table_from_arrays(aid[s], aid_next[s], cnt[s])          # groupby.sum of the slice
            f = eng.filter(t, min_in_part); t.free()
            a, b, c = f.fetch(order="count_desc", head=max_rows_part); f.free()
            pa_.append(a); pb_.append(b); pc_.append(c.astype(np.uint32))
        aid, aid_next, cnt = np.concatenate(pa_), np.concatenate(pb_), np.concatenate(pc_)
        log.debug(f"{len(aid):,} rows after concatenation of parts")
        os.makedirs(f"{dir_stats}/tmp", exist_ok=True)
        pq.write_table(pa.table({"aid": aid, "aid_next": aid_next, "count": cnt}), file_tmp)

    t = eng.table_from_arrays(aid, aid_next, cnt)                            # groupby.sum (:168)
    log.debug(f"{t.rows:,} rows after aggregation")
    f = eng.filter(t, config.MIN_COUNT_TO_SAVE.get(name, 1)); t.free()      # (:172)
    a, b, c = f.fetch(order="count_desc", head=config.MAX_CO_EVENT_PAIRS_TO_SAVE_DISK)   # (:173-175)
    f.free()
    log.debug(f"{len(a):,} rows after filtering and chopping to first "
              f"{config.MAX_CO_EVENT_PAIRS_TO_SAVE_DISK:,} rows with most counts")
    df = pd.DataFrame({"aid": a, "aid_next": b, "count": c})
    df.to_parquet(f"{dir_stats}/{name}.parquet")                             # via pandas (:179)
    log.debug(f"df saved to {dir_stats}/{name}.parquet")


# ---- fused path (extension): whole population in HBM, no part files ----------------------------------
def count_population(dir_sessions, names=None, min_counts=None) -> Dict[str, Table]:
    """All parts of one population counted in one go: phase 1 + 2 without the per-part files.  The parquet parts
    are handed to the engine as they were read (one host buffer per column per part, nothing concatenated on the
    host); it copies them on a second stream and counts group by group behind the copies
    (``ottocov_count_parts``).  Equal to the phased path whenever the reference's lossy merge steps do not trigger
    (exact mode).  min_counts: per-name thresholds fused into the reduce (default 1 = keep every pair, like the
    part files)."""
    files = sorted(glob.glob(f"{dir_sessions}/*.parquet"))
    names = list(names or config.CO_EVENTS_TO_COUNT)
    parts = [_read_events(f) for f in files]
    eng = get_engine()
    tabs = eng.count_parts(parts, names, list(min_counts) if min_counts is not None else [1] * len(names))
    return dict(zip(names, tabs))


def _layout(alias: str):
    """Directory contract of the stage (reference :194-198)."""
    root = f"{config.DIR_DATA}/{alias}"
    populations = ("train_sessions", "test_sessions")
    sessions = {p: f"{root}-parquet/{p}" for p in populations}
    stats = f"{root}-counts-co-event"
    return populations, sessions, stats


def main(argv=None):
    ap = argparse.ArgumentParser(description="co-event counting on a B200 (drop-in for model.count_co_events)")
    ap.add_argument("--data_split_alias", default="train-test")
    for phase in ("count", "merge", "merge_train_test"):       # reference :187-189, parsed as int here
        ap.add_argument(f"--{phase}", default=1, type=int)
    args = ap.parse_args(argv)
    populations, sessions, stats = _layout(args.data_split_alias)
    t0 = time.time()

    if args.count == 1:                                         # phase 1: one set of part files per input part
        tic = time.time()
        for p in populations:
            count_co_events_all_files(sessions[p], f"{stats}/{p}")
        log.info(f"count - time elapsed: {time.time() - tic:.2f} s")

    if args.merge == 1:                                         # phase 2: per population
        tic = time.time()
        for name in config.CO_EVENTS_TO_COUNT:
            for p in populations:
                concat_files_w_stats(name, f"{stats}/{p}")
        log.info(f"merge - time elapsed: {time.time() - tic:.2f} s")

    if args.merge_train_test == 1:                              # phase 3: train + test
        for name in config.CO_EVENTS_TO_COUNT:
            concat_files_w_stats(name=name, dir_stats=stats,
                                 files_stats=[f"{stats}/{p}/{name}.parquet" for p in populations])

    log.info(f"count_co_events - total time elapsed: {time.time() - t0:.2f} s")


if __name__ == "__main__":
    logging.basicConfig(format="%(asctime)s - %(name)s - %(levelname)s - %(message)s", level=logging.DEBUG)
    main()
