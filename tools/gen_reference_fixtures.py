#!/usr/bin/env python
"""Generate reference-run golden fixtures: tests/golden/ref_*.json.

Runs the UNMODIFIED reference functions -- /root/reference/model/count_co_events.py::self_merge,
self_merge_big_df, count_co_events and the top-N block of model/retrieve.py::get_df_count_for_co_event_type --
on small seeded inputs and writes inputs + outputs as JSON.  tests/test_oracle_cpu.py::test_reference_run_fixtures
and tests/test_gpu_parity.py::test_reference_run_fixtures_gpu consume the files when they exist; that is what moves
the oracle from "parity unpinned" to pinned by the reference itself.

Needs polars of the reference's vintage (API: DataFrame.groupby, with_columns, .rank(reverse=True); polars
~0.15-0.16, January 2023 -- requirements.txt:19 does not pin it).  polars is NOT installable in the build
container (no network, no wheel), so this script could not be run there; it is committed so that anyone with that
environment can produce the fixtures:

    pip install 'polars>=0.15,<0.17' tqdm
    python tools/gen_reference_fixtures.py --reference /path/to/otto-recommender [--out tests/golden]

The reference module is imported FROM the reference tree as it lies there; only its `config` import (which has side
effects: creates artifacts/, opens a log file, config.py:13-27) is satisfied by the reference's own config.py
loaded with the working directory pointed at a temporary folder.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np


def load_reference(ref_root: str):
    try:
        import polars as pl  # noqa: F401
    except ImportError as e:
        raise SystemExit("polars is required to run the reference (pip install 'polars>=0.15,<0.17'): %s" % e)
    ref_root = os.path.abspath(ref_root)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="ref_fixture_")
    os.chdir(tmp)                                    # config.py creates artifacts/ and logs in the cwd
    try:
        sys.path.insert(0, ref_root)                 # `import config` inside the module resolves to the reference's
        spec = importlib.util.spec_from_file_location("ref_count_co_events",
                                                      os.path.join(ref_root, "model", "count_co_events.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                 # model/__init__.py (IPython / plotly imports) is NOT executed
    finally:
        os.chdir(cwd)
    import config as ref_config
    return mod, ref_config


def small_events(seed: int, n_sessions: int, n_aids: int, max_len: int):
    """Same generator as tests/conftest.py::small_events (dense, window edges, duplicates), sorted by (session, ts)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len + 1, n_sessions)
    sess = np.repeat(np.arange(n_sessions) * 7 + 3, lens)
    n = len(sess)
    start = np.repeat(rng.integers(1_660_000_000, 1_660_000_000 + 5 * 86400, n_sessions), lens)
    gaps = rng.choice([0, 1, 30, 600, 43200, 43201, 86400, 86401, 100_000], n) * rng.integers(0, 2, n)
    ts = start + rng.integers(0, 200_000, n) // 4 + gaps
    aid = rng.integers(0, n_aids, n)
    typ = rng.choice([0, 0, 0, 0, 1, 1, 2], n)
    pick = rng.integers(0, n, int(n * 0.05))
    sess, ts, aid, typ = (np.concatenate([x, x[pick]]) for x in (sess, ts, aid, typ))
    order = np.lexsort((ts, sess))
    return sess[order].astype(np.int32), aid[order].astype(np.int32), ts[order].astype(np.int32), typ[order].astype(np.int8)


def reference_topn(pl, df_count, first_n: int):
    """The per-aid block of model/retrieve.py:41-51, verbatim expressions."""
    df = df_count.sort(["aid"])
    df = df.with_columns([pl.col("count").rank("ordinal", reverse=True).over("aid").cast(pl.Int16).alias("rank")])
    return df.filter(pl.col("rank") <= first_n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    args = ap.parse_args()
    mod, cfg = load_reference(args.reference)
    import polars as pl
    os.makedirs(args.out, exist_ok=True)
    b1 = json.load(open(os.path.join(args.out, "b1_events.json")))
    cases = {"b1": tuple(np.array(b1["rows"])[:, i] for i in range(4))}
    for seed, (ns, na, ml) in {11: (60, 12, 12), 12: (200, 40, 25), 13: (400, 300, 40)}.items():
        cases[f"seed{seed}"] = small_events(seed, ns, na, ml)
    for name, (s, a, t, y) in cases.items():
        df = pl.DataFrame({"session": np.asarray(s, np.int32), "aid": np.asarray(a, np.int32),
                           "ts": np.asarray(t, np.int32), "type": np.asarray(y, np.int8)})
        df = df.unique()                                                   # count_co_events.py:92
        merged = mod.self_merge_big_df(df, n_sessions_in_part=37)          # :41-57 (slicing must not matter)
        counts = mod.count_co_events(merged)                               # :60-77
        out = {"polars_version": pl.__version__, "rows": np.stack([s, a, t, y], 1).astype(int).tolist(),
               "n_events_after_unique": int(df.shape[0]), "counts": {}, "topn": {}}
        for kind, d in counts.items():
            d = d.sort(["aid", "aid_next"])
            out["counts"][kind] = [[int(x) for x in r] for r in zip(d["aid"], d["aid_next"], d["count"])]
            first_n = cfg.RETRIEVAL_FIRST_N_CO_COUNTS[kind]
            top = reference_topn(pl, d.with_columns([pl.col("count").cast(pl.Int32)]), first_n)
            # ties are resolved by file order in the reference: record only what is tie-independent
            out["topn"][kind] = {"first_n": first_n,
                                 "kept_counts_per_aid": {str(int(k)): sorted((int(c) for c in g["count"]), reverse=True)
                                                         for k, g in ((k, top.filter(pl.col("aid") == k)) for k in top["aid"].unique())}}
        path = os.path.join(args.out, f"ref_{name}.json")
        json.dump(out, open(path, "w"))
        print("wrote", path, {k: len(v) for k, v in out["counts"].items()})


if __name__ == "__main__":
    main()
