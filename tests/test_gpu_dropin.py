"""GPU suite (-m gpu): the drop-in surface -- `python -m model.count_co_events` directory contract,
part/merged parquet schemas, two-level thresholds, resume, and the consumer's per-aid top-N -- against
the dataframe-shaped restatement of the reference (oracle/ref_restatement.py)."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from conftest import small_events
from oracle import ref_restatement as rr
from otto_recommender_b200 import count_co_events as cce
from otto_recommender_b200 import retrieve
from otto_recommender_b200.config import CoEventConfig

pytestmark = pytest.mark.gpu


def _write_parts(dir_sessions, s, a, t, y, n_parts):
    os.makedirs(dir_sessions, exist_ok=True)
    ids = np.unique(s)
    cuts = np.linspace(0, len(ids), n_parts + 1).astype(int)
    tables = []
    for i in range(n_parts):
        lo, hi = ids[cuts[i]], ids[cuts[i + 1] - 1]
        m = (s >= lo) & (s <= hi)
        tab = pa.table({"session": pa.array(s[m], pa.int32()), "aid": pa.array(a[m], pa.int32()),
                        "ts": pa.array(t[m], pa.int32()), "type": pa.array(y[m], pa.int8())})
        pq.write_table(tab, f"{dir_sessions}/{cuts[i]:07d}_{cuts[i + 1]:07d}.parquet")
        tables.append(tab)
    return tables


@pytest.fixture()
def data_dir(tmp_path, engine):
    cfg = CoEventConfig(DIR_DATA=str(tmp_path))
    cce.set_config(cfg)
    cce._engine = engine
    yield tmp_path
    cce._engine = None


def test_cli_three_phases_match_reference(data_dir):
    alias = "tt"
    s, a, t, y = small_events(41, n_sessions=4000, n_aids=25, max_len=30, shuffle=True, dup_frac=0.02)
    is_test = (s % 5 == 0)
    parts = {}
    for pop, m, k in (("train_sessions", ~is_test, 3), ("test_sessions", is_test, 2)):
        parts[pop] = _write_parts(f"{data_dir}/{alias}-parquet/{pop}", s[m], a[m], t[m], y[m], k)

    cce.main(["--data_split_alias", alias])

    stats = f"{data_dir}/{alias}-counts-co-event"
    for name in rr.CO_EVENTS_TO_COUNT:
        merged = {}
        for pop, tabs in parts.items():
            per_part = [rr.count_part(tab)[name] for tab in tabs]
            # phase 1: one file per part, schema (aid i32, aid_next i32, count u32), any row order
            files = sorted(os.listdir(f"{stats}/{pop}/{name}"))
            assert len(files) == len(tabs)
            for f, want in zip(files, per_part):
                got = pq.read_table(f"{stats}/{pop}/{name}/{f}")
                assert got.schema.names == ["aid", "aid_next", "count"]
                assert got.schema.field("count").type == pa.uint32() and got.schema.field("aid").type == pa.int32()
                assert rr.table_to_dict(got) == rr.table_to_dict(want)
            # phase 2: per population, thresholded, count desc
            merged[pop] = rr.merge_counts(name, per_part)
            got = pq.read_table(f"{stats}/{pop}/{name}.parquet")
            assert got.schema.names == ["aid", "aid_next", "count"] and got.schema.field("count").type == pa.int32()
            for c in ("aid", "aid_next", "count"):
                assert np.array_equal(got[c].to_numpy(), merged[pop][c].to_numpy()), (name, pop, c)
        # phase 3: train + test; thresholds were applied per population first (two-level)
        want = rr.merge_counts(name, [merged["train_sessions"], merged["test_sessions"]])
        got = pq.read_table(f"{stats}/{name}.parquet")
        for c in ("aid", "aid_next", "count"):
            assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), (name, c)
        assert want.num_rows > 0 or name in ("buy_to_buy", "cart_to_buy")
        # consumer: per-aid top-N with derived columns (retrieve.py:18-63)
        if want.num_rows:
            df = retrieve.get_df_count_for_co_event_type(name, stats)
            top = rr.top_n_per_aid(want, rr.RETRIEVAL_FIRST_N[name])
            assert np.array_equal(df["aid"].to_numpy(), top["aid"].to_numpy())
            assert np.array_equal(df["aid_next"].to_numpy(), top["aid_next"].to_numpy())
            assert np.array_equal(df[f"{name}_count"].to_numpy(), top["count"].to_numpy())
            assert np.array_equal(df[f"{name}_rank"].to_numpy(), top["rank"].to_numpy())
            cnt = top["count"].to_numpy().astype(np.float64)
            mx = np.maximum.reduceat(cnt, np.r_[0, np.flatnonzero(np.diff(top["aid"].to_numpy())) + 1])
            seg = np.r_[0, np.cumsum(np.diff(top["aid"].to_numpy()) != 0)]
            assert np.array_equal(df[f"{name}_count_rel"].to_numpy(), (cnt / mx[seg] * 100).astype(np.int8))
            assert list(df.columns) == ["aid", "aid_next", f"{name}_count", f"{name}_count_pop", f"{name}_perc_pop",
                                        f"{name}_rank", f"{name}_count_rel"]
            feat = rr.count_features(want, name)                       # all seven columns of retrieve.py:18-63
            for col, arr in feat.items():
                assert np.array_equal(df[col].to_numpy(), arr), (name, col)
            # candidate join of session aids with the top-N rows (retrieve.py:75-91)
            q = pa.table({"aid": pa.array(np.r_[a[::7], 10**6], pa.int32())}).to_pandas()
            pairs = retrieve.get_pairs_co_event_type(q, df, 0)
            want_pairs = q.drop_duplicates().merge(df[["aid", "aid_next"]], on="aid", how="inner")
            key = lambda d: set(zip(d["aid"].tolist(), d["aid_next"].tolist()))
            assert key(pairs) == key(want_pairs) and len(pairs) == len(want_pairs)


def test_lossy_merge_branches_with_small_triggers(data_dir, engine):
    """concat_files_w_stats' row-count-triggered steps (count_co_events.py:131-132, :135-166), reached by
    shrinking the triggers: drop count < 2 for click_to_* once the concatenation is large, then aggregate by
    positional slices and truncate each; plus the final head(N).  Tie order is canonical on both sides."""
    alias = "ls"
    s, a, t, y = small_events(47, n_sessions=3000, n_aids=40, max_len=25)
    tabs = _write_parts(f"{data_dir}/{alias}-parquet/train_sessions", s, a, t, y, 4)
    out = f"{data_dir}/{alias}-counts-co-event/train_sessions"
    cce.count_co_events_all_files(f"{data_dir}/{alias}-parquet/train_sessions", out)
    for name in ("click_to_click", "click_to_cart_or_buy", "cart_to_cart"):
        # the sliced step cuts the concatenation by POSITION, so both sides must see the part rows in the
        # same order: use the part files phase 1 wrote (their content is checked against the oracle elsewhere)
        for tab, f in zip(tabs, sorted(os.listdir(f"{out}/{name}"))):
            assert rr.table_to_dict(pq.read_table(f"{out}/{name}/{f}")) == rr.table_to_dict(rr.count_part(tab)[name])
        per_part = [pq.read_table(f"{out}/{name}/{f}") for f in sorted(os.listdir(f"{out}/{name}"))]
        n_rows = sum(p.num_rows for p in per_part)
        kw = dict(rows_trigger_min_in_part=n_rows // 2, max_rows_groupby=n_rows // 3, optim_rows_groupby=n_rows // 5 + 1,
                  max_pairs_to_save=max(n_rows // 50, 5))
        cfg = CoEventConfig(DIR_DATA=str(data_dir), ROWS_TRIGGER_MIN_COUNT_IN_PART=kw["rows_trigger_min_in_part"],
                            MAX_ROWS_POLARS_GROUPBY=kw["max_rows_groupby"], OPTIM_ROWS_POLARS_GROUPBY=kw["optim_rows_groupby"],
                            MAX_CO_EVENT_PAIRS_TO_SAVE_DISK=kw["max_pairs_to_save"],
                            MIN_COUNT_TO_SAVE={"click_to_click": 3, "click_to_cart_or_buy": 2, "cart_to_cart": 2})
        cce.set_config(cfg)
        want = rr.merge_counts(name, per_part, min_count_to_save=cfg.MIN_COUNT_TO_SAVE[name], **kw)
        cce.concat_files_w_stats(name, out)
        got = pq.read_table(f"{out}/{name}.parquet")
        assert got.num_rows == want.num_rows and got.num_rows <= kw["max_pairs_to_save"]
        for c in ("aid", "aid_next", "count"):
            assert np.array_equal(got[c].to_numpy(), want[c].to_numpy()), (name, c)
        # the sliced aggregation leaves its cache behind (:165-166) and a second call re-uses it (:106-111)
        assert os.path.exists(f"{out}/tmp/{name}.parquet")
        cce.concat_files_w_stats(name, out)
        again = pq.read_table(f"{out}/{name}.parquet")
        assert again.num_rows > 0
    cce.set_config(CoEventConfig(DIR_DATA=str(data_dir)))


def test_resume_skips_existing_parts(data_dir):
    alias = "rs"
    s, a, t, y = small_events(43, n_sessions=300, n_aids=20)
    _write_parts(f"{data_dir}/{alias}-parquet/train_sessions", s, a, t, y, 2)
    out = f"{data_dir}/{alias}-counts-co-event/train_sessions"
    cce.count_co_events_all_files(f"{data_dir}/{alias}-parquet/train_sessions", out)
    f = sorted(os.listdir(f"{out}/click_to_click"))[0]
    before = os.path.getmtime(f"{out}/click_to_click/{f}")
    cce.count_co_events_all_files(f"{data_dir}/{alias}-parquet/train_sessions", out)            # all exist: skipped
    assert os.path.getmtime(f"{out}/click_to_click/{f}") == before
    os.remove(f"{out}/buy_to_buy/{f}")
    cce.count_co_events_all_files(f"{data_dir}/{alias}-parquet/train_sessions", out)            # one missing: recomputed
    assert os.path.exists(f"{out}/buy_to_buy/{f}")
    assert os.path.getmtime(f"{out}/click_to_click/{f}") > before


def test_frame_level_count_co_events(data_dir):
    s, a, t, y = small_events(44, n_sessions=200, n_aids=15)
    tab = rr.events_table(s, a, t, y)
    got = cce.count_co_events(tab)
    want = rr.count_part(tab)
    for name in rr.CO_EVENTS_TO_COUNT:
        g = got[name]
        assert list(g.columns) == ["aid", "aid_next", "count"]
        assert {(int(x), int(z)): int(c) for x, z, c in zip(g["aid"], g["aid_next"], g["count"])} == \
            rr.table_to_dict(want[name])


def test_fused_population_equals_phased(data_dir):
    alias = "fp"
    s, a, t, y = small_events(45, n_sessions=1500, n_aids=30)
    tabs = _write_parts(f"{data_dir}/{alias}-parquet/train_sessions", s, a, t, y, 3)
    fused = cce.count_population(f"{data_dir}/{alias}-parquet/train_sessions")
    for name, tab in fused.items():
        want = rr.merge_counts(name, [rr.count_part(x)[name] for x in tabs], exact=True, min_count_to_save=1)
        assert tab.to_dict() == rr.table_to_dict(want), name
