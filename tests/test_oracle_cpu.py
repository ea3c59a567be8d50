"""CPU suite (-m "not gpu"): the oracles against the golden vectors, against each other, and the
algebraic invariants of SURVEY.md App. B.2."""
import json
import os

import numpy as np
import pytest

from conftest import small_events
from oracle import c_oracle, ref_restatement as rr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "b1_events.json")


def _gold():
    g = json.load(open(GOLD))
    rows = np.array(g["rows"])
    return rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], g["expected"]


def _dict(a, b, c):
    return {(int(x), int(y)): int(z) for x, y, z in zip(a, b, c)}


def test_golden_b1_c_oracle():
    s, a, t, y, exp = _gold()
    for name, rows in exp.items():
        oa, ob, oc, emitted, nded = c_oracle.count_name(s, a, t, y, name)
        assert _dict(oa, ob, oc) == {(r[0], r[1]): r[2] for r in rows}, name
        assert emitted == sum(r[2] for r in rows)
        assert nded == 14          # 15 rows, one exact duplicate


def test_golden_b1_restatement():
    s, a, t, y, exp = _gold()
    res = rr.count_events_all_names(s, a, t, y)
    for name, rows in exp.items():
        assert rr.table_to_dict(res[name]) == {(r[0], r[1]): r[2] for r in rows}, name


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracles_agree(seed):
    s, a, t, y = small_events(seed, n_sessions=200)
    res = rr.count_events_all_names(s, a, t, y, n_sessions_in_part=37)   # slicing must not matter
    for name in c_oracle.NAMES:
        oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, name)
        assert rr.table_to_dict(res[name]) == _dict(oa, ob, oc), name


def test_invariants_b2():
    s, a, t, y = small_events(5, n_sessions=250)
    for name in ("click_to_click", "cart_to_cart", "buy_to_buy"):
        d = _dict(*c_oracle.count_name(s, a, t, y, name)[:3])
        for (x, z), c in d.items():
            assert d[(z, x)] == c                      # symmetric kinds
            if x == z:
                assert c % 2 == 0
    cb = _dict(*c_oracle.count_name(s, a, t, y, "click_to_cart_or_buy")[:3])
    c1 = _dict(*c_oracle.count(s, a, t, y, 0, 0b010, 86400)[:3])
    c2 = _dict(*c_oracle.count(s, a, t, y, 0, 0b100, 86400)[:3])
    keys = set(c1) | set(c2)
    assert set(cb) == keys
    for k in keys:
        assert cb[k] == c1.get(k, 0) + c2.get(k, 0)
    # row order of the input does not matter
    p = np.random.default_rng(0).permutation(len(s))
    d1 = _dict(*c_oracle.count_name(s, a, t, y, "click_to_click")[:3])
    d2 = _dict(*c_oracle.count_name(s[p], a[p], t[p], y[p], "click_to_click")[:3])
    assert d1 == d2


def test_edge_cases_oracle():
    z = np.zeros(0, np.int32)
    oa, ob, oc, em, nd = c_oracle.count_name(z, z, z, z.astype(np.int8), "click_to_click")
    assert len(oa) == 0 and em == 0 and nd == 0
    # singletons only: no pairs
    s = np.arange(10); a = np.arange(10); t = np.full(10, 1_660_000_000); y = np.zeros(10)
    assert c_oracle.count_name(s, a, t, y, "click_to_click")[3] == 0
    # one session, all events at one timestamp, same aid but different types
    s = np.zeros(3); a = np.full(3, 7); t = np.full(3, 100); y = np.array([0, 1, 2])
    assert _dict(*c_oracle.count_name(s, a, t, y, "click_to_cart_or_buy")[:3]) == {(7, 7): 2}
    assert c_oracle.count_name(s, a, t, y, "click_to_click")[3] == 0


def test_merge_and_topn_restatement_vs_numpy():
    s, a, t, y = small_events(9, n_sessions=400, n_aids=30)
    half = s < np.median(s)
    parts = []
    for m in (half, ~half):
        parts.append(rr.count_events_all_names(s[m], a[m], t[m], y[m])["click_to_click"])
    merged = rr.merge_counts("click_to_click", parts, exact=True, min_count_to_save=3)
    whole = c_oracle.count_name(s, a, t, y, "click_to_click")
    ka, kb, kc = c_oracle.merge_tables([whole[:3]], min_count=3)
    sa, sb, sc = c_oracle.sort_count_desc(ka, kb, kc)
    assert np.array_equal(merged["aid"].to_numpy(), sa)
    assert np.array_equal(merged["aid_next"].to_numpy(), sb)
    assert np.array_equal(merged["count"].to_numpy(), sc)
    top = rr.top_n_per_aid(merged, 5)
    ta, tb, tc, tr = c_oracle.top_n(ka, kb, kc, 5)
    assert np.array_equal(top["aid"].to_numpy(), ta)
    assert np.array_equal(top["aid_next"].to_numpy(), tb)
    assert np.array_equal(top["count"].to_numpy(), tc)
    assert np.array_equal(top["rank"].to_numpy(), tr)


# ---- reference-run pin --------------------------------------------------------------------------------------
def _ref_fixtures():
    import glob
    return sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.json")))


def check_against_reference_fixture(path, count_fn, topn_fn):
    """Shared by the CPU (oracles) and GPU (CUDA path) tests: tables equal the reference's own output; the top-N is
    compared tie-aware (the reference keeps file order among equal counts)."""
    g = json.load(open(path))
    rows = np.array(g["rows"])
    s, a, t, y = (rows[:, i] for i in range(4))
    for kind, want in g["counts"].items():
        got = count_fn(s, a, t, y, kind)
        assert got == {(r[0], r[1]): r[2] for r in want}, (path, kind)
        kept = topn_fn(s, a, t, y, kind, g["topn"][kind]["first_n"])
        for aid, counts in g["topn"][kind]["kept_counts_per_aid"].items():
            assert sorted(kept.get(int(aid), []), reverse=True) == counts, (path, kind, aid)


def test_reference_run_fixtures():
    """tests/golden/ref_*.json are written by tools/gen_reference_fixtures.py from the UNMODIFIED reference functions
    under polars.  polars is not installable in the build container, so the files may be absent: then this test
    skips LOUDLY and the oracle stays pinned only by hand-verified vectors (parity unpinned by the reference)."""
    files = _ref_fixtures()
    if not files:
        pytest.skip("PARITY UNPINNED: no tests/golden/ref_*.json -- run tools/gen_reference_fixtures.py where polars "
                    "(~0.15-0.16) is available to pin the oracle against the reference's own output")

    def count_fn(s, a, t, y, kind):
        oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, kind)
        d = _dict(oa, ob, oc)
        assert d == rr.table_to_dict(rr.count_events_all_names(s, a, t, y)[kind])
        return d

    def topn_fn(s, a, t, y, kind, first_n):
        oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, kind)
        ta, tb, tc, _ = c_oracle.top_n(oa, ob, oc, first_n)
        out = {}
        for x, c in zip(ta.tolist(), tc.tolist()):
            out.setdefault(x, []).append(c)
        return out

    for f in files:
        check_against_reference_fixture(f, count_fn, topn_fn)
