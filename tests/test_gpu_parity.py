"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracles on
identical inputs.  Integer work throughout => every comparison is bit-exact equality."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import small_events
from oracle import c_oracle
from oracle import ref_restatement as rr
from otto_recommender_b200 import OttocovError
from otto_recommender_b200.dist import hash_dest
from otto_recommender_b200.retrieve import topn_long
from otto_recommender_b200.synth import SynthSpec, generate_numpy

pytestmark = pytest.mark.gpu
NAMES = list(c_oracle.NAMES)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "b1_events.json")


def _dict(a, b, c):
    return {(int(x), int(y)): int(z) for x, y, z in zip(a, b, c)}


def _check_all_names(engine, s, a, t, y, names=NAMES, **kw):
    info = engine.load_events(s, a, t, y)
    for name in names:
        oa, ob, oc, emitted, nded = c_oracle.count_name(s, a, t, y, name)
        tab = engine.count(name, **kw)
        ga, gb, gc = tab.fetch()
        ci = engine.count_info()
        assert info["n_events"] == nded
        assert ci["n_pairs"] == emitted, name
        assert np.array_equal(ga, oa) and np.array_equal(gb, ob), name          # sorted by (aid, aid_next)
        assert np.array_equal(gc.astype(np.uint32), oc), name
        assert tab.total() == emitted
        tab.free()


# ---- radix sort ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 33, 4095, 4096, 4097, 100_003, 3_000_001])
def test_sort_keys_only(engine, n):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    keys = torch.randint(0, 2**62, (n,), generator=g, device="cuda", dtype=torch.int64)
    want = torch.sort(keys).values
    engine.sort_u64(keys.data_ptr(), None, n, 0, 62)
    torch.cuda.synchronize()
    assert torch.equal(keys, want)


@pytest.mark.parametrize("lo,hi", [(0, 7), (3, 24), (32, 53), (10, 42)])
def test_sort_pairs_partial_bits_is_stable(engine, lo, hi):
    n = 700_001
    g = torch.Generator(device="cuda"); g.manual_seed(lo * 100 + hi)
    keys = torch.randint(0, 2**62, (n,), generator=g, device="cuda", dtype=torch.int64)
    vals = torch.arange(n, device="cuda", dtype=torch.int32)
    field = (keys >> lo) & ((1 << (hi - lo)) - 1)
    order = torch.sort(field, stable=True).indices
    want_k, want_v = keys[order], vals[order]
    engine.sort_u64(keys.data_ptr(), vals.data_ptr(), n, lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(keys, want_k)
    assert torch.equal(vals, want_v)


def test_sort_skewed_digits(engine):
    # all keys share most digits (one bin takes everything in several passes)
    n = 1_000_000
    keys = (torch.arange(n, device="cuda", dtype=torch.int64) % 3) << 40
    keys = keys[torch.randperm(n, device="cuda")].contiguous()
    want = torch.sort(keys).values
    engine.sort_u64(keys.data_ptr(), None, n, 0, 48)
    torch.cuda.synchronize()
    assert torch.equal(keys, want)


# ---- golden vector + randomized differential tests -------------------------------------------------------
def test_golden_b1(engine):
    g = json.load(open(GOLD))
    rows = np.array(g["rows"])
    engine.load_events(rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3])
    for name, exp in g["expected"].items():
        tab = engine.count(name)
        assert tab.to_dict() == {(r[0], r[1]): r[2] for r in exp}, name
        tab.free()
    assert engine.events_info()["n_events"] == 14


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("shuffle", [False, True])
def test_random_vs_oracle(engine, seed, shuffle):
    s, a, t, y = small_events(seed, n_sessions=400, shuffle=shuffle)
    _check_all_names(engine, s, a, t, y)
    assert engine.events_info()["was_sorted"] == (0 if shuffle else 1)


def test_arbitrary_specs_vs_oracle(engine):
    """Any (source type, next-type set, window), not just the five configured kinds."""
    s, a, t, y = small_events(61, n_sessions=500, n_aids=60, max_len=35)
    info = engine.load_events(s, a, t, y)
    rng = np.random.default_rng(5)
    specs = [(0, [0, 1, 2], 86400), (2, [0], 50_000), (1, [0, 2], 1), (0, [0], 0), (2, [1, 2], 43_201), (1, [1], 86_400 * 3)]
    specs += [(int(rng.integers(0, 3)), sorted(set(rng.integers(0, 3, rng.integers(1, 4)).tolist())),
               int(rng.choice([7, 600, 3600, 43_200, 86_399, 86_400]))) for _ in range(6)]
    for this, nxt, w in specs:
        mask = sum(1 << k for k in nxt)
        oa, ob, oc, emitted, _ = c_oracle.count(s, a, t, y, this, mask, min(w, 86400))
        for mc in (1, 2):
            tab = engine.count(type_this=this, next_types=nxt, window=w, min_count=mc)
            ga, gb, gc = tab.fetch()
            keep = oc >= mc
            assert engine.count_info()["n_pairs"] == emitted, (this, nxt, w)
            assert np.array_equal(ga, oa[keep]) and np.array_equal(gb, ob[keep]) and np.array_equal(gc.astype(np.uint32), oc[keep]), (this, nxt, w, mc)


def test_long_sessions_and_chunking(engine):
    # a few sessions of ~2000 events inside one window: quadratic tail, tiles span many records
    s, a, t, y = small_events(7, n_sessions=6, n_aids=400, max_len=2000, span=40_000)
    _check_all_names(engine, s, a, t, y, names=["click_to_click", "click_to_cart_or_buy", "cart_to_buy"])
    # same result when the output space is cut into many pair-budget chunks
    _check_all_names(engine, s, a, t, y, names=["click_to_click", "click_to_cart_or_buy"], pair_budget=3 * 2048)
    assert engine.count_info()["n_chunks"] > 1


def test_edge_cases(engine):
    z = np.zeros(0, np.int32)
    info = engine.load_events(z, z, z, z.astype(np.int8))
    assert info["n_events"] == 0
    for name in NAMES:
        tab = engine.count(name)
        assert tab.rows == 0
        ax, nv, ay, ac = engine.topk(tab, 20)
        assert len(ax) == 0
        tab.free()
    # singletons only
    s = np.arange(10); a = np.arange(10); t = np.full(10, 1_660_000_000); y = np.zeros(10, np.int8)
    engine.load_events(s, a, t, y)
    assert engine.count("click_to_click").rows == 0
    # a single event
    engine.load_events(s[:1], a[:1], t[:1], y[:1])
    assert engine.count("click_to_click").rows == 0
    # same (session, ts), same aid, three types; plus exact duplicates of each
    s = np.zeros(6); a = np.full(6, 7); t = np.full(6, 100); y = np.array([0, 1, 2, 0, 1, 2])
    _check_all_names(engine, s, a, t, y)
    # extreme timestamps / ids (int32 edges): windows must saturate, not wrap
    s = np.array([-2**31, -2**31, 2**31 - 1, 2**31 - 1, 5, 5])
    t = np.array([-2**31, -2**31 + 86400, 2**31 - 1, 2**31 - 1 - 43200, 0, 43201])
    a = np.array([1, 2, 3, 4, 2**31 - 1, 0]); y = np.zeros(6)
    _check_all_names(engine, s, a, t, y, names=["click_to_click"])
    # only carts / only orders
    s, a, t, y = small_events(21, n_sessions=50)
    _check_all_names(engine, s, a, t, np.ones_like(y))
    _check_all_names(engine, s, a, t, np.full_like(y, 2))


def test_invariants_on_gpu(engine):
    s, a, t, y = small_events(13, n_sessions=500, n_aids=80)
    engine.load_events(s, a, t, y)
    for name in ("click_to_click", "cart_to_cart", "buy_to_buy"):
        d = engine.count(name).to_dict()
        for (x, z), c in d.items():
            assert d[(z, x)] == c
    cb = engine.count("click_to_cart_or_buy")
    c1 = engine.count(type_this=0, next_types=[1], window=86400)
    c2 = engine.count(type_this=0, next_types=[2], window=86400)
    m = engine.merge([c1, c2])
    assert m.to_dict() == cb.to_dict()
    # window larger than 24 h is clamped by the +-24 h pre-filter (count_co_events.py:33-36)
    w1 = engine.count(type_this=0, next_types=[0], window=86400).to_dict()
    w2 = engine.count(type_this=0, next_types=[0], window=10 * 86400).to_dict()
    assert w1 == w2


def test_fused_threshold_equals_filter(engine):
    s, a, t, y = small_events(23, n_sessions=800, n_aids=60)
    engine.load_events(s, a, t, y)
    for name in ("click_to_click", "click_to_cart_or_buy"):
        full = engine.count(name)
        for thr in (1, 2, 3, 7, 1000):
            want = engine.filter(full, thr)
            got = engine.count(name, min_count=thr)
            assert engine.count_info()["n_pairs"] == full.total()
            for x, z in zip(got.fetch(), want.fetch()):
                assert np.array_equal(x, z)
            # chunked: thresholds must still apply to complete sums
            got2 = engine.count(name, min_count=thr, pair_budget=4096)
            for x, z in zip(got2.fetch(), want.fetch()):
                assert np.array_equal(x, z)


@pytest.mark.parametrize("name", ["click_to_click", "cart_to_cart", "buy_to_buy"])
def test_symmetric_shortcut_is_exact(engine, name):
    s, a, t, y = small_events(31, n_sessions=700, n_aids=70, max_len=60)
    engine.load_events(s, a, t, y)
    oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
    for kw in ({}, {"pair_budget": 2048}):
        full = engine.count(name, symmetric=False, **kw)
        half = engine.count(name, symmetric=True, **kw)          # canonical half pairs + mirror
        assert engine.count_info()["n_pairs"] == emitted
        for x, z, w in zip(half.fetch(), full.fetch(), (oa, ob, oc)):
            assert np.array_equal(x, z) and np.array_equal(x.astype(w.dtype), w)
        for thr in (2, 4, 9):                                     # diagonal rows count both orders
            ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=thr)
            for sym in (True, False, None):
                got = engine.count(name, min_count=thr, symmetric=sym, **kw).fetch()
                assert np.array_equal(got[0], ka) and np.array_equal(got[1], kb) and np.array_equal(got[2], kc)
    # an asymmetric kind ignores the request
    cb = engine.count("click_to_cart_or_buy", symmetric=True)
    assert cb.to_dict() == _dict(*c_oracle.count_name(s, a, t, y, "click_to_cart_or_buy")[:3])


def test_long_runs_span_tiles(engine):
    # one pair repeated far beyond a reduce tile (2048 keys): carries must cross many tiles
    n = 30_000
    s = np.zeros(n); a = np.where(np.arange(n) % 2 == 0, 5, 9); t = np.full(n, 1000) + np.arange(n) % 7
    y = np.zeros(n)
    # events are distinct only by ts (7 values) -> after dedup 14 events; use distinct ts instead
    t = 1_660_000_000 + np.arange(n) % 40_000
    _check_all_names(engine, s, a, t, y, names=["click_to_click"])
    tab = engine.table_from_arrays(np.full(50_000, 3), np.full(50_000, 4), np.ones(50_000))
    assert tab.to_dict() == {(3, 4): 50_000}


def test_device_resident_inputs(engine):
    s, a, t, y = small_events(17, n_sessions=300, shuffle=True)
    ts_ = [torch.from_numpy(x).cuda() for x in (s, a, t, y)]
    engine.load_events(*ts_)
    oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, "click_to_click")
    tab = engine.count("click_to_click")
    ga, gb, gc = tab.fetch(device=True)
    assert np.array_equal(ga.cpu().numpy(), oa) and np.array_equal(gb.cpu().numpy(), ob)
    assert np.array_equal(gc.cpu().numpy().astype(np.uint32), oc)


# ---- tables: merge / filter / order / top-K / partition --------------------------------------------------------
def test_merge_filter_order(engine):
    s, a, t, y = small_events(9, n_sessions=600, n_aids=40)
    half = s < np.median(s)
    tabs = []
    for m in (half, ~half):
        engine.load_events(s[m], a[m], t[m], y[m])
        tabs.append(engine.count("click_to_click"))
    merged = engine.merge(tabs)
    whole = c_oracle.count_name(s, a, t, y, "click_to_click")
    assert merged.to_dict() == _dict(*whole[:3])
    for thr in (1, 2, 5, 50, 10**6):
        f = engine.filter(merged, thr)
        ka, kb, kc = c_oracle.merge_tables([whole[:3]], min_count=thr)
        ga, gb, gc = f.fetch()
        assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc)
        sa, sb, sc = c_oracle.sort_count_desc(ka, kb, kc)
        ga, gb, gc = f.fetch(order="count_desc")
        assert np.array_equal(ga, sa) and np.array_equal(gb, sb) and np.array_equal(gc, sc)
        ga, gb, gc = f.fetch(order="count_desc", head=7)
        assert np.array_equal(ga, sa[:7]) and np.array_equal(gc, sc[:7])
    # round trip through arrays (unsorted, with duplicate keys that must be summed)
    ka, kb, kc = whole[:3]
    p = np.random.default_rng(1).permutation(len(ka))
    t2 = engine.table_from_arrays(np.concatenate([ka[p], ka[:100]]), np.concatenate([kb[p], kb[:100]]),
                                  np.concatenate([kc[p], kc[:100]]))
    want = _dict(ka, kb, kc)
    for i in range(100):
        want[(int(ka[i]), int(kb[i]))] += int(kc[i])
    assert t2.to_dict() == want


@pytest.mark.parametrize("k", [1, 10, 20, 32])
def test_topk_vs_oracle(engine, k):
    # head-heavy aids: segments from 1 row to several thousand rows, lots of equal counts (ties)
    rng = np.random.default_rng(k)
    n = 400_000
    ax = np.minimum(rng.zipf(1.3, n), 5000) - 1
    ay = rng.integers(0, 3000, n)
    cnt = rng.integers(1, 6, n)
    tab = engine.table_from_arrays(ax, ay, cnt)
    ka, kb, kc = c_oracle.merge_tables([(ax, ay, cnt)])
    assert tab.rows == len(ka)
    ta, tb, tc, tr = c_oracle.top_n(ka, kb, kc, k)
    gx, nv, gy, gc = engine.topk(tab, k)
    la, lb, lc, lr = topn_long(gx, nv, gy, gc)
    assert np.array_equal(la, ta) and np.array_equal(lb, tb) and np.array_equal(lc, tc) and np.array_equal(lr, tr)
    assert np.all(gy[np.arange(k)[None, :] >= nv[:, None]] == -1)


def test_topk_lookup_candidates(engine):
    rng = np.random.default_rng(3)
    ax = rng.integers(0, 2000, 60_000) * 3            # only multiples of 3 exist
    ay = rng.integers(0, 500, 60_000); cnt = rng.integers(1, 9, 60_000)
    tab = engine.table_from_arrays(ax, ay, cnt)
    gx, nv, gy, gc = engine.topk(tab, 10)
    q = rng.integers(-5, 6100, 5000).astype(np.int32)
    qnv, qy, qc = engine.topk_lookup(q)
    row = {int(a): i for i, a in enumerate(gx)}
    for i, a in enumerate(q):
        if int(a) in row:
            r = row[int(a)]
            assert qnv[i] == nv[r] and np.array_equal(qy[i], gy[r]) and np.array_equal(qc[i], gc[r])
        else:
            assert qnv[i] == 0 and np.all(qy[i] == -1) and np.all(qc[i] == 0)


def test_count_features_on_device(engine):
    """ottocov_count_features == the restatement of get_df_count_for_co_event_type (retrieve.py:18-63), all seven
    columns: (1) a canonical file (count desc, aid asc, aid_next asc) with heavy ties; (2) distinct counts in an
    arbitrary FILE order -- perc_pop must follow the file row, the quantile the sorted counts."""
    import pyarrow as pa
    rng = np.random.default_rng(9)
    for case in ("canonical_ties", "arbitrary_order"):
        n = 60_000
        key = rng.choice(400 * 400, n, replace=False)
        aid, nxt = (key // 400).astype(np.int32), (key % 400).astype(np.int32) * 3
        if case == "canonical_ties":
            cnt = np.minimum(rng.geometric(0.05, n), 300).astype(np.int32)
            o = np.lexsort((nxt, aid, -cnt.astype(np.int64)))
        else:
            cnt = (rng.permutation(n) + 5).astype(np.int32)
            o = rng.permutation(n)
        aid, nxt, cnt = aid[o], nxt[o], cnt[o]
        tab = pa.table({"aid": aid, "aid_next": nxt, "count": cnt})
        for first_n in (1, 10, 20, 32):
            want = rr.count_features(tab, "click_to_click", first_n)
            got = engine.count_features(aid, nxt, cnt, first_n)
            ren = {"aid": "aid", "aid_next": "aid_next", "click_to_click_count": "count", "click_to_click_count_pop": "count_pop",
                   "click_to_click_perc_pop": "perc_pop", "click_to_click_rank": "rank", "click_to_click_count_rel": "count_rel"}
            for col, arr in want.items():
                assert np.array_equal(got[ren[col]], arr), (case, first_n, col)
    e = engine.count_features(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), 10)
    assert all(len(v) == 0 for v in e.values())
    one = engine.count_features(np.array([7]), np.array([9]), np.array([4]), 10)       # q == min: the division guard
    w1 = rr.count_features(pa.table({"aid": np.array([7], np.int32), "aid_next": np.array([9], np.int32),
                                     "count": np.array([4], np.int32)}), "click_to_click", 10)
    assert one["count_pop"].tolist() == w1["click_to_click_count_pop"].tolist() and one["perc_pop"].tolist() == [10000]


def test_partition_by_hash(engine):
    s, a, t, y = small_events(3, n_sessions=500, n_aids=500)
    engine.load_events(s, a, t, y)
    tab = engine.count("click_to_click")
    ka, kb, kc = tab.fetch()
    for R in (1, 2, 3, 8):
        keys = torch.empty(tab.rows, dtype=torch.int64, device="cuda")
        cnts = torch.empty(tab.rows, dtype=torch.int32, device="cuda")
        rows = engine.partition(tab, R, keys.data_ptr(), cnts.data_ptr())
        torch.cuda.synchronize()
        dest = hash_dest(ka, R)
        assert rows == np.bincount(dest, minlength=R).tolist()
        order = np.argsort(dest, kind="stable")
        want = (ka.astype(np.int64) << 32 | kb.astype(np.int64))[order]
        assert np.array_equal(keys.cpu().numpy(), want)
        assert np.array_equal(cnts.cpu().numpy(), kc[order])
        # emulate the exchange on one GPU: every "rank" re-reduces what it would receive
        off = np.concatenate([[0], np.cumsum(rows)])
        got = {}
        for r in range(R):
            sh = engine.table_from_packed(keys[off[r]:off[r + 1]].clone(), cnts[off[r]:off[r + 1]].clone())
            got.update(sh.to_dict())
        assert got == _dict(ka, kb, kc)


@pytest.mark.parametrize("name,mc", [("click_to_click", 1), ("click_to_click", 3), ("click_to_cart_or_buy", 2),
                                     ("cart_to_cart", 1)])
def test_exchange_first_emulated_ranks(engine, name, mc):
    """The multi-GPU flow on one GPU: keys grouped by destination rank, each 'rank' reduces its slice,
    symmetric halves are mirrored; the union must equal the plain single-GPU table."""
    s, a, t, y = small_events(51, n_sessions=900, n_aids=300, max_len=50)
    info = engine.load_events(s, a, t, y)
    want = engine.count(name, min_count=mc, symmetric=False).to_dict()
    for R in (1, 2, 5, 8):
        n_keys, sym = engine.expand_prepare(name, min_count=mc)
        assert sym == (name in ("click_to_click", "cart_to_cart"))
        buf_a = torch.empty(max(n_keys, 1), dtype=torch.int64, device="cuda")
        buf_b = torch.empty(max(n_keys, 1), dtype=torch.int64, device="cuda")
        grouped, rows = engine.expand_run(R, buf_a, buf_b)
        assert sum(rows) == n_keys
        off = np.concatenate([[0], np.cumsum(rows)])
        got = {}
        for r in range(R):
            keys_r = grouped[off[r]:off[r + 1]].clone()
            half = engine.reduce_pairs(keys_r, keys_r.numel(), info["aid_bits"], mc, symmetric=sym, strip_dest=R > 1)
            ha, hb, hc = half.fetch()
            assert np.all(hash_dest(ha, R) == r)
            pieces = [half]
            if sym:
                pieces.append(engine.mirror(half, transpose_only=True))
            for p in pieces:
                for k, v in p.to_dict().items():
                    assert k not in got
                    got[k] = v
        assert got == want, (name, mc, R)


def _session_shards(s, a, t, y, R):
    """contiguous session ranges balanced by the pair-work proxy, as bench.py / dist.py shard them"""
    from otto_recommender_b200.dist import shard_bounds
    order = np.lexsort((t, s))
    s, a, t, y = s[order], a[order], t[order], y[order]
    us, lens = np.unique(s, return_counts=True)
    b = shard_bounds(lens, R)
    cuts = np.concatenate([[0], np.cumsum(lens)])[b]
    return [tuple(x[cuts[r]:cuts[r + 1]] for x in (s, a, t, y)) for r in range(R)]


@pytest.mark.parametrize("R", [2, 3, 8])
def test_scatter_exchange_emulated_vs_oracle(engine, R):
    """The default multi-GPU flow (expansion fused with the first bucket pass and the exchange: every key is stored
    straight into a stripe of its owner's receive area; mirrored rows of symmetric kinds pushed the same way) with
    the R ranks played by one engine.  All five kinds against the plain-C oracle run on the WHOLE event set: every
    row on its owner rank, the union equal to the oracle's thresholded table."""
    from otto_recommender_b200.dist import count_scatter_emulated
    for seed, kw in ((55, dict(n_sessions=900, n_aids=300, max_len=50)), (56, dict(n_sessions=2500, n_aids=40, max_len=30))):
        s, a, t, y = small_events(seed, **kw)
        info = engine.load_events(s, a, t, y)
        shards = _session_shards(s, a, t, y, R)
        for name in NAMES:
            oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
            for mc in (1, 3):
                ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
                tabs, regrows = count_scatter_emulated(engine, shards, name, mc, info["aid_bits"])
                got = {}
                for r, tab in enumerate(tabs):
                    ga, gb, gc = tab.fetch()
                    assert np.all(hash_dest(ga, R) == r), "row on the wrong rank"
                    assert np.all(np.diff(ga.astype(np.int64) << 32 | gb) > 0)            # sorted, distinct
                    for k, v in zip(zip(ga.tolist(), gb.tolist()), gc.tolist()):
                        assert k not in got
                        got[k] = v
                    tab.free()
                assert got == _dict(ka, kb, kc), (name, mc, R, seed)


def test_scatter_exchange_regrows_consistently(engine):
    """Stripes far too small: every rank reads the same published need, the plan grows, the step repeats; the mirror
    stripes likewise.  Same table as the oracle."""
    from otto_recommender_b200.dist import count_scatter_emulated
    s, a, t, y = small_events(57, n_sessions=1500, n_aids=200, max_len=40)
    info = engine.load_events(s, a, t, y)
    R = 4
    shards = _session_shards(s, a, t, y, R)
    for name in ("click_to_click", "click_to_cart_or_buy"):
        oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, name)
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=2)
        tabs, regrows = count_scatter_emulated(engine, shards, name, 2, info["aid_bits"], stripe_cap=64, mirror_cap=8)
        assert regrows >= (2 if name == "click_to_click" else 1)
        got = {}
        for tab in tabs:
            got.update(tab.to_dict())
        assert got == _dict(ka, kb, kc), name


def test_push_keys_emulated_peers(engine):
    """ottocov_push_keys with every 'peer' receive buffer living at an offset of one local tensor: the
    result must be exactly the grouped output of the two-buffer expand_run."""
    s, a, t, y = small_events(52, n_sessions=1200, n_aids=400, max_len=40)
    engine.load_events(s, a, t, y)
    for R in (2, 3, 8):
        n_keys, _ = engine.expand_prepare("click_to_click", min_count=2)
        a0 = torch.empty(n_keys, dtype=torch.int64, device="cuda"); b0 = torch.empty_like(a0)
        grouped, rows = engine.expand_run(R, a0, b0)
        want = grouped.clone()
        n2, _ = engine.expand_prepare("click_to_click", min_count=2)
        assert n2 == n_keys
        a1 = torch.empty(n_keys, dtype=torch.int64, device="cuda")
        keys, rows2 = engine.expand_run(R, a1, None)
        assert rows2 == rows
        recv = torch.full((n_keys + 8 * R,), -1, dtype=torch.int64, device="cuda")     # 8-key guard gaps
        off = np.concatenate([[0], np.cumsum(rows)])
        ptrs = [recv.data_ptr() + 8 * (int(off[d]) + 8 * d) for d in range(R)]
        engine.push_keys(keys, n_keys, ptrs)
        torch.cuda.synchronize()
        for d in range(R):
            lo = int(off[d]) + 8 * d
            assert torch.equal(recv[lo:lo + rows[d]], want[int(off[d]):int(off[d + 1])])
            assert bool((recv[lo + rows[d]:lo + rows[d] + 8] == -1).all())               # nothing spilled past the run


def test_errors_are_loud(engine):
    s, a, t, y = small_events(1, n_sessions=20)
    bad = y.copy(); bad[3] = 5
    with pytest.raises(OttocovError) as e:
        engine.load_events(s, a, t, bad)
    assert e.value.code == -3
    neg = a.copy(); neg[0] = -4
    with pytest.raises(OttocovError) as e:
        engine.load_events(s, neg, t, y)
    assert e.value.code == -3
    with pytest.raises(OttocovError) as e:      # failed load leaves no events behind
        engine.count("click_to_click")
    assert e.value.code == -4
    engine.load_events(s, a, t, y)
    tab = engine.count("click_to_click")
    with pytest.raises(OttocovError) as e:
        engine.topk(tab, 33)
    assert e.value.code == -2


# ---- bucketed hash reduce (hash_reduce.cu) == full sort + run-length reduce == oracle --------------------------
@pytest.mark.parametrize("name", NAMES)
def test_hash_reduce_vs_oracle(engine, name):
    s, a, t, y = small_events(41, n_sessions=900, n_aids=90, max_len=50)
    engine.load_events(s, a, t, y)
    oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
    for mc in (1, 2, 5, 100_000):
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
        for sym in (None, True, False):
            for kw in ({}, {"pair_budget": 4096}):
                got = engine.count(name, min_count=mc, symmetric=sym, hashed=True, **kw)
                assert engine.count_info()["n_pairs"] == emitted
                ga, gb, gc = got.fetch()
                assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc), (name, mc, sym, kw)
                got.free()


def test_hash_reduce_wide_keys(engine):
    """aids of 27 bits: 54-bit keys do not fit the packed table word (42-bit tag), the wide variant runs."""
    s, a, t, y = small_events(43, n_sessions=700, n_aids=90, max_len=40)
    a = a * 1_000_003
    info = engine.load_events(s, a, t, y)
    assert info["aid_bits"] == 27
    for name in ("click_to_click", "click_to_cart_or_buy"):
        oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
        for mc in (1, 3):
            ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
            ga, gb, gc = engine.count(name, min_count=mc, hashed=True).fetch()
            assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc), (name, mc)


def test_hash_reduce_many_buckets(engine):
    """Enough keys for several bucket passes and thousands of reduce tiles; hashed == sorted, row for row."""
    d = generate_numpy(SynthSpec(n_sessions=150_000, seed=11))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    engine.load_events(s, a, t, y)
    oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, "click_to_click")
    for name in ("click_to_click", "click_to_cart_or_buy"):
        for mc in (1, 2, 3):
            want = engine.count(name, min_count=mc, hashed=False)
            p_sort = engine.count_info()["sort_passes"]
            got = engine.count(name, min_count=mc, hashed=True)
            ci = engine.count_info()
            assert 1 <= ci["sort_passes"] < p_sort            # bucket bits only: fewer distribution passes
            for x, z in zip(got.fetch(device=True), want.fetch(device=True)):
                assert torch.equal(x, z)
            if name == "click_to_click":
                ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
                ga, gb, gc = got.fetch()
                assert ci["n_pairs"] == emitted
                assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc)
            got.free(); want.free()


def test_hash_reduce_hot_pairs(engine):
    """A pair repeated far beyond a tile: its bucket's tail is streamed by one CTA (warp-aggregated adds);
    a bucket of many millions of keys raises the fallback flag and the sort path takes over.  Same table."""
    for n, expect_fallback in ((2_000, False), (9_000, True)):
        s = np.zeros(n); a = np.where(np.arange(n) % 2 == 0, 5, 9); y = np.zeros(n)
        t = 1_660_000_000 + np.arange(n) % 40_000
        # plus ordinary sessions so that ordinary buckets surround the hot ones
        s2, a2, t2, y2 = small_events(5, n_sessions=300, n_aids=50)
        S = np.concatenate([s, s2 + 1]); A = np.concatenate([a, a2]); T = np.concatenate([t, t2]); Y = np.concatenate([y, y2])
        engine.load_events(S, A, T, Y)
        want = engine.count("click_to_click", min_count=3, hashed=False)
        p_sort = engine.count_info()["sort_passes"]
        got = engine.count("click_to_click", min_count=3, hashed=True)
        fell_back = engine.count_info()["sort_passes"] > p_sort
        assert fell_back == expect_fallback, engine.count_info()
        assert got.to_dict() == want.to_dict()
        d = got.to_dict()
        assert d[(5, 9)] == d[(9, 5)] and d[(5, 9)] >= (n // 2) ** 2


@pytest.mark.parametrize("mc", [1, 3])
def test_reduce_pairs_hashed(engine, mc):
    """ottocov_reduce_pairs (the receive side of the multi-GPU exchange) with either reduce strategy."""
    s, a, t, y = small_events(53, n_sessions=900, n_aids=300, max_len=50)
    info = engine.load_events(s, a, t, y)
    for name in ("click_to_click", "click_to_cart_or_buy"):
        for R in (1, 4):
            tabs = {}
            for hashed in (False, True, None):
                n_keys, sym = engine.expand_prepare(name, min_count=mc)
                buf_a = torch.empty(max(n_keys, 1), dtype=torch.int64, device="cuda")
                buf_b = torch.empty(max(n_keys, 1), dtype=torch.int64, device="cuda")
                grouped, rows = engine.expand_run(R, buf_a, buf_b)
                off = np.concatenate([[0], np.cumsum(rows)])
                got = {}
                for r in range(R):
                    keys_r = grouped[off[r]:off[r + 1]].clone()
                    half = engine.reduce_pairs(keys_r, keys_r.numel(), info["aid_bits"], mc, symmetric=sym,
                                               strip_dest=R > 1, hashed=hashed)
                    ha, hb, hc = half.fetch()
                    assert np.all(np.diff(ha.astype(np.int64) << 32 | hb) > 0)       # sorted, distinct
                    got.update(half.to_dict())
                tabs[hashed] = got
            assert tabs[True] == tabs[False] == tabs[None], (name, R, mc)


def _fused_cases(engine, seed=13, n_sessions=40_000):
    """All five kinds on an OTTO-shaped slice big enough for >= 2 bucket passes of click_to_click and
    click_to_cart_or_buy (the first pass then runs inside the expansion), against the plain-C oracle."""
    d = generate_numpy(SynthSpec(n_sessions=n_sessions, seed=seed))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    engine.load_events(s, a, t, y)
    fused_seen = 0
    big_seen = _fused_cases.big_seen = [0]      # of those, counted by the whole-bucket kernel (count_info "fused" == 2)
    for name in NAMES:
        oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
        for mc in (1, 2, 3):
            ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
            for sym in (None, False):
                for budget in (None, max(emitted // 7, 5000)):          # digit-range chunks when the budget is small
                    got = engine.count(name, min_count=mc, symmetric=sym, hashed=True, pair_budget=budget)
                    ci = engine.count_info()
                    assert ci["n_pairs"] == emitted
                    fused_seen += 1 if ci["fused"] else 0
                    big_seen[0] += 1 if ci["fused"] == 2 else 0
                    if ci["fused"] and budget is not None:
                        assert ci["n_chunks"] >= 3, ci
                    ga, gb, gc = got.fetch()
                    assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc), (name, mc, sym, budget)
                    got.free()
    return fused_seen


def test_fused_first_pass_vs_oracle(engine):
    assert _fused_cases(engine) >= 12          # click_to_click and click_to_cart_or_buy take the fused path


def test_fused_first_pass_overflow_and_unfused_children():
    """(1) OTTOCOV_FUSE_SLACK_PCT=-60: every digit region is too small, the expansion raises the overflow flag and the
    chunk is re-run with regions sized by the fill counters; (2) OTTOCOV_NO_FUSED_PASS=1: the unfused bucket passes.
    Both knobs are read once per process, so the cases run in child processes; same tables as the oracle."""
    import subprocess
    import sys
    if any(os.environ.get(k) for k in ("OTTOCOV_FUSE_SLACK_PCT", "OTTOCOV_NO_FUSED_PASS", "OTTOCOV_NO_FUSED_LOADER", "OTTOCOV_HRB")):
        pytest.skip("already inside the child run")
    # (3) OTTOCOV_NO_FUSED_LOADER=1: the loader's general path (separate statistics pass) on inputs the fused
    # validate+dedup+split pass would otherwise take.
    for extra, sel in (({"OTTOCOV_FUSE_SLACK_PCT": "-60"}, "fused_child or hash_reduce_hot or hash_reduce_many"),
                       ({"OTTOCOV_NO_FUSED_PASS": "1"}, "fused_child or hash_reduce_hot or hash_reduce_many"),
                       ({"OTTOCOV_NO_FUSED_LOADER": "1"}, "golden_b1 or random_vs_oracle or edge_cases or errors_are_loud"),
                       # (4) the whole-bucket reduce (one CTA per bucket, keys staged in shared memory, 2^ROUND_BITS table rounds), which
                       # by default only engages at sizes where it saves a pass (> 33 M keys): forced on small inputs with
                       # small / medium buckets; the hot-pair test then streams a bucket that does not fit the staging area
                       ({"OTTOCOV_HRB": "2", "OTTOCOV_HRB_AVG": "64"}, "fused_child or hash_reduce_hot or hash_reduce_many"),
                       ({"OTTOCOV_HRB": "2", "OTTOCOV_HRB_AVG": "2000"}, "fused_child or hash_reduce_many"),
                       ({"OTTOCOV_HRB": "2", "OTTOCOV_HRB_AVG": "2000", "OTTOCOV_HRB_ROUND_BITS": "0"}, "fused_child"),
                       ({"OTTOCOV_HRB": "0"}, "fused_child")):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                            "-k", sel],
                           env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, str(extra) + r.stdout[-3000:] + r.stderr[-2000:]


def test_fused_child(engine):
    if os.environ.get("OTTOCOV_FUSE_SLACK_PCT"):
        assert _fused_cases(engine, seed=14, n_sessions=30_000) >= 12
    elif os.environ.get("OTTOCOV_NO_FUSED_PASS"):
        assert _fused_cases(engine, seed=14, n_sessions=30_000) == 0
    elif os.environ.get("OTTOCOV_HRB"):
        assert _fused_cases(engine, seed=14, n_sessions=30_000) >= 12
        assert (_fused_cases.big_seen[0] >= 6) == (os.environ["OTTOCOV_HRB"] == "2"), _fused_cases.big_seen
    else:
        pytest.skip("runs inside test_fused_first_pass_overflow_and_unfused_children")


def test_asymmetric_time_range(engine):
    """config.MIN_TIME_TO_NEXT / MAX_TIME_TO_NEXT other than the reference's +-24 h (count_co_events.py:33-36 filters
    MIN <= ts_next - ts <= MAX before the per-kind |dt| <= W): e.g. MIN = 0 means "the next event cannot precede this
    one".  Counts are no longer symmetric; the general window kernel serves these configs."""
    from otto_recommender_b200 import Engine
    from otto_recommender_b200.config import CoEventConfig
    s, a, t, y = small_events(61, n_sessions=500, n_aids=40, max_len=40)
    for lo, hi in ((0, 86400), (1, 86400), (-86400, -1), (-3600, 7200), (100, 50), (-86400, 86400)):
        eng = Engine(0, config=CoEventConfig(MIN_TIME_TO_NEXT=lo, MAX_TIME_TO_NEXT=hi))
        try:
            eng.load_events(s, a, t, y)
            for name, (th, mask, w) in c_oracle.NAMES.items():
                oa, ob, oc, emitted, _ = c_oracle.count(s, a, t, y, th, mask, w, dt_min=lo, dt_max=hi)
                ga, gb, gc = eng.count(name).fetch()
                assert eng.count_info()["n_pairs"] == emitted, (name, lo, hi)
                assert np.array_equal(ga, oa) and np.array_equal(gb, ob) and np.array_equal(gc.astype(np.uint32), oc), (name, lo, hi)
                ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=2)
                ha, hb, hc = eng.count(name, min_count=2, hashed=True).fetch()
                assert np.array_equal(ha, ka) and np.array_equal(hb, kb) and np.array_equal(hc, kc), (name, lo, hi)
        finally:
            eng.close()


def test_pipelined_host_load(engine):
    """Host columns of >= 4 M rows are copied in chunks on a second stream and validated / de-duplicated / split
    chunk by chunk behind the copies (one scan chained over the launches).  Same arrays, same info and same
    tables as the single-shot load of device-resident columns; unsorted or invalid input falls back."""
    d = generate_numpy(SynthSpec(n_sessions=300_000, seed=21))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    assert len(s) >= 4 << 20
    # sprinkle exact duplicates (also across chunk boundaries: every 1/12 of the rows) and keep (session, ts) order
    rng = np.random.default_rng(3)
    dup = np.sort(np.concatenate([rng.integers(0, len(s), 50_000), (len(s) * np.arange(1, 12) // 12) & ~31]))
    idx = np.sort(np.concatenate([np.arange(len(s)), dup]))
    s, a, t, y = s[idx], a[idx], t[idx], y[idx]
    dev = [torch.from_numpy(x).cuda() for x in (s, a, t, y)]
    info_d = engine.load_events(*dev)
    want = {n: engine.count(n, min_count=2, hashed=False).fetch() for n in ("click_to_click", "click_to_cart_or_buy", "cart_to_buy")}
    info_h = engine.load_events(s, a, t, y)                         # host arrays: the pipelined path
    assert info_h == info_d and info_h["was_sorted"] == 1 and info_h["n_events"] < info_h["n_rows_in"]
    for n, w in want.items():
        for x, z in zip(engine.count(n, min_count=2, hashed=False).fetch(), w):
            assert np.array_equal(x, z), n
    # unsorted host input: the optimistic split is dropped, the general path sorts first
    p = rng.permutation(len(s))
    info_u = engine.load_events(s[p], a[p], t[p], y[p])
    assert info_u["was_sorted"] == 0 and info_u["n_events"] == info_d["n_events"] and info_u["n_by_type"] == info_d["n_by_type"]
    for x, z in zip(engine.count("click_to_cart_or_buy", min_count=2).fetch(), want["click_to_cart_or_buy"]):
        assert np.array_equal(x, z)
    # invalid input is still reported
    bad = y.copy(); bad[len(bad) // 2] = 7
    with pytest.raises(OttocovError) as e:
        engine.load_events(s, a, t, bad)
    assert e.value.code == -3
    neg = a.copy(); neg[-5] = -1
    with pytest.raises(OttocovError) as e:
        engine.load_events(s, neg, t, y)
    assert e.value.code == -3


def test_count_parts_streamed(engine):
    """ottocov_count_parts: the parts of one population handed over as separate host buffers, copied on a second stream
    and counted group by group behind the copies (expansion fused with the first bucket pass per group; the remaining
    passes + hash reduce once over the regions of all groups).  Same tables as load_events(concatenation) + count, for
    every kind, thresholded or not; rows inside a part may come in any order; empty parts are fine."""
    from otto_recommender_b200 import Engine
    d = generate_numpy(SynthSpec(n_sessions=60_000, seed=17))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    names = list(NAMES)
    mcs = [3, 2, 1, 2, 1]
    info = engine.load_events(s, a, t, y)
    want = [engine.count(n, min_count=mc).fetch() for n, mc in zip(names, mcs)]
    oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, "click_to_click")
    ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=3)
    assert np.array_equal(want[0][0], ka) and np.array_equal(want[0][2], kc)
    rng = np.random.default_rng(5)
    for n_parts in (1, 7, 23):
        parts = Engine.split_at_sessions(s, a, t, y, n_parts)
        assert sum(len(p[0]) for p in parts) == len(s)
        assert all(p[0][0] != q[0][-1] for p, q in zip(parts[1:], parts[:-1]))        # cut at session boundaries
        if n_parts == 7:                                                              # any row order inside a part; an empty part
            parts = [tuple(x[o] for x in p) for p in parts for o in [rng.permutation(len(p[0]))]]
            parts.insert(3, tuple(x[:0] for x in parts[0]))
        tabs = engine.count_parts(parts, names, mcs)
        ci = engine.count_info()
        assert ci["fused"] == 1
        # the session column crosses as (id, first row) runs unless the part's rows are shuffled (then it is copied raw)
        if n_parts != 7:
            assert 9 * len(s) < ci["h2d_bytes"] < 11 * len(s), ci
        else:
            assert ci["h2d_bytes"] == 13 * len(s), ci
        got_info = engine.events_info()
        for k in ("n_rows_in", "n_events", "n_by_type", "aid_bits", "aid_max", "session_min", "session_max", "ts_min", "ts_max"):
            assert got_info[k] == info[k], k
        for n, tab, w in zip(names, tabs, want):
            for x, z in zip(tab.fetch(), w):
                assert np.array_equal(x, z), (n, n_parts)
            tab.free()
        with pytest.raises(OttocovError) as e:          # the events of the parts are gone: count needs a new load
            engine.count("click_to_click")
        assert e.value.code == -4
    # invalid data inside a part is reported, nothing is left behind
    bad = y.copy(); bad[len(bad) // 3] = 9
    with pytest.raises(OttocovError) as e:
        engine.count_parts(Engine.split_at_sessions(s, a, t, bad, 5), ["click_to_click"], [2])
    assert e.value.code == -3
    tiny = small_events(3, n_sessions=30)               # too few keys for a fused pass: the plain path serves the call
    o = np.lexsort((tiny[2], tiny[0]))
    tiny = tuple(x[o] for x in tiny)
    engine.load_events(*tiny)
    w = engine.count("click_to_cart_or_buy", min_count=1).fetch()
    g = engine.count_parts(Engine.split_at_sessions(*tiny, 3), ["click_to_cart_or_buy"], [1])[0].fetch()
    assert all(np.array_equal(x, z) for x, z in zip(g, w))


def test_hash_reduce_packed_relative_tags():
    """Keys of more than 42 bits (the 4x-scale config: 23-bit aids, 46-bit keys) use the packed table word with
    tags RELATIVE to the tile's first bucket, which needs >= 2^(kb - 34) buckets -- hundreds of millions of keys at
    the default bucket size.  With OTTOCOV_HR_AVG=16 (16 keys per bucket) a million keys are enough; the knob is
    read once per process, so the check runs in a child process."""
    import subprocess
    import sys
    if os.environ.get("OTTOCOV_HR_AVG"):
        pytest.skip("already inside the child run")
    env = dict(os.environ, OTTOCOV_HR_AVG="16")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "wide_keys_small_buckets_child or hash_reduce_vs_oracle"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_wide_keys_small_buckets_child(engine):
    if os.environ.get("OTTOCOV_HR_AVG") != "16":
        pytest.skip("runs inside test_hash_reduce_packed_relative_tags")
    s, a, t, y = small_events(47, n_sessions=4000, n_aids=3000, max_len=50)
    a = a * 2500                                   # 23-bit aids -> 46-bit keys
    info = engine.load_events(s, a, t, y)
    assert info["aid_bits"] == 23
    for name in ("click_to_click", "click_to_cart_or_buy"):
        oa, ob, oc, emitted, _ = c_oracle.count_name(s, a, t, y, name)
        for mc in (1, 2):
            ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=mc)
            for sym in (True, False):
                tab = engine.count(name, min_count=mc, symmetric=sym, hashed=True)
                ci = engine.count_info()
                ga, gb, gc = tab.fetch()
                assert np.array_equal(ga, ka) and np.array_equal(gb, kb) and np.array_equal(gc, kc), (name, mc, sym)
                # bucket passes only (no fall-back to the full 46-bit sort) when there are enough keys for >= 2^12 buckets
                if ci["n_pairs"] >= 400_000:
                    assert ci["sort_passes"] <= 3, ci


# ---- BASELINE config 1: 100k-session synthetic slice, full pair table + top-20 ----------------------------------
def test_config1_100k_sessions(engine):
    d = generate_numpy(SynthSpec(n_sessions=100_000, seed=42))
    s, a, t, y = d["session"], d["aid"], d["ts"], d["type"]
    info = engine.load_events(s, a, t, y)
    assert info["was_sorted"] == 1
    for name in NAMES:
        oa, ob, oc, emitted, nded = c_oracle.count_name(s, a, t, y, name)
        tab = engine.count(name)
        ga, gb, gc = tab.fetch()
        assert engine.count_info()["n_pairs"] == emitted and info["n_events"] == nded
        assert np.array_equal(ga, oa) and np.array_equal(gb, ob) and np.array_equal(gc.astype(np.uint32), oc), name
        # threshold 2 + top-20 (the full-size thresholds 10/5 leave almost nothing on one part)
        f = engine.filter(tab, 2)
        ka, kb, kc = c_oracle.merge_tables([(oa, ob, oc)], min_count=2)
        ta, tb, tc, tr = c_oracle.top_n(ka, kb, kc, 20)
        la, lb, lc, lr = topn_long(*engine.topk(f, 20))
        assert np.array_equal(la, ta) and np.array_equal(lb, tb) and np.array_equal(lc, tc), name
    # shuffled rows give the same tables (row order of the input is irrelevant)
    p = np.random.default_rng(0).permutation(len(s))
    engine.load_events(s[p], a[p], t[p], y[p])
    assert engine.events_info()["was_sorted"] == 0
    oa, ob, oc, _, _ = c_oracle.count_name(s, a, t, y, "click_to_cart_or_buy")
    ga, gb, gc = engine.count("click_to_cart_or_buy").fetch()
    assert np.array_equal(ga, oa) and np.array_equal(gb, ob) and np.array_equal(gc.astype(np.uint32), oc)


# ---- size-independent properties at a larger scale (oracle too slow there) -----------------------------------------
def test_properties_at_scale(engine):
    from otto_recommender_b200.synth import generate
    d = generate(SynthSpec(n_sessions=1_500_000, seed=7), "cuda")
    info = engine.load_events(d["session"], d["aid"], d["ts"], d["type"])
    assert sum(info["n_by_type"]) == info["n_events"] <= info["n_rows_in"]
    cc = engine.count("click_to_click")
    ci = engine.count_info()
    assert cc.total() == ci["n_pairs"] and cc.rows == ci["n_unique"]
    # symmetric kind: the transposed table is the same table
    ka, kb, kc = cc.fetch(device=True)
    tr = engine.table_from_arrays(kb, ka, kc)
    a2, b2, c2 = tr.fetch(device=True)
    assert torch.equal(a2, ka) and torch.equal(b2, kb) and torch.equal(c2, kc)
    # chunked == unchunked
    cc2 = engine.count("click_to_click", pair_budget=ci["n_pairs"] // 5 + 1)
    assert engine.count_info()["n_chunks"] >= 5
    a3, b3, c3 = cc2.fetch(device=True)
    assert torch.equal(a3, ka) and torch.equal(b3, kb) and torch.equal(c3, kc)
    # click_to_cart_or_buy == click_to_cart + click_to_buy
    cb = engine.count("click_to_cart_or_buy")
    m = engine.merge([engine.count(type_this=0, next_types=[1], window=86400),
                      engine.count(type_this=0, next_types=[2], window=86400)])
    for x, z in zip(cb.fetch(device=True), m.fetch(device=True)):
        assert torch.equal(x, z)
    # top-20: per aid non-increasing counts, n_valid = min(20, segment size), ids inside the segment
    f = engine.filter(cc, 2)
    ax, nv, ay, ac = engine.topk(f, 20, device=True)
    assert bool((ac[:, :-1] >= ac[:, 1:]).all())
    fa, fb, fc = f.fetch(device=True)
    seg = torch.unique_consecutive(fa, return_counts=True)
    assert torch.equal(seg[0], ax) and torch.equal(torch.clamp(seg[1], max=20).to(torch.int32), nv)
    assert int(ac[:, 0].max()) == int(fc.max())


def test_weighted_time_decay_extension(engine):
    """EXTENSION (north_star config 4; NO reference counterpart): every counted pair contributes
    w = max(0.10, 1 - |dt| / window) to score(aid, aid_next).  Pinned by this repo's float64 oracle
    (oracle/cov_oracle.c::cov_oracle_score): same keys, same integer counts, scores within 1e-5 relative (the
    fixed-point sums are in fact within 3e-7); top-k by (score desc, aid_next asc) is checked against the table."""
    cases = [small_events(71, n_sessions=600, n_aids=60, max_len=40)]
    d = generate_numpy(SynthSpec(n_sessions=20_000, seed=23, force_long_click_session=465))     # long-tail shape of config 4
    cases.append((d["session"], d["aid"], d["ts"], d["type"]))
    for ci, (s, a, t, y) in enumerate(cases):
        engine.load_events(s, a, t, y)
        for name in NAMES if ci == 0 else ("click_to_cart_or_buy", "click_to_click"):
            oa, ob, osc, oc = c_oracle.score_name(s, a, t, y, name)
            ia, ib, ic = engine.count(name).fetch()
            assert np.array_equal(ia, oa) and np.array_equal(ib, ob) and np.array_equal(ic.astype(np.uint32), oc)
            for mc in (1, 3):
                wt = engine.count_weighted(name, min_count=mc)
                ga, gb, gs, gc = wt.fetch()
                keep = oc >= mc
                assert np.array_equal(ga, oa[keep]) and np.array_equal(gb, ob[keep]) and np.array_equal(gc.astype(np.uint32), oc[keep])
                if len(gs):
                    rel = np.abs(gs - osc[keep]) / osc[keep]
                    assert rel.max() <= 1e-5, (name, mc, rel.max())            # north_star's tolerance
                    assert rel.max() <= 1e-6                                    # what 24 fractional bits really give
                    assert np.all(gs >= 0.1 * gc - 1e-9) and np.all(gs <= gc + 1e-9)
                for k in (1, 5, 20):
                    ta, tb, ts_, tr = wt.topk(k)
                    o = np.lexsort((gb, -gs, ga))                               # fixed-point sums are exact in float64
                    xa, xb, xs = ga[o], gb[o], gs[o]
                    start = np.r_[True, xa[1:] != xa[:-1]] if len(xa) else np.zeros(0, bool)
                    seg = np.maximum.accumulate(np.where(start, np.arange(len(xa)), 0)) if len(xa) else np.zeros(0, np.int64)
                    rank = np.arange(len(xa)) - seg + 1
                    kk = rank <= k
                    assert np.array_equal(ta, xa[kk]) and np.array_equal(tb, xb[kk]) and np.array_equal(ts_, xs[kk]) and np.array_equal(tr, rank[kk])
                wt.free()


def test_reference_run_fixtures_gpu(engine):
    """The CUDA path against the reference's OWN output (tests/golden/ref_*.json, tools/gen_reference_fixtures.py);
    skipped loudly while the fixtures cannot be generated (polars absent)."""
    from test_oracle_cpu import _ref_fixtures, check_against_reference_fixture
    files = _ref_fixtures()
    if not files:
        pytest.skip("PARITY UNPINNED: no tests/golden/ref_*.json (polars not installable here); see tools/gen_reference_fixtures.py")

    def count_fn(s, a, t, y, kind):
        engine.load_events(s, a, t, y)
        return engine.count(kind).to_dict()

    def topn_fn(s, a, t, y, kind, first_n):
        engine.load_events(s, a, t, y)
        la, lb, lc, _ = topn_long(*engine.topk(engine.count(kind), first_n))
        out = {}
        for x, c in zip(la.tolist(), lc.tolist()):
            out.setdefault(x, []).append(c)
        return out

    for f in files:
        check_against_reference_fixture(f, count_fn, topn_fn)
