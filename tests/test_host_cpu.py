"""CPU suite (-m "not gpu"): the C-ABI library loads and exports every declared symbol, the product
fails loudly without a GPU (no CPU fallback), config mirrors the reference, generator sanity."""
import ctypes
import os
import re

import numpy as np
import pytest

import otto_recommender_b200 as pkg
from otto_recommender_b200 import _lib
from otto_recommender_b200.config import DEFAULT_CONFIG
from otto_recommender_b200.dist import hash_dest, shard_bounds
from otto_recommender_b200.synth import SynthSpec, generate_numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "ottocov.h")).read()
    declared = set(re.findall(r"\b(ottocov_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in ottocov.h"
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in ottocov.h but not exported"
    # the ctypes table binds exactly the declared functions
    assert declared == set(_lib.SYMBOLS)
    assert _lib.load_library().ottocov_version() == 200


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.OttocovError) as e:
        pkg.Engine(device=0)
    assert "no CPU fallback" in str(e.value)
    from otto_recommender_b200 import count_co_events as cce
    cce._engine = None
    with pytest.raises(pkg.OttocovError):
        cce.count_co_events({"session": [1], "aid": [1], "ts": [1], "type": [0]})


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "otto_recommender_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "cov_oracle" not in src, f


def test_hash_dest_matches_c(built_lib):
    lib = _lib.load_library()
    aids = np.array([0, 1, 2, 12345, 1_799_999, 2**31 - 1, 77777], dtype=np.int64)
    for r in (1, 2, 3, 4, 8):
        want = np.array([lib.ottocov_hash_dest(int(a), r) for a in aids])
        assert np.array_equal(hash_dest(aids, r), want)
    d = hash_dest(np.arange(200_000), 8)
    share = np.bincount(d, minlength=8) / len(d)
    assert share.min() > 0.11 and share.max() < 0.14


def test_config_matches_reference():
    ref = "/root/reference/config.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this box")
    ns = {}
    src = open(ref).read()
    # evaluate only the co-count block (lines 39-96): no logging / mkdir side effects
    block = "\n".join(src.splitlines()[38:96])
    exec(block, ns)
    c = DEFAULT_CONFIG
    assert c.MIN_TIME_TO_NEXT == ns["MIN_TIME_TO_NEXT"] and c.MAX_TIME_TO_NEXT == ns["MAX_TIME_TO_NEXT"]
    assert c.MAP_MAX_TIME_TO_NEXT == ns["MAP_MAX_TIME_TO_NEXT"]
    assert c.MIN_COUNT_TO_SAVE == ns["MIN_COUNT_TO_SAVE"]
    assert c.MIN_COUNT_IN_PART == ns["MIN_COUNT_IN_PART"]
    assert c.MAX_CO_EVENT_PAIRS_TO_SAVE_DISK == ns["MAX_CO_EVENT_PAIRS_TO_SAVE_DISK"]
    assert list(c.CO_EVENTS_TO_COUNT) == ns["CO_EVENTS_TO_COUNT"]
    assert {k: (v[0], list(v[1])) for k, v in c.MAP_NAME_COUNT_TYPE.items()} == \
        {k: (v[0], list(v[1])) for k, v in ns["MAP_NAME_COUNT_TYPE"].items()}
    assert c.RETRIEVAL_FIRST_N_CO_COUNTS == ns["RETRIEVAL_FIRST_N_CO_COUNTS"]
    assert c.OPTIM_ROWS_POLARS_GROUPBY == ns["OPTIM_ROWS_POLARS_GROUPBY"]
    assert c.MAX_ROWS_POLARS_GROUPBY == ns["MAX_ROWS_POLARS_GROUPBY"]
    assert c.spec("click_to_cart_or_buy") == (0, 0b110, 86400)
    assert c.spec("click_to_click") == (0, 0b001, 43200)


def test_synth_shape_and_determinism():
    d1 = generate_numpy(SynthSpec(n_sessions=20_000, seed=3))
    d2 = generate_numpy(SynthSpec(n_sessions=20_000, seed=3))
    for k in d1:
        assert np.array_equal(d1[k], d2[k])
    assert d1["session"].dtype == np.int32 and d1["type"].dtype == np.int8
    lens = np.bincount(d1["session"])
    assert 12 < lens.mean() < 22 and 4 <= np.median(lens) <= 9 and lens.max() <= 510
    share = np.bincount(d1["type"], minlength=3) / len(d1["type"])
    assert abs(share[0] - 0.8985) < 0.01 and abs(share[1] - 0.078) < 0.01
    key = d1["session"].astype(np.int64) * 2**32 + d1["ts"]
    assert np.all(np.diff(key) >= 0)                      # ETL order: by session, then time
    d3 = generate_numpy(SynthSpec(n_sessions=2_000, seed=3, force_long_click_session=465))
    lens3 = np.bincount(d3["session"][d3["type"] == 0])
    assert lens3.max() >= 465


def test_shard_bounds_balanced():
    rng = np.random.default_rng(0)
    lens = np.clip(np.round(np.exp(2.0 + 1.3 * rng.standard_normal(100_000))), 1, 498).astype(np.int64)
    for r in (1, 2, 4, 8):
        b = shard_bounds(lens, r)
        assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0) and len(b) == r + 1
        w = lens * np.minimum(lens, 32)
        per = np.array([w[b[i]:b[i + 1]].sum() for i in range(r)])
        assert per.max() / per.mean() < 1.02


def test_key_mix_is_a_bijection():
    """The hash path of the reduce-by-key rests on the key mix being invertible (equal keys stay equal, distinct
    keys stay distinct, the plain key comes back exactly).  Host code of the library: no GPU needed."""
    from otto_recommender_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(0)
    for ab in (1, 2, 3, 7, 11, 16, 21, 23, 24, 28):
        lim = 1 << ab
        if ab <= 3:                                    # exhaustive: every key of 2 ab bits maps to a distinct mixed key
            seen = set()
            for x in range(lim):
                for y in range(lim):
                    h = lib.ottocov_key_mix(ab, x, y)
                    assert h < (1 << (2 * ab)) and lib.ottocov_key_unmix(ab, h) == (x << 32 | y)
                    seen.add(h)
            assert len(seen) == lim * lim
        xs = rng.integers(0, lim, 2000); ys = rng.integers(0, lim, 2000)
        edge = [(0, 0), (lim - 1, lim - 1), (0, lim - 1), (lim - 1, 0)]
        for x, y in list(zip(xs.tolist(), ys.tolist())) + edge:
            h = lib.ottocov_key_mix(ab, x, y)
            assert h < (1 << (2 * ab))
            assert lib.ottocov_key_unmix(ab, h) == (x << 32 | y)
    # the top bits of the mixed key spread a skewed key set evenly: 2^8 buckets of a 100 k-key Zipf-like set
    ab = 21
    a = np.minimum((rng.pareto(1.1, 100_000) * 50).astype(np.int64), (1 << ab) - 1)
    b = np.minimum((rng.pareto(1.1, 100_000) * 50).astype(np.int64), (1 << ab) - 1)
    keys = np.unique(a << 32 | b)
    top = np.array([lib.ottocov_key_mix(ab, int(k >> 32), int(k & 0xFFFFFFFF)) >> (2 * ab - 8) for k in keys])
    counts = np.bincount(top, minlength=256)
    assert counts.max() < 2.0 * len(keys) / 256 and counts.min() > 0.4 * len(keys) / 256
    assert lib.ottocov_key_mix(29, 0, 0) == 2 ** 64 - 1      # out of range is loud


def test_columns_wider_than_the_abi_are_range_checked():
    """The C ABI takes int32 / int8 columns; wider inputs are narrowed only after a range check (raw OTTO
    millisecond timestamps or 64-bit session ids must not wrap silently)."""
    from otto_recommender_b200.engine import _as_col
    ok = np.array([1_660_000_000, 1_660_000_100], dtype=np.int64)
    p, where, keep = _as_col(ok, np.int32, "int32")
    assert keep.dtype == np.int32 and keep.tolist() == ok.tolist()
    with pytest.raises(ValueError):
        _as_col(ok * 1000, np.int32, "int32")                      # milliseconds
    with pytest.raises(ValueError):
        _as_col(np.array([0, 3, 200], dtype=np.int64), np.int8, "int8")
    with pytest.raises(ValueError):
        _as_col(np.array([0.5, 1.0]), np.int32, "int32")
    import torch
    with pytest.raises(ValueError):
        _as_col(torch.tensor([1_660_000_000_000], dtype=torch.int64), np.int32, "int32")
    _, _, t = _as_col(torch.tensor([5, 6], dtype=torch.int64), np.int32, "int32")
    assert t.dtype == torch.int32
