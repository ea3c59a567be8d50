"""Synthetic OTTO-shaped event generator.

Shapes follow the reference's published data statistics: 12.9 M sessions / 220 M events /
1.8 M aids (README.md:10-12), aids-per-session distribution of model/w2vec_aids.py:228
(mean 15.4, median 6, 95 % 62, 99 % 152, max 498), schema and dtypes of
etl/jsonl_to_parquet.py:23-29 (session i32, aid i32, ts i32 seconds, type i8).

Written with torch ops only so the same code runs on the CPU (tests, small fixtures) and on
the GPU (bench at full size, < 1 s).  Deterministic per (seed, device type).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

TS0 = 1_659_304_800          # 2022-08-01 00:00 CEST, start of the OTTO train window
FOUR_WEEKS = 28 * 86400
P_CLICK, P_CART, P_ORDER = 0.8985, 0.0780, 0.0235


@dataclass
class SynthSpec:
    n_sessions: int = 100_000
    n_aids: int = 1_800_000
    seed: int = 42
    first_session: int = 0
    mu: float = 2.0               # LogNormal(mu, sigma) session length
    sigma: float = 1.3
    max_len: int = 498
    p_local: float = 0.6          # event drawn from the session's topic neighbourhood
    p_repeat: float = 0.25        # event revisits an earlier aid of the session
    local_scale: float = 20.0
    zipf_shift: float = 200.0     # p(rank) ~ 1/(rank+shift): flat head like the real catalogue
    dup_frac: float = 0.001       # exact duplicate rows injected
    force_long_click_session: int = 0   # config 4: one session with this many clicks (465)
    order: str = "sorted"         # "sorted" (ETL order) | "shuffled"


def _zipf_rank(u: torch.Tensor, n: int, q: float) -> torch.Tensor:
    # inverse CDF of p(r) ~ 1/(r+q) on [0, n)
    r = q * torch.pow(torch.tensor((n + q) / q, dtype=torch.float64, device=u.device), u.double()) - q
    return r.long().clamp_(0, n - 1)


def generate(spec: SynthSpec, device: str | torch.device = "cpu") -> Dict[str, torch.Tensor]:
    """-> dict(session i32, aid i32, ts i32, type i8) on `device`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(spec.seed)
    S, N = spec.n_sessions, spec.n_aids

    ln = torch.exp(spec.mu + spec.sigma * torch.randn(S, generator=g, device=dev, dtype=torch.float32))
    ln = ln.round().clamp_(1, spec.max_len).long()
    if spec.force_long_click_session:
        ln[S // 2] = spec.force_long_click_session
    off = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    torch.cumsum(ln, 0, out=off[1:])
    E = int(off[-1].item())
    sess = torch.repeat_interleave(torch.arange(S, device=dev), ln, output_size=E)
    pos = torch.arange(E, device=dev) - off[sess]

    # types
    u = torch.rand(E, generator=g, device=dev)
    typ = (u >= P_CLICK).to(torch.int8) + (u >= P_CLICK + P_CART).to(torch.int8)
    if spec.force_long_click_session:
        a, b = int(off[S // 2].item()), int(off[S // 2 + 1].item())
        typ[a:b] = 0

    # timestamps: 80 % short gaps Exp(90 s), 20 % long gaps LogNormal(median 1 day, 1.2)
    u = torch.rand(E, generator=g, device=dev)
    short = -90.0 * torch.log1p(-torch.rand(E, generator=g, device=dev).clamp_(max=0.999999))
    long_ = torch.exp(math.log(86400.0) + 1.2 * torch.randn(E, generator=g, device=dev))
    gap = torch.where(u < 0.8, short, long_).clamp_(0, 30 * 86400.0).long()
    gap[pos == 0] = 0
    if spec.force_long_click_session:
        # keep the long session inside one window so the quadratic tail is exercised
        gap[a:b] = torch.randint(0, 60, (b - a,), generator=g, device=dev)
        gap[a] = 0
    cs = torch.cumsum(gap, 0)
    base = cs[off[:-1]]
    start = TS0 + torch.randint(0, FOUR_WEEKS, (S,), generator=g, device=dev)
    ts = (start[sess] + cs - base[sess]).clamp_(max=2**31 - 1)

    # aids: topic-local + global popularity + in-session revisits
    centre = _zipf_rank(torch.rand(S, generator=g, device=dev), N, spec.zipf_shift)
    glob = _zipf_rank(torch.rand(E, generator=g, device=dev), N, spec.zipf_shift)
    lap = torch.rand(E, generator=g, device=dev) - 0.5
    offs = (-spec.local_scale * torch.sign(lap) * torch.log1p(-2 * lap.abs().clamp_(max=0.499999))).round().long()
    local = (centre[sess] + offs).remainder(N)
    rank = torch.where(torch.rand(E, generator=g, device=dev) < spec.p_local, local, glob)
    rep = (torch.rand(E, generator=g, device=dev) < spec.p_repeat) & (pos > 0)
    src = off[sess] + (torch.rand(E, generator=g, device=dev) * pos.float()).long().clamp_(min=0)
    src = torch.minimum(src, off[sess] + pos - 1).clamp_(min=0)
    rank = torch.where(rep, rank[src], rank)
    # fixed pseudo-random bijection rank -> aid so popular items are spread over the id space
    gp = torch.Generator(device="cpu"); gp.manual_seed(spec.seed + 7919)
    if N <= 50_000_000:
        perm = torch.randperm(N, generator=gp).to(dev)
        aid = perm[rank]
    else:
        aid = rank
    sess = sess + spec.first_session

    # exact duplicate rows (count_co_events.py:92 drops them)
    nd = int(E * spec.dup_frac)
    if nd > 0:
        pick = torch.randint(0, E, (nd,), generator=g, device=dev)
        sess = torch.cat([sess, sess[pick]]); aid = torch.cat([aid, aid[pick]])
        ts = torch.cat([ts, ts[pick]]); typ = torch.cat([typ, typ[pick]])
    n = sess.numel()
    if spec.order == "shuffled":
        order = torch.randperm(n, generator=gp).to(dev)
    else:
        key = sess * (1 << 32) + ts
        order = torch.sort(key, stable=True).indices
    return {
        "session": sess[order].to(torch.int32).contiguous(),
        "aid": aid[order].to(torch.int32).contiguous(),
        "ts": ts[order].to(torch.int32).contiguous(),
        "type": typ[order].contiguous(),
    }


def generate_numpy(spec: SynthSpec):
    d = generate(spec, "cpu")
    return {k: v.numpy() for k, v in d.items()}
