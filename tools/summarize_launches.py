"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py <csv> [title]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
lines = [l for l in open(path) if l.startswith('"')]
agg = collections.OrderedDict()
total = 0.0
n = 0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v
    name = row["Kernel Name"]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ms
    a[1] += 1
    total += ms
    n += 1
OURS = ("rs_", "scan_onepass", "expand_kernel", "expand_scatter", "region_table", "publish_", "mirror_push", "rle_expand", "aid_max", "feat_", "sum_hist", "stripe_off", "init_minmax", "pop_", "hash_reduce", "rle_kernel", "ev_", "window", "tile_search", "topk",
        "hr_", "mix_", "unmix", "unpack", "order_keys", "table_stats", "pack_keys", "stamp", "strip", "unstamp")
ours = sum(v[0] for k, v in agg.items() if any(o in k for o in OURS))
print(f"# ncu launch list summary of: {title}")
print(f"# {n} launches, {total:.1f} ms total device time; cold-cache and serialised under ncu: compare SHARES, not absolutes")
print(f"# kernels of libottocov.so: {ours:.1f} ms ({100 * ours / total:.1f} %); the rest is torch generating the synthetic data set")
for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    mine = any(o in k for o in OURS)
    share = f"{100 * ms / ours:5.1f}% of ours" if mine else "  (torch)     "
    print(f"{ms:10.2f} ms {c:5d} launches {100 * ms / total:5.1f}% of all, {share}   {re.sub(r'[(].*', '', k)[:90]}")
