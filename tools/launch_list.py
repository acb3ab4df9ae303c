#!/usr/bin/env python
"""Aggregates an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X python bench.py --steps 1
--warmup 1 ...`) per kernel for ONE full-batch classify call: calls start at k_flags (the first kernel of the fused path);
the second call that takes more than `min_ms` is the timed step (the first is the warm-up).
usage: python tools/launch_list.py launches.csv [min_ms = 30] > profiles/rNN_launches_c3_step.csv"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v * 1e3 if r[iu] in ("s", "second") else v
    seq.append((r[ik], ms))
starts = [i for i, (k, _) in enumerate(seq) if re.search(r"\bk_flags\b", k)] + [len(seq)]
min_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
calls = [(a, b) for a, b in zip(starts[:-1], starts[1:]) if sum(ms for _, ms in seq[a:b]) > min_ms]
a, b = calls[1] if len(calls) > 1 else calls[0]
agg = {}
for k, ms in seq[a:b]:
    k = re.sub(r"^void\s+", "", k)
    k = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", k)
    k = re.sub(r"\(.*$", "", k)
    k = re.sub(r"cub::(CUB_[0-9_A-Za-z]+::)?", "cub::", k)
    k = re.sub(r"(cub::\w+)<.*", r"\1", k)
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + ms)
tot = sum(t for _, t in agg.values())
print("# one timed step under ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.2f ms serialised "
      "(%d full-batch calls in the list)" % (b - a, tot, len(calls)))
print("kernel,launches,ms,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.4f,%.4f" % (k, n, t, t / tot))
