#!/usr/bin/env python
"""Summarise ncu CSV exports.

  python tools/ncu_table.py launches <launch-list.csv>      # `ncu --metrics gpu__time_duration.sum --csv --log-file ...`
  python tools/ncu_table.py full <raw-page.csv> [regex]     # `ncu -i x.ncu-rep --page raw --csv`

`launches` prints time, share and average per kernel; `full` prints one line per profiled launch with the metrics the
roofline discussion in DESIGN.md uses (duration, tensor-pipe %, L2 / DRAM throughput %, DRAM bytes, active-cycle share)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    for junk in ("void ", "lbc::", "fast::", "ref::"):
        name = name.replace(junk, "")
    return name


def launches(path):
    hdr, agg, n = None, collections.OrderedDict(), 0
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1e-3)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.2f ms serialised" % (n, tot / 1e3))
    print("%-58s %6s %9s %7s %9s" % ("kernel", "count", "ms", "share", "us/launch"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %6d %9.3f %6.1f%% %9.1f" % (k[:58], a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0]))
    return agg


FULL_COLS = [
    ("us", "gpu__time_duration.sum"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dramR_MB", "dram__bytes_read.sum"),
    ("dramW_MB", "dram__bytes_write.sum"),
    ("regs", "launch__registers_per_thread"),
    ("smem_KB", "launch__shared_mem_per_block_dynamic"),
    ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]


def full(path, pattern=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(lab, idx[m]) for lab, m in FULL_COLS if m in idx]
    act = (idx.get("sm__cycles_active.avg"), idx.get("sm__cycles_elapsed.avg"))
    print("%-44s %-12s " % ("kernel", "grid") + " ".join("%9s" % c for c, _ in cols) + "   active/elapsed")
    out = []
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        vals = []
        for _, i in cols:
            try:
                vals.append(float(r[i]))
            except ValueError:
                vals.append(float("nan"))
        ratio = float("nan")
        if act[0] is not None and act[1] is not None:
            try:
                ratio = float(r[act[0]]) / float(r[act[1]])
            except (ValueError, ZeroDivisionError):
                pass
        print("%-44s %-12s " % (name[:44], r[idx["Grid Size"]].replace(" ", "")) + " ".join("%9.2f" % v for v in vals) +
              "   %.2f" % ratio)
        out.append((name, vals))
    return out


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "full"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
