#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into small text summaries under profiles/.

    python tools/summarize_profiles.py r01_v5      # tag used in the output file names

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list) and every
gpurun_out/prof_*.ncu-rep (ncu --set full), writes profiles/<tag>_launches.txt and
profiles/<tag>_<kernel>.txt. Needs only the ncu CLI (no GPU).
"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"

RAW_METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(tag):
    f = SRC / "launches.csv"
    if not f.exists():
        return
    lines = [l for l in f.read_text().splitlines() if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list: 2 forward passes of cfg-2 (64 x 448x448, T=2), --clock-control none, cold-cache serialised",
           f"# total {sum(a[0] for a in agg.values())} launches, {tot:.0f} us (compare SHARES with bench.py's CUDA-event shares, not absolutes)",
           f"{'us':>10} {'launches':>8} {'share':>7}  kernel"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t:10.1f} {c:8d} {100 * t / tot:6.1f}%  {k[:120]}")
    (OUT / f"{tag}_launches.txt").write_text("\n".join(out) + "\n")
    print("wrote", OUT / f"{tag}_launches.txt")


def full(tag, rep):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full --clock-control none, {rep.name}: one section per captured launch"]
    for r in rows[2:]:
        out.append(f"\n== {r[idx['Kernel Name']][:110]}  (id {r[idx['ID']]})")
        for m in RAW_METRICS:
            if m in idx:
                out.append(f"  {m:72s} {r[idx[m]]:>16s} {units[idx[m]]}")
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--kernel-id", ":::1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 3:
        h = srows[1]
        ix = {c: i for i, c in enumerate(h)}
        data, seen = [], set()
        for r in srows[2:]:
            if len(r) == len(h) and r[ix["Address"]] not in seen:
                seen.add(r[ix["Address"]])
                data.append(r)
        num = lambda r, c: int(r[ix[c]]) if r[ix[c]].lstrip("-").isdigit() else 0
        tot = sum(num(r, "# Samples") for r in data) or 1
        stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = sorted(((sum(num(r, c) for r in data), c) for c in stall), reverse=True)[:6]
        out.append(f"\n-- warp-stall samples of the first captured launch ({tot} samples): " + ", ".join(f"{c}={v}" for v, c in agg))
        out.append("-- hottest SASS instructions (samples, share, instruction, executed)")
        for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:15]:
            out.append(f"  {num(r, '# Samples'):7d} {100 * num(r, '# Samples') / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} x{r[ix['Instructions Executed']]}")
    name = rep.stem.replace("prof_", "")
    (OUT / f"{tag}_{name}.txt").write_text("\n".join(out) + "\n")
    print("wrote", OUT / f"{tag}_{name}.txt")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    OUT.mkdir(exist_ok=True)
    launches(tag)
    for rep in sorted(SRC.glob("prof_*.ncu-rep")):
        full(tag, rep)
