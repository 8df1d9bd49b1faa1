#!/usr/bin/env python
"""Turns the raw ncu outputs that gpurun brings back (gpurun_out/, scratch) into the small text summaries
committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python profiles/summarize.py kernels  gpurun_out/prof.ncu-rep            > profiles/rNN_kernels.txt
  python profiles/summarize.py stalls   gpurun_out/prof.ncu-rep <kernel>   > profiles/rNN_stalls_<kernel>.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
    allt = sum(v[1] for v in tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {sum(v[0] for v in tot.values())} launches, "
          f"{allt / 1e6:.3f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':60s} {'n':>5s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:5d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / allt:6.1f}%")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def kernels(rep):
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none ; one block per captured launch")
    for r in rows:
        print(f"\n{r[idx['Kernel Name']].split('(')[0]}")
        for m in METRICS:
            if m in idx:
                print(f"  {m:70s} {r[idx[m]]:>18s} {units[idx[m]]}")
        try:
            rd = float(r[idx['dram__bytes_read.sum']].replace(",", "")); wr = float(r[idx['dram__bytes_write.sum']].replace(",", ""))
            u = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx['dram__bytes_read.sum']]]
            uw = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx['dram__bytes_write.sum']]]
            t = float(r[idx['gpu__time_duration.sum']].replace(",", ""))
            tu = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}[units[idx['gpu__time_duration.sum']]]
            print(f"  {'=> dram traffic (read+write) per launch':70s} {(rd * u + wr * uw) / 1e9:18.4f} GB")
            print(f"  {'=> dram GB/s under ncu':70s} {(rd * u + wr * uw) / 1e9 / (t * tu):18.1f}")
        except Exception:
            pass


def stalls(rep, kernel, ntop=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or not r[idx["# Samples"]].isdigit():
            continue
        if r[idx["Address"]] in seen:
            continue  # the page repeats per captured launch: keep the first
        seen.add(r[idx["Address"]])
        data.append(r)
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[idx[h]]) for r in data) for h in reasons}
    print(f"# {kernel}: {tot} warp-state samples over {len(data)} SASS instructions")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {k:25s} {100 * v / max(tot, 1):5.1f}%")
    print("# hottest instructions (program order)")
    for pos, r in enumerate(data):
        r.append(pos)
    top = sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:ntop]
    for r in sorted(top, key=lambda r: r[-1]):
        st = max(reasons, key=lambda h: int(r[idx[h]]))
        print(f"  @{r[-1]:5d} {100 * int(r[idx['# Samples']]) / max(tot, 1):5.1f}% {st:22s} {r[idx['Source']].strip()[:90]}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernels":
        kernels(sys.argv[2])
    elif cmd == "stalls":
        stalls(sys.argv[2], sys.argv[3])
