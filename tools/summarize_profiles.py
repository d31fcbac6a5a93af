"""Turns ncu outputs into the tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <launches.csv> <out.md> <title> [first_id last_id]
      launch list (ncu --metrics gpu__time_duration.sum --csv): per-launch table of ONE step + per-kernel shares
  python tools/summarize_profiles.py full <report.ncu-rep> <out.md> <raw_out.csv> <title>
      ncu --set full capture: one row per profiled launch (time, DRAM bytes, throughputs, occupancy, registers)
"""
import csv
import json
import os
import re
import subprocess
import sys


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def launches(path, out, title, lo=None, hi=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ci = {h: i for i, h in enumerate(hdr)}
    recs = [(int(r[ci["ID"]]), short(r[ci["Kernel Name"]]), r[ci["Grid Size"]], r[ci["Block Size"]], float(r[ci["Metric Value"]]) / 1e3)
            for r in rows if r[ci["Metric Name"]] == "gpu__time_duration.sum"]
    if lo is None:
        # one step = the launches from one k_copy_flag to the next; the second complete one (the first is the
        # bit-exactness check through mci_filter_run, the later ones are warm-up / timed resident steps)
        starts = [k for k, r in enumerate(recs) if r[1].startswith("k_copy_flag")] + [len(recs)]
        segs = [(a, b) for a, b in zip(starts, starts[1:]) if any(r[1].startswith("k_dt_emit") or r[1].startswith("k_dt_tree") or r[1].startswith("k_apply") for r in recs[a:b])]
        a, b = segs[1] if len(segs) > 1 else segs[0]
        step = recs[a:b]
    else:
        step = [r for r in recs if lo <= r[0] <= hi]
    total = sum(r[4] for r in step)
    with open(out, "w") as f:
        f.write("# %s\n\nPer-launch times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES with the CUDA-event "
                "table in bench.py's JSON (`kernels`), not absolutes. Raw CSV: `%s`.\n\n" % (title, os.path.relpath(path)))
        f.write("| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|\n")
        for k, r in enumerate(step):
            f.write("| %d | %s | %s | %s | %.1f | %.1f%% |\n" % (k, r[1], r[2], r[3], r[4], 100 * r[4] / total))
        f.write("\nstep total (kernels only, under ncu): %.1f us; %d launches\n\n| kernel | launches | us per step | share |\n|---|---|---|---|\n"
                % (total, len(step)))
        agg = {}
        for r in step:
            a = agg.setdefault(r[1], [0, 0.0])
            a[0] += 1
            a[1] += r[4]
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (k, n, t, 100 * t / total))


def full(rep, out, raw_out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(raw_out, "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=1.0):
        if name not in ci or r[ci[name]] in ("", "n/a"):
            return float("nan")
        v = float(r[ci[name]].replace(",", ""))
        u = units[ci[name]]
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return v * mult * scale

    traffic = {}
    with open(out, "w") as f:
        f.write("# %s\n\nRaw metrics: `%s`.\n\n| kernel | time us | DRAM read MB | DRAM write MB | DRAM %% of peak | SM throughput %% | "
                "issue active %% | warps active %% | regs | grid x block | warp-inst |\n|---|---|---|---|---|---|---|---|---|---|---|\n"
                % (title, os.path.relpath(raw_out)))
        for r in rows:
            name = short(r[ci["Kernel Name"]])
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            f.write("| %s | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %d | %s x %s | %d |\n" % (
                name, val(r, "gpu__time_duration.sum"), rd / 1e6, wr / 1e6,
                val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                int(val(r, "launch__registers_per_thread")), r[ci["Grid Size"]].strip("()").split(",")[0],
                r[ci["Block Size"]].strip("()").split(",")[0], int(val(r, "smsp__inst_executed.sum"))))
            key = name.split("<")[0]
            if key not in traffic:
                traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "source": os.path.relpath(raw_out)}
    return traffic


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4], *(int(v) for v in sys.argv[5:7]))
    else:
        t = full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur.update(t)
        json.dump(cur, open(p, "w"), indent=1)
