"""Source-line hot spots of a kernel from an ncu report taken with --import-source on.

ncu's CSV export of the source page carries the per-instruction metrics only in the SASS view, so the SASS rows are
mapped back to source lines through nvdisasm's line info of the same cubin (built with -lineinfo):

  cuobjdump -xelf all itkcontourinterpolation_b200/_build/mci_dt.o          # -> mci_dt.sm_100a.cubin
  ncu -i report.ncu-rep --page source --print-source sass --csv --kernel-name regex:k_dt_tree > sass.csv
  python tools/ncu_source_hotspots.py sass.csv mci_dt.sm_100a.cubin '<mangled function name>' [top=25] [block=0]

Prints, per source line, the share of warp-stall samples and of executed warp instructions.
"""
import csv
import re
import subprocess
import sys


def line_table(cubin, func):
    out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    table, cur, inside = {}, None, False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            inside = ln.strip() == ".text.%s:" % func
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def blocks(path):
    cur, hdr = None, None
    res = []
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            res.append(cur)
        elif row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and row[0].startswith("0x"):
            cur["rows"].append(row)
    return res


def main():
    sass, cubin, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    b = blocks(sass)[which]
    ci = {n: i for i, n in enumerate(b["hdr"])}
    table = line_table(cubin, func)
    base = int(b["rows"][0][0], 16)
    agg = {}
    tot_s = tot_i = 0.0
    for r in b["rows"]:
        off = int(r[0], 16) - base
        key = table.get(off, ("?", 0))
        s = float(r[ci["# Samples"]] or 0)
        i = float(r[ci["Instructions Executed"]] or 0)
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += s
        a[1] += i
        tot_s += s
        tot_i += i
    print("%s: %d SASS rows, %d stall samples, %d warp instructions" % (b["name"][:60], len(b["rows"]), tot_s, tot_i))
    src_cache = {}
    for (fn, ln), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        try:
            if fn not in src_cache:
                import glob
                hits = glob.glob("**/" + fn, recursive=True)
                src_cache[fn] = open(hits[0]).read().splitlines() if hits else []
            text = src_cache[fn][ln - 1].strip()[:100] if src_cache[fn] else ""
        except Exception:
            pass
        print("%5.1f%% samples %5.1f%% inst  %s:%d  %s" % (100 * s / max(tot_s, 1), 100 * i / max(tot_i, 1), fn, ln, text))


if __name__ == "__main__":
    main()
