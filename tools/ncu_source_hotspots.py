#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report (needs -lineinfo).
usage: tools/ncu_source_hotspots.py report.ncu-rep kernel_regex [top_n] [launch_index]
HOTSPOT_SORT=1 sorts by executed instructions instead of stall samples."""
import collections
import csv
import io
import os
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    sort_col = int(os.environ.get("HOTSPOT_SORT", "0"))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # split into (file, function) sections; a kernel launch = consecutive sections until the function name repeats its first file
    launches = []; cur = None; fpath = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1]; continue
        if len(r) == 2 and r[0] == "Function Name":
            if cur is None or (cur["files"] and fpath == cur["files"][0]):
                cur = {"files": [], "lines": collections.defaultdict(lambda: [0, 0, 0, ""])}; launches.append(cur)
            cur["files"].append(fpath); cur["hdr"] = None; continue
        if cur is None:
            continue
        if cur["hdr"] is None:
            cur["hdr"] = r; continue
        h = cur["hdr"]
        try:
            ln = r[0]; src = r[1]
            smp = int(r[h.index("# Samples")] or 0); ins = int(r[h.index("Instructions Executed")] or 0)
            thr = int(r[h.index("Thread Instructions Executed")] or 0)
        except (ValueError, IndexError):
            continue
        if not ln.strip():
            continue   # SASS rows under a CUDA line: already included in the line's totals
        key = (fpath.split("/")[-1], ln)
        e = cur["lines"][key]; e[0] += smp; e[1] += ins; e[2] += thr
        if src.strip():
            e[3] = src.strip()
    L = launches[which]["lines"]
    ts = sum(v[0] for v in L.values()); ti = sum(v[1] for v in L.values()); tt = sum(v[2] for v in L.values())
    print("launch %d/%d: samples %d, warp-instr %d, thread-instr %d (lanes/instr %.1f)" % (which, len(launches), ts, ti, tt, tt / max(ti, 1)))
    for (f, ln), v in sorted(L.items(), key=lambda kv: -kv[1][sort_col])[:top_n]:
        print("%5.1f%% smp %5.1f%% inst lanes %4.1f  %s:%s  %s" % (100.0 * v[0] / max(ts, 1), 100.0 * v[1] / max(ti, 1), v[2] / max(v[1], 1), f, ln, v[3][:90]))


if __name__ == "__main__":
    main()
