"""Per-source-line instruction and stall-sample totals from an `ncu --page source --csv --print-source cuda,sass` export.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/ncu_lines.py src.csv [kernel-substring] [top-n]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    launches = []  # one dict per kernel launch section (a launch restarts the file list: a file name repeats)
    cur_file, cur_fn, hdr, agg = None, None, None, None
    seen_files = set()
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            if cur_fn != row[1] or cur_file in seen_files:
                agg = defaultdict(lambda: [0, 0, ""])
                launches.append((row[1], agg))
                seen_files = set()
            seen_files.add(cur_file)
            cur_fn = row[1]
            continue
        if row[0] == "Line No":
            hdr = {h: i for i, h in enumerate(row)}
            i_inst = row.index("Instructions Executed")
            i_samp = row.index("# Samples")
            continue
        if hdr is None or agg is None:
            continue
        if row[2] != "-":  # SASS row under a source line: the source-line row already aggregates them
            continue
        try:
            inst = int(row[i_inst])
            samp = int(row[i_samp])
        except (ValueError, IndexError):
            continue
        key = (cur_file, int(row[0]))
        agg[key][0] += inst
        agg[key][1] += samp
        agg[key][2] = row[1].strip()
    for name, agg in launches:
        if want not in name:
            continue
        tot_i = sum(v[0] for v in agg.values())
        tot_s = sum(v[1] for v in agg.values())
        print("== %s: %.1f M warp instructions, %d samples" % (name[:70], tot_i / 1e6, tot_s))
        for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1), f, ln, v[2][:90]))


if __name__ == "__main__":
    main()
