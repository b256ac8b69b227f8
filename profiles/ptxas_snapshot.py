"""Writes profiles/r3_ptxas_final.md: registers, stack and spills per kernel from the `ptxas -v` logs the csrc
Makefile leaves under se3-icp_b200/csrc/build_logs/ (git-ignored; this table is the tracked snapshot)."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def collect():
    rows = []
    for f in sorted(glob.glob(os.path.join(ROOT, "se3-icp_b200", "csrc", "build_logs", "*.ptxas.log"))):
        cur, stack = None, None
        for line in open(f):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                stack = m.groups()
                continue
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                sm = re.search(r"(\d+) bytes smem", line)
                if "cub" not in cur:
                    rows.append((os.path.basename(f).split(".")[0], cur, int(m.group(1)), stack, sm.group(1) if sm else "0"))
                cur = None
    return rows


def main():
    rows = collect()
    names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.split("\n")
    out = ["# ptxas -v snapshot of the final build (registers / stack / spills per kernel; cub instantiations omitted)",
           "# regenerate: make -C se3-icp_b200/csrc && python profiles/ptxas_snapshot.py", "",
           "| file | kernel | registers | stack B | spill st / ld B | static smem B |", "|---|---|---:|---:|---:|---:|"]
    for r, n in zip(rows, names):
        n = n.split("(")[0].replace("se3::", "")
        out.append("| %s.cu | %s | %d | %s | %s / %s | %s |" % (r[0], n, r[2], r[3][0], r[3][1], r[3][2], r[4]))
    with open(os.path.join(ROOT, "profiles", "r3_ptxas_final.md"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("%d kernels" % len(rows))


if __name__ == "__main__":
    main()
