"""Warp-stall samples of a profiled kernel, split at its barriers (source page of a .ncu-rep)."""
import csv
import subprocess
import sys


def main():
    report = sys.argv[1]
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    header = rows[1]
    col = {name: header.index(name) for name in ("Source", "# Samples", "stall_long_sb", "stall_short_sb", "stall_barrier",
                                                  "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected",
                                                  "Instructions Executed")}
    # a report with several kernels repeats the "Kernel Name" / header lines: keep the first kernel (or argv[3]-th)
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    starts = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    begin = starts[which] + 2
    end = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print("kernel:", rows[starts[which]][1])
    data = [r for r in rows[begin:end] if len(r) == len(header)]
    total = sum(int(r[col["# Samples"]]) for r in data)
    bars = [k for k, r in enumerate(data) if "BAR" in r[col["Source"]]]
    print("total samples", total, "barriers at", bars)
    previous = 0
    for b in bars + [len(data)]:
        region = data[previous:b]
        line = {name: sum(int(r[col[name]]) for r in region) for name in col if name not in ("Source", "Instructions Executed")}
        print(f"region [{previous},{b})", line)
        previous = b
    print("top instructions by samples:")
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][col["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        print(k, r[col["Source"]].strip()[:64], "| samples", r[col["# Samples"]], "long", r[col["stall_long_sb"]], "short",
              r[col["stall_short_sb"]], "barrier", r[col["stall_barrier"]], "exec", r[col["Instructions Executed"]])


if __name__ == "__main__":
    main()
