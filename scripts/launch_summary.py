"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the last N launches (one NES iteration)."""
import csv
import sys


def main():
    path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 9
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    last = rows[-n:]
    tot = sum(x[3] for x in last)
    for k, g, b, us in last:
        print("%-34s grid %-16s block %-14s %8.1f us %5.1f%%" % (k, g, b, us, 100 * us / tot))
    print("sum %.1f us" % tot)


if __name__ == "__main__":
    main()
