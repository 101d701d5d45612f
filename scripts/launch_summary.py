"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: total time, launches, share.

    python scripts/launch_summary.py gpurun_out/launches.csv "command that was profiled" > profiles/..._summary.txt
"""
import csv
import sys
from collections import defaultdict


def main(path: str, command: str) -> None:
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        tot[r["Kernel Name"]] += v_us
        cnt[r["Kernel Name"]] += 1
    total = sum(tot.values())
    print(f"# {command}")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: SHARES are "
          f"meaningful, not absolute times); {sum(cnt.values())} launches, {total / 1e3:.1f} ms in total")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{t:14.1f} us  x{cnt[name]:<5d}{100 * t / total:6.1f}%  {name[:150]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
