"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel table committed under profiles/.

usage: python tools/summarize_launches.py <launches.csv> <steps> [title] > profiles/<name>_summary.md
"""
from __future__ import annotations

import collections
import csv
import re
import sys


def load(path: str):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg: dict[str, list[float]] = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v  # -> us
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("lcd::", "").replace("void ", "").strip()
        grid = row.get("Grid Size", "")
        agg.setdefault(name, []).append((v, grid))
    return agg


def main() -> int:
    path, steps = sys.argv[1], int(sys.argv[2])
    title = sys.argv[3] if len(sys.argv) > 3 else path
    agg = load(path)
    total = sum(v for vs in agg.values() for v, _ in vs) / steps
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` over the timed region of bench.py "
          f"({steps} step(s); bench.py brackets it with cudaProfilerStart/Stop).  Per-launch times under ncu are serialised and cold-cache: the SHARE "
          "of the step is what must agree with bench.py's live CUDA-event numbers, not the absolute.\n")
    print("| kernel | launches/step | us/launch (grids) | us/step | share |")
    print("|---|---|---|---|---|")
    for name, vs in sorted(agg.items(), key=lambda kv: -sum(v for v, _ in kv[1])):
        tot = sum(v for v, _ in vs)
        per = ", ".join(f"{v:.1f} {g}" for v, g in vs[: len(vs) // steps]) if len(vs) // steps <= 4 else f"{tot / len(vs):.1f} avg"
        print(f"| `{name}` | {len(vs) / steps:g} | {per} | {tot / steps:.1f} | {100 * tot / steps / total:.1f}% |")
    print(f"| **total** | | | {total:.1f} | 100% |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
