"""Per-kernel shares of one forward from an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

    python tools/launches_summary.py profiles/rNN_x_ncu_launches.csv profiles/rNN_x_ncu_launches_summary.md

A forward is a run of launches that starts with `ingest_kernel`; the one summarised is the first timed graph replay
(torch's own fill / copy kernels are skipped); under ncu every launch is serialised and cold-cache, so
SHARES are the quantity to compare with the live step, not the absolute times.
"""
import csv
import re
import sys
from collections import OrderedDict


def main(src, dst, which=5):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    head = rows[0]
    k_name, k_val = head.index("Kernel Name"), head.index("Metric Value")
    launches = [(r[k_name], float(r[k_val].replace(",", "")) / 1e3) for r in rows[1:]]
    ours = re.compile(r"(gemm_tap_kernel<[^>]*>|[a-z0-9_]+_kernel(?:<[^>]*>)?)")
    starts = [i for i, (n, _) in enumerate(launches) if "ingest_kernel" in n]
    per_fwd = starts[1] - starts[0] if len(starts) > 1 else len(launches) - starts[0]
    # bench.py --steps 2 --warmup 3 launches forwards 0-1 eagerly (plan warm-up), 2-4 as warm-up graph replays and 5-6 as the
    # timed single-stream replays; later forwards belong to the other lanes / the 1024-pair leg
    first = starts[which]
    agg = OrderedDict()
    total = 0.0
    n = 0
    for name, us in launches[first:first + per_fwd]:
        if "at::" in name:
            continue
        m = ours.search(name)
        key = m.group(1) if m else name[:60]
        c, t = agg.get(key, (0, 0.0))
        agg[key] = (c + 1, t + us)
        total += us
        n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src.split('/')[-1]}): one forward = launches {first}..{first + per_fwd - 1}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 2 --warmup 3 "
                "--no-cpu-baseline`; serialised, cold-cache launches: compare SHARES with the live step.\n\n")
        f.write(f"Serialised total: {total:.0f} us per forward, {n} launches.\n\n| kernel | launches | us per forward | share |\n|---|---:|---:|---:|\n")
        for key, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{key}` | {c} | {t:.1f} | {100 * t / total:.1f} % |\n")
        gemm = sum(t for k, (c, t) in agg.items() if k.startswith("gemm_tap_kernel"))
        f.write(f"\n`gemm_tap_kernel` (all instantiations): {gemm:.0f} us = {100 * gemm / total:.1f} % of the serialised forward; "
                "template arguments are <BN, MT, TF32, EPI, ROW32, PAIR>.\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 5)
