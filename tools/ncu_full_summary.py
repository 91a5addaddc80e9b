"""Summarise an `ncu --set full ... --page raw --csv` dump of ONE whole forward (tools/profile_forward.py) into
profiles/*.md + *.json: every launch with its op name, duration, tensor-pipe activity, DRAM bytes and occupancy.

    ncu --set full --clock-control none --profile-from-start off -o /tmp/x python tools/profile_forward.py
    ncu -i /tmp/x.ncu-rep --page raw --csv > profiles/rNN_x_ncu_full_raw.csv
    python tools/ncu_full_summary.py profiles/rNN_x_ncu_full_raw.csv profiles/rNN_y_per_op_ms.json profiles/rNN_x_ncu_full_summary

Launch k of the capture is op k of the batch-256 program (names taken from a tools/per_op_ms.py dump of the same code;
the kernel names are cross-checked).  bench.py reads the newest profiles/*_ncu_gemm_summary.json / *_ncu_full_summary.json
for `roofline.traffic` (DRAM bytes of the convolution launches).
"""
import csv
import json
import sys

SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
HBM_GBS = 6459.0          # MEASURED_PEAKS.json (copy bandwidth)


def short(kern):
    kern = kern.replace("<unnamed>::", "").replace("void ", "")
    return kern[:kern.index("(")] if "(" in kern else kern


def main(raw, per_op, out):
    rows = list(csv.reader(open(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}
    names = [o["name"] for o in json.load(open(per_op))["ops"]]

    def val(r, key):
        i = col[key]
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

    launches = []
    for k, r in enumerate(data):
        kern = short(r[col["Kernel Name"]])
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        launches.append({
            "op": names[k] if k < len(names) else f"launch{k}", "kernel": kern, "time_us": t,
            "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_read_mb": rd, "dram_write_mb": wr, "dram_gbs": (rd + wr) * 1e6 / (t * 1e-6) / 1e9,
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "regs": int(val(r, "launch__registers_per_thread")),
            "smem_lsu_wavefronts": int(val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
            "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    total = sum(l["time_us"] for l in launches)
    conv = [l for l in launches if l["op"].startswith(("stem", "s1.", "s2.", "s3.", "s4.")) and "conv" in l["op"]]
    rd, wr = sum(l["dram_read_mb"] for l in conv), sum(l["dram_write_mb"] for l in conv)
    json.dump({"launches": launches, "total_us": total, "dram_read_mb": rd, "dram_write_mb": wr,
               "source": f"ncu --set full --clock-control none, all {len(launches)} launches of one forward, batch 256; "
                         f"dram_* = the {len(conv)} convolution launches (stem + 16 block convolutions)"},
              open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full, one whole forward at batch 256 ({raw.split('/')[-1]})\n\n"
                "`ncu --set full --clock-control none --profile-from-start off python tools/profile_forward.py`, exported with "
                "`--page raw --csv`.  Under ncu every launch runs alone, serialised and with cold caches: use the SHARES and the "
                "per-kernel ratios, not the absolute step time.  `HBM %` = DRAM read + write bytes / duration against the "
                f"measured {HBM_GBS:.0f} GB/s copy bandwidth (MEASURED_PEAKS.json).\n\n"
                "| # | op | kernel | us | share | tensor pipe active % | DRAM rd MB | DRAM wr MB | GB/s | HBM % | warps active % | regs | grid |\n"
                "|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for k, l in enumerate(launches):
            f.write(f"| {k} | {l['op']} | `{l['kernel']}` | {l['time_us']:.1f} | {100 * l['time_us'] / total:.1f} % | "
                    f"{l['tensor_pipe_active_pct']:.1f} | {l['dram_read_mb']:.1f} | {l['dram_write_mb']:.1f} | {l['dram_gbs']:.0f} | "
                    f"{100 * l['dram_gbs'] / HBM_GBS:.0f} | {l['warps_active_pct']:.0f} | {l['regs']} | {l['grid']} |\n")
        f.write(f"\nSum {total:.1f} us over {len(launches)} launches.  Convolutions (stem + 16): "
                f"{sum(l['time_us'] for l in conv):.1f} us, DRAM {rd:.0f} MB read + {wr:.0f} MB written per 256-pair forward.\n\n")
        grp = {}
        for l in launches:
            n = l["op"]
            key = ("stem" if n.startswith("stem") else n[:2] + " convolutions" if n[:2] in ("s1", "s2", "s3", "s4") and "conv" in n
                   else "stage tails" if n.endswith(".tail") else "ingest" if n == "ingest" else "text / fusion / head")
            grp[key] = grp.get(key, 0.0) + l["time_us"]
        f.write("| group | us | share |\n|---|---:|---:|\n")
        for k, v in grp.items():
            f.write(f"| {k} | {v:.1f} | {100 * v / total:.1f} % |\n")


if __name__ == "__main__":
    main(*sys.argv[1:4])
