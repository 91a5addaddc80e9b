"""Summarise an `ncu --page raw --csv` dump of the convolution launches into profiles/*.md + *.json.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > profiles/rNN_x_ncu_gemm_raw.csv
    python tools/ncu_summary.py profiles/rNN_x_ncu_gemm_raw.csv profiles/rNN_x_ncu_gemm_summary "<command line>"

Launch k of the capture is matched to the k-th gemm op of the batch-256 program (stem, then the 16 block convolutions).
bench.py reads the newest profiles/*_ncu_gemm_summary.json for `roofline.traffic`.
"""
import csv
import json
import sys

NAMES = ["stem.conv+pool"] + [f"s{s}.b{b}.conv{c}" for s in (1, 2, 3, 4) for b in (0, 1) for c in (1, 2)]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(raw, out, command):
    rows = list(csv.reader(open(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}

    def val(r, key):
        i = col[key]
        return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

    launches = []
    for k, r in enumerate(data):
        kern = r[col["Kernel Name"]]
        kern = kern[kern.index("gemm_tap_kernel"):] if "gemm_tap_kernel" in kern else kern
        launches.append({
            "op": NAMES[k] if k < len(NAMES) else f"launch{k}", "kernel": kern,
            "time_us": val(r, "gpu__time_duration.sum"),
            "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_read_mb": val(r, "dram__bytes_read.sum"), "dram_write_mb": val(r, "dram__bytes_write.sum"),
            "regs": int(val(r, "launch__registers_per_thread")),
            "smem_lsu_wavefronts": int(val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
            "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    rd, wr = sum(l["dram_read_mb"] for l in launches), sum(l["dram_write_mb"] for l in launches)
    json.dump({"launches": launches, "dram_read_mb": rd, "dram_write_mb": wr,
               "source": f"ncu --set full, {len(launches)} conv launches of one forward, batch 256", "command": command},
              open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full, the {len(launches)} convolution launches of one forward (batch 256, B200, --clock-control none)\n\n")
        f.write(f"Command: `{command}`\n(raw CSV: {raw.split('/')[-1]}; per-launch times under ncu are cold-cache and serialised).\n\n")
        f.write("| op | kernel | grid x block | time us | tensor pipe active % | DRAM read MB | DRAM write MB | regs | smem LSU wavefronts |\n")
        f.write("|---|---|---|---:|---:|---:|---:|---:|---:|\n")
        for l in launches:
            f.write(f"| {l['op']} | `{l['kernel'][:32]}` | {l['grid']} x {l['block']} | {l['time_us']:.1f} | "
                    f"{l['tensor_pipe_active_pct']:.1f} | {l['dram_read_mb']:.1f} | {l['dram_write_mb']:.1f} | {l['regs']} | "
                    f"{l['smem_lsu_wavefronts']} |\n")
        tot = sum(l["time_us"] for l in launches)
        f.write(f"\nSum over the {len(launches)} launches: {tot:.0f} us; DRAM read {rd:.0f} MB + write {wr:.0f} MB = {rd + wr:.0f} MB "
                f"per 256-pair forward ({(rd + wr) / 256:.2f} MB/pair; SURVEY 8(d) algorithmic layer-wise bytes for the "
                f"convolutions: 9.07 MB/pair with every activation spilled to HBM).\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
