"""Key metrics of `ncu --set full` captures (gpurun_out/*.ncu-rep) as a markdown table: python tools_dev/summarize_ncu_full.py out.md rep1 [rep2 ...]"""
import csv, io, subprocess, sys
KEYS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "MUFU pipe %")]
out, reps = sys.argv[1], sys.argv[2:]
lines = ["# `ncu --set full --clock-control none` captures of the final round-2 code (one U-Net evaluation at batch 64 = 128 samples, `tools_dev/unet_once.py`)", "",
         "Read with `ncu -i <rep> --page raw --csv`; cold-cache, serialised (compare shares, not absolutes).", "",
         "| capture | kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
for rep in reps:
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").replace("void ", "").replace("ds::", "")[:60]
        cells = []
        for key, _ in KEYS:
            v = d.get(key, "")
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u.get(key, '')}".strip())
        lines.append(f"| {rep.split('/')[-1]} | `{name}` | " + " | ".join(cells) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
