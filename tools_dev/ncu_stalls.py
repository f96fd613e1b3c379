"""Stall-sample summary of the epilogue region (first LDTM .. last 256-bit store) of a conv_gemm kernel in an ncu report:
python tools_dev/ncu_stalls.py gpurun_out/prof_conv1b.ncu-rep"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:90])
hdr = rows[1]
si, ai, ei = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ci = {h: hdr.index(h) for h in cols}
data = [r for r in rows[2:] if len(r) > max(ci.values())]
tot = sum(int(r[ai] or 0) for r in data)
ld = [i for i, r in enumerate(data) if "LDTM" in r[si]]
st = [i for i, r in enumerate(data) if "STG.E" in r[si] and ".256" in r[si]]
reg = data[ld[0] - 5:st[-1] + 40]
c = collections.Counter()
for r in reg:
    for h in cols:
        c[h] += int(r[ci[h]] or 0)
S = sum(c.values())
print(f"epilogue region: {S} of {tot} samples, {len(reg)} instructions")
for h, v in c.most_common(10):
    print(f"  {h:24s} {v:6d} {100 * v / S:5.1f}%")
for r in sorted(reg, key=lambda r: -int(r[ai] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    top = max(cols, key=lambda h: int(r[ci[h]] or 0))
    print(f"  {r[ai]:>5s} {top:20s} {r[si].strip()[:80]}")
