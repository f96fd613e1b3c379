"""Per-layer difference of one U-Net evaluation between the widest N tiling and the small-job narrow tiling (batch_invariant=False), and each against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ConditionedUnet, weights as W
from oracle import cases, ds_oracle as O

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())

cfg, sd, x, t, cond = cases.unet_case(os.environ.get("CASE", "deployed_w64"))
out = {}
for wide in (1, 0):
    net = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda", batch_invariant=bool(wide))
    net.load_state_dict(sd)
    taps = {}
    eps = net.forward(x.cuda(), t.cuda(), cond.cuda(), taps=taps).cpu()
    out[wide] = (eps, {k: v.float().cpu() for k, v in taps.items()})
ref_taps = {}
with torch.no_grad():
    ref = O.unet_forward(sd, x, t, cond, ref_taps)
print(f"eps: wide vs oracle {rel(out[1][0], ref):.3e}  narrow vs oracle {rel(out[0][0], ref):.3e}  narrow vs wide {rel(out[0][0], out[1][0]):.3e}")
for k in out[1][1]:
    if k in out[0][1]:
        d = rel(out[0][1][k], out[1][1][k])
        r = f"{rel(out[1][1][k], ref_taps[k]):.2e} / {rel(out[0][1][k], ref_taps[k]):.2e}" if k in ref_taps else ""
        print(f"  {k:18s} narrow vs wide {d:.3e}   vs oracle (wide / narrow) {r}")
