"""U-Net forward parity vs the CPU oracle at awkward widths, with per-layer errors for the first failing one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ConditionedUnet, weights as W
from oracle import cases, ds_oracle as O

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

cfg = W.UNET_DEPLOYED
sd = W.unet_random_state_dict(cfg, seed=0)
net = ConditionedUnet(**{k: v for k, v in cfg.items() if k not in ("out_dim", "time_dim")}, device="cuda")
net.load_state_dict(sd)
shown = False
for Wd in (int(a) for a in (sys.argv[1:] or ["30", "20", "36", "44", "17", "33"])):
    x = cases.randn((1, 4, 128, Wd), 11) * 1.5
    t = torch.tensor([947], dtype=torch.long)
    cond = cases.randn((1, 512), 12)
    taps, ref_taps = {}, {}
    eps = net.forward(x.cuda(), t.cuda(), cond.cuda(), taps=taps).cpu()
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, cond, ref_taps)
    e = rel(eps, ref)
    print(f"W={Wd}: eps rel-L2 {e:.3e}", flush=True)
    if e > 1e-2 and not shown:
        shown = True
        for k in ref_taps:
            if k in taps and taps[k].shape == ref_taps[k].shape:
                print(f"    {k:16s} {rel(taps[k], ref_taps[k]):.3e}  {tuple(taps[k].shape)}")
            elif k in taps:
                print(f"    {k:16s} shapes {tuple(taps[k].shape)} vs {tuple(ref_taps[k].shape)}")
