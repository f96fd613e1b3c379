"""One CFG-doubled U-Net evaluation at the benchmark batch (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusynth_b200 import ConditionedUnet, weights as W
B = int(os.environ.get("B", "64"))
net = ConditionedUnet(**{k: v for k, v in W.UNET_DEPLOYED.items() if k not in ("out_dim", "time_dim")}, device="cuda")
net.load_state_dict(W.unet_random_state_dict(seed=0, perturb_norm=False))
pl = net.plan(2 * B, 128, 64, x_batch_mod=B, uniform_time=True)
pl.x.normal_(); pl.cond.normal_(); pl.t.fill_(500)
pl.run_cond()
for _ in range(int(os.environ.get("REPS", "2"))):
    pl.run()
torch.cuda.synchronize()
print("ok", pl.num_launches(), float(pl.eps.abs().mean()))
