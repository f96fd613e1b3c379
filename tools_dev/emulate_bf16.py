"""Design-time numerics model of the CUDA U-Net (bf16 storage + bf16 MMA operands, fp32 accumulate,
GroupNorm(1,C) folded into the following conv).  Predicts parity error vs the fp32 oracle on CPU."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import cases, ds_oracle as O

RES_FP32 = os.environ.get("RES_FP32", "0") == "1"
RA = os.environ.get("RA", "bf16"); RW = os.environ.get("RW", "bf16")
def _rnd(x, kind):
    if kind == "bf16": return x.bfloat16().float()
    if kind == "fp16": return x.half().float()
    return x
def r(x): return _rnd(x, RA)
def rw(x): return _rnd(x, RW)
def rs(x): return x if RES_FP32 else r(x)      # residual stream storage

def gn_fold_conv(h_b, sd, pn, pc, pad):
    # h_b: bf16-stored raw input; stats from it
    w, b = sd[pc + "weight"], sd.get(pc + "bias")
    g, be = sd[pn + "weight"], sd[pn + "bias"]
    mu = h_b.mean(dim=(1, 2, 3), keepdim=True); var = h_b.var(dim=(1, 2, 3), keepdim=True, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    wg = rw(w * g.view(1, -1, 1, 1))
    acc = F.conv2d(h_b, wg, None, padding=pad)
    ones = torch.ones_like(h_b[:1, :1])
    e1 = F.conv2d(ones.expand(1, w.shape[1], -1, -1), wg, None, padding=pad)
    e2 = F.conv2d((be.view(1, -1, 1, 1) * ones).expand(1, w.shape[1], -1, -1), w, b, padding=pad)
    return rstd * acc - rstd * mu * e1 + e2

def block(sd, p, x, temb):
    xb = r(x) if RES_FP32 else x
    h = F.conv2d(xb, sd[p + "ds_conv.weight"], sd[p + "ds_conv.bias"], padding=3, groups=x.shape[1])
    if temb is not None and (p + "mlp.1.weight") in sd:
        h = h + F.linear(F.gelu(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])[:, :, None, None]
    h = r(h)
    y = r(F.gelu(gn_fold_conv(h, sd, p + "net.0.", p + "net.1.", 1)))
    o = gn_fold_conv(y, sd, p + "net.3.", p + "net.4.", 1)
    if (p + "res_conv.weight") in sd:
        res = F.conv2d(xb, rw(sd[p + "res_conv.weight"]), sd[p + "res_conv.bias"])
    else:
        res = x
    return rs(o + res)

def attn(sd, p, x, cemb, heads=4, dh=32):
    B, C, H, W = x.shape; n = H * W
    xb = r(x)
    qkv = gn_fold_conv(xb, sd, p + "fn.norm.", p + "fn.fn.to_qkv.", 0).reshape(B, 3, heads, dh, n)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    k = k + F.linear(cemb, sd[p + "fn.fn.label_key.weight"], sd[p + "fn.fn.label_key.bias"]).view(B, heads, dh, 1)
    q = q + F.linear(cemb, sd[p + "fn.fn.label_query.weight"], sd[p + "fn.fn.label_query.bias"]).view(B, heads, dh, 1)
    q = r(q.softmax(dim=-2) * dh ** -0.5)
    k = r(k); v = r(v)
    ctx = torch.einsum("bhdn,bhen->bhde", k.softmax(dim=-1), v)
    wo = sd[p + "fn.fn.to_out.0.weight"].view(C, heads, dh)
    M = rw(torch.einsum("che,bhde->bchd", wo, ctx).reshape(B, C, heads * dh))
    y = torch.einsum("bck,bkn->bcn", M, q.reshape(B, heads * dh, n)).reshape(B, C, H, W) + sd[p + "fn.fn.to_out.0.bias"].view(1, -1, 1, 1)
    y = r(y)
    return rs(F.group_norm(y, 1, sd[p + "fn.fn.to_out.1.weight"], sd[p + "fn.fn.to_out.1.bias"], 1e-5) + x)

def unet(sd, x, t, cond):
    n_stage = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    n_midl = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("mid_left."))
    cemb = F.linear(cond, sd["label_embedding.embedding.weight"], sd["label_embedding.embedding.bias"])
    hs = []
    x = rs(F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)); hs.append(x)
    temb = O.time_embedding(sd, t, sd["init_conv.weight"].shape[0])
    conv = lambda x, p, **kw: rs(F.conv2d(r(x), rw(sd[p + "weight"]), sd[p + "bias"], **kw))
    for i in range(n_stage):
        p = f"downs.{i}."
        x = block(sd, p + "0.", x, temb); x = attn(sd, p + "1.", x, cemb); hs.append(x)
        x = block(sd, p + "2.", x, temb); x = attn(sd, p + "3.", x, cemb); hs.append(x)
        x = conv(x, p + "4.", stride=2, padding=1); hs.append(x)
    for j in range(n_midl):
        x = block(sd, f"mid_left.{j}.", x, temb); hs.append(x)
    x = block(sd, "mid_mid.0.", x, temb); x = attn(sd, "mid_mid.1.", x, cemb); x = block(sd, "mid_mid.2.", x, temb)
    for j in range(n_midl):
        x = block(sd, f"mid_right.{j}.", torch.cat([hs.pop(), x], 1), temb)
    for i in range(n_stage):
        p = f"ups.{i}."
        x = block(sd, p + "0.", torch.cat([hs.pop(), x], 1), temb); x = attn(sd, p + "1.", x, cemb)
        x = rs(F.conv_transpose2d(r(x), rw(sd[p + "2.weight"]), sd[p + "2.bias"], stride=2, padding=1))
        x = block(sd, p + "3.", torch.cat([hs.pop(), x], 1), temb); x = attn(sd, p + "4.", x, cemb)
        x = block(sd, p + "5.", torch.cat([hs.pop(), x], 1), temb); x = attn(sd, p + "6.", x, cemb)
    x = block(sd, "final_conv.0.", torch.cat([hs.pop(), x], 1), None)
    return F.conv2d(r(x), rw(sd["final_conv.1.weight"]), sd["final_conv.1.bias"], padding=1)

if __name__ == "__main__":
    torch.set_num_threads(8)
    name = sys.argv[1] if len(sys.argv) > 1 else "deployed_w64"
    cfg, sd, x, t, cond = cases.unet_case(name)
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, cond)
        em = unet(sd, x, t, cond)
    print(name, "RA", RA, "RW", RW, "eps rel-L2", float((em - ref).norm() / ref.norm()))
    # CFG-combined one-step latent error at t=999 (first sampling step), two conditions
    B = 1
    from diffusynth_b200 import weights as W
    import numpy as np
    cond, uncond = W.synthetic_conditions(B, 512)
    x0 = cases.randn((B, 4, 128, 64), 3)
    xin = torch.cat([x0, x0]); tt = torch.full((2,), 999, dtype=torch.long); cc = torch.cat([uncond[None], cond])
    with torch.no_grad():
        er = O.unet_forward(sd, xin, tt, cc); ee = unet(sd, xin, tt, cc)
    sch = O.Schedule(1000); sch.respace(list(np.linspace(0, 999, 4, dtype=np.int32)))
    co = O.ddim_coefficients(sch, 3, 0.0)
    a = O.ddim_update(x0, er[:1], er[1:], 6, co, torch.zeros_like(x0)); b = O.ddim_update(x0, ee[:1], ee[1:], 6, co, torch.zeros_like(x0))
    print("   eps_u", float((ee[:1]-er[:1]).norm()/er[:1].norm()), "eps_c", float((ee[1:]-er[1:]).norm()/er[1:].norm()), "latent(CFG=6)", float((a-b).norm()/a.norm()),
          "  |eps_c-eps_u|/|eps|", float((er[1:]-er[:1]).norm()/er[1:].norm()))
