"""Pins the text-front-end restatement (oracle.ds_oracle.clap_text_features): the CLAP tower against the installed transformers'
own ClapModel (third-party, unpinned in the reference's requirements.txt:5), the ProjectionHead against the reference's class
(model/multimodal_model.py:35-47) where the reference tree exists and against a golden vector minted from it everywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import ds_oracle as O
from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text.npz")
SMALL = dict(vocab_size=120, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
             max_position_embeddings=40, pad_token_id=1, layer_norm_eps=1e-12, projection_dim=64)


def _ids():
    ids = torch.tensor([[0, 17, 33, 99, 5, 2, 1, 1], [0, 8, 2, 1, 1, 1, 1, 1], [0, 50, 51, 52, 53, 54, 55, 2]])
    return ids, ids.ne(1).long()


def _hf_sd(model):
    sd = {"text_encoder." + k: v.detach().clone() for k, v in model.state_dict().items() if k.startswith(("text_model.", "text_projection."))}
    return sd


def test_tower_matches_transformers_clap():
    transformers = pytest.importorskip("transformers")
    from transformers import ClapConfig, ClapModel
    torch.manual_seed(0)
    cfg = ClapConfig(text_config={k: v for k, v in SMALL.items()}, projection_dim=SMALL["projection_dim"])
    cfg.text_config.projection_dim = SMALL["projection_dim"]
    model = ClapModel(cfg).eval()
    ids, mask = _ids()
    with torch.no_grad():
        out = model.get_text_features(input_ids=ids, attention_mask=mask)
    want = out.pooler_output if hasattr(out, "pooler_output") else out          # transformers >= 5 returns the output object
    got = O.clap_text_features(_hf_sd(model), ids, mask, heads=SMALL["num_attention_heads"], num_projection_layers=0)
    assert tuple(got.shape) == (3, SMALL["projection_dim"])
    assert float((got - want).abs().max()) < 2e-6


def _head_sd(seed=3, dim_in=64, dim=64, layers=2):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(layers):
        p = f"text_projection.layers.{i}."
        di = dim_in if i == 0 else dim
        sd[p + "projection.weight"] = torch.randn((dim, di), generator=g) * di ** -0.5
        sd[p + "projection.bias"] = 0.1 * torch.randn((dim,), generator=g)
        sd[p + "fc.weight"] = torch.randn((dim, dim), generator=g) * dim ** -0.5
        sd[p + "fc.bias"] = 0.1 * torch.randn((dim,), generator=g)
        sd[p + "layer_norm.weight"] = 1 + 0.1 * torch.randn((dim,), generator=g)
        sd[p + "layer_norm.bias"] = 0.1 * torch.randn((dim,), generator=g)
    return sd


def _head_only(sd, y, layers=2):
    """The ProjectionHead part of the restatement on given CLAP features."""
    import torch.nn.functional as F
    for j in range(layers):
        p = f"text_projection.layers.{j}."
        projected = F.linear(y, sd[p + "projection.weight"], sd[p + "projection.bias"])
        y = F.linear(F.gelu(projected), sd[p + "fc.weight"], sd[p + "fc.bias"]) + projected
        y = F.layer_norm(y, (y.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    return y


def test_projection_head_golden():
    """tests/golden/text.npz was minted from the reference's ProjectionHead (oracle/make_golden.py --text)."""
    z = np.load(GOLD)
    sd = _head_sd()
    y = torch.from_numpy(z["features"])
    got = _head_only(sd, y)
    assert float((got - torch.from_numpy(z["head_out"])).abs().max()) < 2e-6


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_projection_head_matches_reference_live():
    import sys, types
    ref_loader.load()
    sys.modules.setdefault("model.timbre_encoder_pretrain", types.ModuleType("model.timbre_encoder_pretrain"))
    if not hasattr(sys.modules["model.timbre_encoder_pretrain"], "get_timbre_encoder"):
        sys.modules["model.timbre_encoder_pretrain"].get_timbre_encoder = lambda *a, **k: None
    from model.multimodal_model import ProjectionHead
    head = ProjectionHead(embedding_dim=64, projection_dim=64, dropout=0.1, num_layers=2).eval()
    sd = _head_sd()
    head.load_state_dict({k.replace("text_projection.", ""): v for k, v in sd.items()})
    y = torch.nn.functional.normalize(torch.randn((5, 64), generator=torch.Generator().manual_seed(1)), dim=-1)
    with torch.no_grad():
        want = head(y)
    assert float((_head_only(sd, y) - want).abs().max()) < 2e-6
