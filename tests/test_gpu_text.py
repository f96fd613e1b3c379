"""Text-conditioning front end on the GPU (CLAP text tower + ProjectionHead, batched over distinct prompts)
against the oracle restatement (pinned to transformers' ClapModel and the reference's ProjectionHead in tests/test_oracle_text.py).
Tolerance: relative L2 <= 1e-2 (16-bit GEMM operands, fp32 accumulate; LayerNorm / soft-max / tail in fp32)."""
import pytest
import torch

from oracle import ds_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _prompts(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 50000, (B, L), generator=g)
    ids[:, 0] = 0
    lens = torch.randint(3, L + 1, (B,), generator=g)
    lens[0] = L
    for b in range(B):
        ids[b, lens[b] - 1] = 2
        ids[b, lens[b]:] = 1
    return ids, ids.ne(1).long()


@pytest.mark.parametrize("layers", [2, 0])
def test_text_features_match_oracle(layers):
    from diffusynth_b200.text import CLAP_TEXT, TextEncoder, text_random_state_dict
    sd = text_random_state_dict(CLAP_TEXT, num_projection_layers=layers, seed=7)
    enc = TextEncoder(num_projection_layers=layers, device="cuda").load_state_dict(sd)
    ids, mask = _prompts(5, 12, 1)
    got = enc.get_text_features(ids, mask)
    hid = enc._plan(5, 12)["hidden"].float().reshape(5, 12, 768).cpu()
    taps = {}
    want = O.clap_text_features(sd, ids, mask, num_projection_layers=layers, taps=taps)
    valid = mask.bool()
    r_h = rel(hid[valid], taps["last_hidden_state"][valid])
    r = rel(got, want)
    print(f"\nprojection layers {layers}: last hidden state rel-L2 {r_h:.2e}, text features [5,512] rel-L2 {r:.2e}")
    assert tuple(got.shape) == (5, 512) and torch.isfinite(got).all()
    assert r_h < 1e-2 and r < 1e-2


def test_text_batch_of_distinct_prompts():
    """64 distinct prompts in one pass equal the same prompts encoded one by one (the reference encodes one prompt and repeats it)."""
    from diffusynth_b200.text import TextEncoder, text_random_state_dict
    sd = text_random_state_dict(seed=7)
    enc = TextEncoder(device="cuda").load_state_dict(sd)
    ids, mask = _prompts(64, 16, 2)
    allp = enc.get_text_features(ids, mask)
    one = torch.cat([enc.get_text_features(ids[b:b + 1], mask[b:b + 1]) for b in (0, 7, 63)])
    r = rel(allp[[0, 7, 63]], one)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        enc.get_text_features(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    print(f"\n64 prompts x 16 tokens: {e0.elapsed_time(e1) / 5:.2f} ms per batch ({enc.num_launches(64, 16)} launches); batched vs single rel-L2 {r:.2e}")
    assert r < 2e-3


def test_generate_from_tokens():
    """text2sound.py:89-134 with the text front end on the device: B distinct tokenised prompts + one negative prompt -> waveforms."""
    from diffusynth_b200 import TextToTimbre
    from diffusynth_b200.text import TextEncoder, text_random_state_dict
    enc = TextEncoder(device="cuda").load_state_dict(text_random_state_dict(seed=7))
    pipe = TextToTimbre.random_init(device="cuda", seed=0)
    ids, mask = _prompts(3, 10, 5)
    nid, nmask = _prompts(1, 4, 6)
    out = pipe.generate_from_tokens(enc, ids, mask, nid, nmask, steps=2, cfg_scale=6.0, width=64, seed=3)
    cond = enc.get_text_features(ids, mask)
    uncond = enc.get_text_features(nid, nmask)[0]
    ref = pipe.generate(cond, uncond, steps=2, cfg_scale=6.0, width=64, seed=3)
    assert tuple(out.waveforms.shape) == (3, 65280) and torch.isfinite(out.waveforms).all()
    assert torch.equal(out.waveforms, ref.waveforms)
