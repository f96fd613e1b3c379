"""Import the UNMODIFIED reference from /root/reference (read-only), for pinning the oracle
and minting golden vectors.  TEST INFRASTRUCTURE ONLY; works only where the reference exists
(this container, not the GPU box).

The reference's ``model/diffusion.py`` imports ``metrics.IS`` (absent from the reference repo)
and ``tools`` imports matplotlib/librosa (absent from this image); empty stand-in modules are
registered before import, nothing in the reference is edited."""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DIFFUSYNTH_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "DiffSynthSampler.py"))


def load():
    """Returns a namespace with the reference classes/functions used on the path."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for m in ("matplotlib", "matplotlib.pyplot", "librosa", "metrics", "metrics.IS"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["metrics.IS"].get_inception_score = lambda *a, **k: None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from model.DiffSynthSampler import DiffSynthSampler
        from model.diffusion import ConditionedUnet
        from model.VQGAN import VQGAN
        import tools
    ns = types.SimpleNamespace(DiffSynthSampler=DiffSynthSampler, ConditionedUnet=ConditionedUnet,
                               VQGAN=VQGAN, tools=tools)
    return ns


def feed_noise(sampler, draws):
    """Make a reference sampler consume host-generated noise: every internal
    ``randn(max_batchsize, C, H, train_width)`` draw (DiffSynthSampler.py:111) is replaced by
    the next entry of ``draws`` (draw 0 must be passed as ``initial_noise=``)."""
    orig = sampler.get_deterministic_noise_tensor_repeat
    state = {"k": 1}

    def patched(batchsize, width, reference_noise=None):
        if reference_noise is None:
            reference_noise = draws[state["k"]][:batchsize]
            state["k"] += 1
        return orig(batchsize, width, reference_noise=reference_noise)

    sampler.get_deterministic_noise_tensor_repeat = patched
    return sampler


def load_glue():
    """webUI/natural_language_guided_4/utils.py, imported UNMODIFIED from its file (the webUI package itself needs gradio).
    ``librosa`` is absent from this image: the stand-in module gets ``istft`` = the oracle's restatement of
    librosa.istft, so the reference's decode loop (utils.py:194-267) runs end to end."""
    load()
    import importlib.util
    from oracle import ds_oracle as O
    sys.modules["librosa"].istft = lambda D, hop_length=256, win_length=1024: O.istft(D, hop_length, win_length)
    path = os.path.join(REFERENCE_ROOT, "webUI", "natural_language_guided_4", "utils.py")
    spec = importlib.util.spec_from_file_location("_ds_ref_glue", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
