"""Run the reference's ``src/nomad_audio/nomad.py`` VERBATIM under import shims.

TEST INFRASTRUCTURE ONLY, and only usable in the build container (it reads
``/root/reference``, which does not exist on the GPU box).  Used by
``oracle/make_golden.py`` to produce the committed fixtures that pin
``oracle/w2v_oracle.py``.

Three blockers keep the reference from importing as-is:
``import fairseq`` (absent), import-time ``urlretrieve`` of two checkpoints (no
network) and ``torchaudio.load`` (needs torchcodec, absent).  The shims:

1. a stub ``fairseq`` module whose ``checkpoint_utils.load_model_ensemble_and_task``
   returns an adapter around ``torchaudio.models.wav2vec2_base()`` (the in-image
   mirror of fairseq's architecture) with fairseq's call signature
   ``forward(source, mask=False, features_only=True) -> {'x', 'layer_results'}``;
2. pre-created ``pt-models/wav2vec_small.pt`` (only ``isfile`` is checked,
   ``nomad.py:22``) and ``pt-models/nomad_best_model.pt`` holding the seeded
   state_dict (``nomad.py:30,65``), re-keyed to the adapter's parameter names;
3. ``torchaudio.load`` replaced by a stdlib ``wave`` reader (int16/32768 -> f32).
"""
from __future__ import annotations

import importlib.util
import os
import re
import sys
import types
import wave as _wave

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("NOMAD_REFERENCE_ROOT", "/root/reference")


def fairseq_to_torchaudio_key(key: str):
    """fairseq wav2vec2 parameter name -> torchaudio name (cf. torchaudio
    ``utils/import_fairseq.py:51-115``)."""
    if key == "mask_emb":
        return None
    m = re.match(r"feature_extractor\.conv_layers\.0\.2\.(weight|bias)", key)
    if m:
        return f"feature_extractor.conv_layers.0.layer_norm.{m.group(1)}"
    m = re.match(r"feature_extractor\.conv_layers\.(\d+)\.0\.(weight|bias)", key)
    if m:
        return f"feature_extractor.conv_layers.{m.group(1)}.conv.{m.group(2)}"
    m = re.match(r"post_extract_proj\.(weight|bias)", key)
    if m:
        return f"encoder.feature_projection.projection.{m.group(1)}"
    m = re.match(r"layer_norm\.(weight|bias)", key)
    if m:
        return f"encoder.feature_projection.layer_norm.{m.group(1)}"
    if key == "encoder.pos_conv.0.bias":
        return "encoder.transformer.pos_conv_embed.conv.bias"
    if key == "encoder.pos_conv.0.weight_g":
        return "encoder.transformer.pos_conv_embed.conv.parametrizations.weight.original0"
    if key == "encoder.pos_conv.0.weight_v":
        return "encoder.transformer.pos_conv_embed.conv.parametrizations.weight.original1"
    m = re.match(r"encoder\.layer_norm\.(weight|bias)", key)
    if m:
        return f"encoder.transformer.layer_norm.{m.group(1)}"
    m = re.match(r"encoder\.layers\.(\d+)\.self_attn\.((k_|v_|q_|out_)proj\.(weight|bias))", key)
    if m:
        return f"encoder.transformer.layers.{m.group(1)}.attention.{m.group(2)}"
    m = re.match(r"encoder\.layers\.(\d+)\.self_attn_layer_norm\.(weight|bias)", key)
    if m:
        return f"encoder.transformer.layers.{m.group(1)}.layer_norm.{m.group(2)}"
    m = re.match(r"encoder\.layers\.(\d+)\.fc1\.(weight|bias)", key)
    if m:
        return f"encoder.transformer.layers.{m.group(1)}.feed_forward.intermediate_dense.{m.group(2)}"
    m = re.match(r"encoder\.layers\.(\d+)\.fc2\.(weight|bias)", key)
    if m:
        return f"encoder.transformer.layers.{m.group(1)}.feed_forward.output_dense.{m.group(2)}"
    m = re.match(r"encoder\.layers\.(\d+)\.final_layer_norm\.(weight|bias)", key)
    if m:
        return f"encoder.transformer.layers.{m.group(1)}.final_layer_norm.{m.group(2)}"
    raise ValueError(f"unexpected fairseq key {key}")


class _GradMultiply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


class FairseqW2VAdapter(torch.nn.Module):
    """torchaudio wav2vec2_base behind fairseq's ``Wav2Vec2Model`` call signature."""

    def __init__(self, feature_grad_mult: float = 1.0):
        super().__init__()
        import torchaudio
        self.model = torchaudio.models.wav2vec2_base()
        self.feature_grad_mult = feature_grad_mult

    def remove_pretraining_modules(self):
        return None

    def forward(self, source, mask=False, features_only=True, padding_mask=None):
        assert not mask and features_only
        feats, _ = self.model.feature_extractor(source, None)
        if self.feature_grad_mult != 1.0:
            feats = _GradMultiply.apply(feats, self.feature_grad_mult)
        outs = self.model.encoder.extract_features(feats, None)
        return {"x": outs[-1], "layer_results": [(o.transpose(0, 1), None, None) for o in outs]}


def adapter_state_dict(sd_fairseq):
    """TripletModel state_dict (fairseq keys) -> same tensors under the adapter's names."""
    out = {}
    for k, v in sd_fairseq.items():
        if k.startswith("ssl_model."):
            nk = fairseq_to_torchaudio_key(k[len("ssl_model."):])
            if nk is not None:
                out["ssl_model.model." + nk] = v.clone()
        else:
            out[k] = v.clone()
    return out


def wave_load(filepath):
    """``torchaudio.load`` semantics for PCM16 wav: (channels, N) float32 = int16/32768."""
    with _wave.open(str(filepath), "rb") as w:
        ch, sr, n = w.getnchannels(), w.getframerate(), w.getnframes()
        assert w.getsampwidth() == 2
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(n, ch).T
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0), sr


def load_reference_nomad(sd_fairseq, workdir: str, feature_grad_mult: float = 1.0):
    """Import the reference module in ``workdir`` and return ``(module, Nomad instance)``.

    ``os.getcwd()`` is changed to ``workdir`` because the reference uses the
    cwd-relative ``./pt-models`` (``nomad.py:15-33``).
    """
    import torchaudio

    os.makedirs(os.path.join(workdir, "pt-models"), exist_ok=True)
    open(os.path.join(workdir, "pt-models", "wav2vec_small.pt"), "wb").close()
    torch.save(adapter_state_dict(sd_fairseq), os.path.join(workdir, "pt-models", "nomad_best_model.pt"))

    fs = types.ModuleType("fairseq")
    cu = types.ModuleType("fairseq.checkpoint_utils")
    cu.load_model_ensemble_and_task = lambda paths: ([FairseqW2VAdapter(feature_grad_mult)], None, None)
    fs.checkpoint_utils = cu
    sys.modules["fairseq"] = fs
    sys.modules["fairseq.checkpoint_utils"] = cu
    torchaudio.load = wave_load

    os.chdir(workdir)
    spec = importlib.util.spec_from_file_location(
        "reference_nomad", os.path.join(REFERENCE_ROOT, "src", "nomad_audio", "nomad.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(4321)  # LossNetLayers head is freshly random-initialised (nomad.py:238-241)
    inst = mod.Nomad(device="cpu")
    return mod, inst
