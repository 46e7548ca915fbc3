"""NOMAD checkpoint handling: the fairseq-keyed ``TripletModel`` state_dict.

The reference saves ``TripletModel.state_dict()`` (reference
``src/training/train_triplet.py:177``) and loads it back in
``src/nomad_audio/nomad.py:63-65``.  Keys are ``ssl_model.<fairseq wav2vec2
names>`` plus ``embedding_layer.1.{weight,bias}`` (``nomad.py:219-222``).  The
file holds tensors only, so plain ``torch.load`` reads it without fairseq.

There is no network in the build/bench environment and the real checkpoint is
not shipped with the reference, so :func:`random_state_dict` emits a seeded
random-init state_dict of the *identical* architecture (wav2vec 2.0 base).
Timing is weight independent; parity is checked oracle-vs-kernel on the same
dict.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch

# wav2vec 2.0 base hyper-parameters (fairseq ``wav2vec_small.pt``; mirrored in
# torchaudio ``models/wav2vec2/model.py:420-437``)
CONV_KERNELS = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDES = (5, 2, 2, 2, 2, 2, 2)
CONV_DIM = 512
EMBED_DIM = 768
FFN_DIM = 3072
NUM_HEADS = 12
HEAD_DIM = 64
NUM_LAYERS = 12
POS_KERNEL = 128
POS_GROUPS = 16
SSL_OUT_DIM = 768  # nomad.py:54
EMB_DIM = 256      # nomad.py:55
NUM_LOSS_TERMS = 13  # nomad.py:264


def expected_shapes() -> "OrderedDict[str, tuple]":
    """All tensors of a NOMAD ``TripletModel`` state_dict, in file order."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    p = "ssl_model."
    s[p + "mask_emb"] = (EMBED_DIM,)
    for i, k in enumerate(CONV_KERNELS):
        cin = 1 if i == 0 else CONV_DIM
        s[p + f"feature_extractor.conv_layers.{i}.0.weight"] = (CONV_DIM, cin, k)
        if i == 0:
            s[p + "feature_extractor.conv_layers.0.2.weight"] = (CONV_DIM,)
            s[p + "feature_extractor.conv_layers.0.2.bias"] = (CONV_DIM,)
    s[p + "post_extract_proj.weight"] = (EMBED_DIM, CONV_DIM)
    s[p + "post_extract_proj.bias"] = (EMBED_DIM,)
    s[p + "encoder.pos_conv.0.bias"] = (EMBED_DIM,)
    s[p + "encoder.pos_conv.0.weight_g"] = (1, 1, POS_KERNEL)
    s[p + "encoder.pos_conv.0.weight_v"] = (EMBED_DIM, EMBED_DIM // POS_GROUPS, POS_KERNEL)
    for l in range(NUM_LAYERS):
        q = p + f"encoder.layers.{l}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[q + f"self_attn.{n}.weight"] = (EMBED_DIM, EMBED_DIM)
            s[q + f"self_attn.{n}.bias"] = (EMBED_DIM,)
        s[q + "self_attn_layer_norm.weight"] = (EMBED_DIM,)
        s[q + "self_attn_layer_norm.bias"] = (EMBED_DIM,)
        s[q + "fc1.weight"] = (FFN_DIM, EMBED_DIM)
        s[q + "fc1.bias"] = (FFN_DIM,)
        s[q + "fc2.weight"] = (EMBED_DIM, FFN_DIM)
        s[q + "fc2.bias"] = (EMBED_DIM,)
        s[q + "final_layer_norm.weight"] = (EMBED_DIM,)
        s[q + "final_layer_norm.bias"] = (EMBED_DIM,)
    s[p + "encoder.layer_norm.weight"] = (EMBED_DIM,)
    s[p + "encoder.layer_norm.bias"] = (EMBED_DIM,)
    s[p + "layer_norm.weight"] = (CONV_DIM,)
    s[p + "layer_norm.bias"] = (CONV_DIM,)
    s["embedding_layer.1.weight"] = (EMB_DIM, SSL_OUT_DIM)
    s["embedding_layer.1.bias"] = (EMB_DIM,)
    return s


def random_state_dict(seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random weights of the identical architecture (CPU, fp32).

    Scales follow fairseq's own initialisers (kaiming-normal convs, N(0, σ)
    linears, pos-conv N(0, sqrt(4/(k·C)))) with norm affine parameters and
    biases perturbed away from (1, 0) so every parameter influences the output.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)

    def normal(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in expected_shapes().items():
        leaf = name.split(".")[-1]
        if name.endswith("mask_emb"):
            t = torch.rand(shape, generator=g, dtype=torch.float32)
        elif "feature_extractor.conv_layers" in name and name.endswith(".0.weight"):
            fan_in = shape[1] * shape[2]
            t = normal(shape, math.sqrt(2.0 / fan_in))
        elif "pos_conv.0.weight_v" in name:
            t = normal(shape, math.sqrt(4.0 / (POS_KERNEL * EMBED_DIM)))
        elif "pos_conv.0.weight_g" in name:
            t = torch.ones(shape)  # placeholder: filled below once weight_v exists
        elif "layer_norm" in name or ".0.2." in name:
            t = 1.0 + normal(shape, 0.1) if leaf == "weight" else normal(shape, 0.1)
        elif leaf == "bias":
            t = normal(shape, 0.05)
        elif name == "embedding_layer.1.weight":
            t = normal(shape, 0.03)
        else:  # transformer / projection linears
            t = normal(shape, 0.04)
        sd[name] = t
    # weight_g = per-tap norm of v (what weight_norm(dim=2) initialises to), scaled
    v = sd["ssl_model.encoder.pos_conv.0.weight_v"]
    nrm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    sd["ssl_model.encoder.pos_conv.0.weight_g"] = nrm * (0.75 + 0.5 * torch.rand((1, 1, POS_KERNEL), generator=g))
    return sd


def validate_state_dict(sd) -> None:
    exp = expected_shapes()
    for name, shape in exp.items():
        if name.endswith("mask_emb"):
            continue  # unused by the scoring/loss path
        if name not in sd:
            raise KeyError(f"NOMAD checkpoint is missing tensor {name}")
        if tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"NOMAD checkpoint tensor {name} has shape {tuple(sd[name].shape)}, expected {shape}")


def load_state_dict(path: str | None = None, seed: int = 1234):
    """Load ``nomad_best_model.pt`` (``nomad.py:29,65``) if present, else seeded random init.

    Returns ``(state_dict, source)`` where source is ``"checkpoint:<path>"`` or
    ``"random-init(seed=<n>)"``.
    """
    cands = [path] if path else []
    cands += [os.environ.get("NOMAD_B200_CHECKPOINT", ""), os.path.join("pt-models", "nomad_best_model.pt")]
    for c in cands:
        if c and os.path.isfile(c) and os.path.getsize(c) > 0:
            sd = torch.load(c, map_location="cpu")
            sd = OrderedDict((k, v.detach().to(torch.float32).contiguous()) for k, v in sd.items())
            validate_state_dict(sd)
            return sd, f"checkpoint:{c}"
    return random_state_dict(seed), f"random-init(seed={seed})"


def fold_pos_conv_weight(sd) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / ||v||, norm over (out, in) per tap."""
    v = sd["ssl_model.encoder.pos_conv.0.weight_v"].to(torch.float64)
    g = sd["ssl_model.encoder.pos_conv.0.weight_g"].to(torch.float64)
    nrm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return (g * v / nrm).to(torch.float32)


def conv_out_lengths(n_samples: int) -> list:
    """Frames after each of the 7 conv layers (torchaudio components.py:96)."""
    out = []
    t = int(n_samples)
    for k, s in zip(CONV_KERNELS, CONV_STRIDES):
        t = (t - k) // s + 1 if t >= k else 0
        out.append(t)
    return out


def flops_embed(n_samples: int) -> float:
    """Algorithmic forward FLOPs F(N) of one utterance: conv stack, projection,
    positional conv, 12 x (qkv + out-proj, FFN, attention core), head."""
    T = conv_out_lengths(n_samples)
    T6 = T[6]
    return (2.0 * (5120 * T[0] + 786432 * (T[1] + T[2] + T[3] + T[4]) + 524288 * (T[5] + T[6]))
            + 786432.0 * T6 + 9437184.0 * T6
            + 12.0 * (4718592.0 * T6 + 9437184.0 * T6 + 3072.0 * T6 * T6) + 393216.0)
