"""CPU oracle for the NOMAD hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product
(``nomad_b200``) never does: it fails loudly when the CUDA library is missing.

This is a plain ``torch`` (CPU, fp32 or fp64) restatement of the arithmetic the
reference delegates to fairseq's ``Wav2Vec2Model`` (not vendored in the
reference; ``requirements.txt:4`` asks for ``fairseq>=0.12.2``, no exact pin)
and to ``scipy.spatial.distance.cdist``.  Each function cites the reference
line it follows and, for the un-vendored fairseq arithmetic, the in-image
mirror ``torchaudio/models/wav2vec2/components.py`` (``$TA``), which is
bit-identical to fairseq's architecture for wav2vec 2.0 base.

PARITY PINNING: the reference ships no test vectors for this path except the
README table (``README.md:69-81``), which needs the real checkpoint (absent,
no network).  The oracle is instead pinned against outputs of the reference's
own ``src/nomad_audio/nomad.py`` executed *verbatim* under import shims
(``oracle/ref_shims.py``) in the build container; ``oracle/make_golden.py``
commits those outputs under ``tests/golden/`` and ``tests/test_oracle.py``
checks this file against them.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

CONV_KERNELS = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDES = (5, 2, 2, 2, 2, 2, 2)
NUM_LAYERS = 12
NUM_HEADS = 12
POS_KERNEL = 128
POS_GROUPS = 16
P = "ssl_model."


class _GradMultiply(torch.autograd.Function):
    """fairseq ``modules/grad_multiply.py`` (upstream): identity fwd, scale bwd."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


def _w(sd, name, dtype):
    return sd[name].to(dtype)


def conv_feature_encoder(sd, wav: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B, N) waveform -> (B, T, 512).

    fairseq ``ConvFeatureExtractionModel`` (mode "default"): 7 bias-free convs,
    GroupNorm(512, 512) after conv 0 only, exact-erf GELU after every conv.
    Mirror: ``$TA/components.py:69-143`` (block), ``:561-585`` (shapes/norm).
    Called from reference ``nomad.py:226`` / ``:245``.
    """
    x = wav.to(dtype).unsqueeze(1)
    for i, (k, s) in enumerate(zip(CONV_KERNELS, CONV_STRIDES)):
        x = F.conv1d(x, _w(sd, P + f"feature_extractor.conv_layers.{i}.0.weight", dtype), stride=s)
        if i == 0:
            x = F.group_norm(
                x, 512,
                _w(sd, P + "feature_extractor.conv_layers.0.2.weight", dtype),
                _w(sd, P + "feature_extractor.conv_layers.0.2.bias", dtype), eps=1e-5)
        x = F.gelu(x)
    return x.transpose(1, 2)


def pos_conv_weight(sd, dtype=torch.float32) -> torch.Tensor:
    """weight_norm(dim=2) fold, ``$TA/components.py:212``: w = g * v / ||v||_{(0,1)}."""
    v = _w(sd, P + "encoder.pos_conv.0.weight_v", torch.float64)
    g = _w(sd, P + "encoder.pos_conv.0.weight_g", torch.float64)
    return (g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()).to(dtype)


def encoder_layers(sd, feats: torch.Tensor, dtype=torch.float32) -> List[torch.Tensor]:
    """(B, T, 512) conv features -> list of the 12 post-LN layer outputs (B, T, 768).

    fairseq ``Wav2Vec2Model.forward(mask=False, features_only=True)`` after the
    conv encoder: LayerNorm(512) -> post_extract_proj (``$TA:179-181``) ->
    ``x + GELU(pos_conv(x))[:T]`` (``$TA:228-234``) -> encoder LayerNorm
    (``$TA:421-428,756``) -> 12 post-LN layers (``$TA:384-401``).
    """
    x = feats.to(dtype)
    x = F.layer_norm(x, (512,), _w(sd, P + "layer_norm.weight", dtype), _w(sd, P + "layer_norm.bias", dtype), 1e-5)
    x = F.linear(x, _w(sd, P + "post_extract_proj.weight", dtype), _w(sd, P + "post_extract_proj.bias", dtype))
    B, T, C = x.shape
    pc = F.conv1d(x.transpose(1, 2), pos_conv_weight(sd, dtype), _w(sd, P + "encoder.pos_conv.0.bias", dtype),
                  padding=POS_KERNEL // 2, groups=POS_GROUPS)[..., :T]
    x = x + F.gelu(pc).transpose(1, 2)
    x = F.layer_norm(x, (C,), _w(sd, P + "encoder.layer_norm.weight", dtype),
                     _w(sd, P + "encoder.layer_norm.bias", dtype), 1e-5)
    outs = []
    hd = C // NUM_HEADS
    for l in range(NUM_LAYERS):
        q_ = P + f"encoder.layers.{l}."
        lin = lambda t, n: F.linear(t, _w(sd, q_ + n + ".weight", dtype), _w(sd, q_ + n + ".bias", dtype))
        q = lin(x, "self_attn.q_proj").view(B, T, NUM_HEADS, hd).transpose(1, 2) * (hd ** -0.5)
        k = lin(x, "self_attn.k_proj").view(B, T, NUM_HEADS, hd).transpose(1, 2)
        v = lin(x, "self_attn.v_proj").view(B, T, NUM_HEADS, hd).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, C)
        x = x + lin(a, "self_attn.out_proj")
        x = F.layer_norm(x, (C,), _w(sd, q_ + "self_attn_layer_norm.weight", dtype),
                         _w(sd, q_ + "self_attn_layer_norm.bias", dtype), 1e-5)
        h = lin(F.gelu(lin(x, "fc1")), "fc2")
        x = F.layer_norm(x + h, (C,), _w(sd, q_ + "final_layer_norm.weight", dtype),
                         _w(sd, q_ + "final_layer_norm.bias", dtype), 1e-5)
        outs.append(x)
    return outs


def head(x_last: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean over T -> ReLU -> Linear(768, 256) -> L2 normalise (``nomad.py:228-230``)."""
    x = torch.mean(x_last, 1)
    x = F.linear(F.relu(x), w.to(x_last.dtype), b.to(x_last.dtype))
    return F.normalize(x, dim=1)


def ssl_layers(sd, wav: torch.Tensor, dtype=torch.float32, feature_grad_mult: float = 1.0) -> List[torch.Tensor]:
    if wav.dim() == 3:
        wav = wav.squeeze(1)  # nomad.py:225 / :244
    feats = conv_feature_encoder(sd, wav, dtype)
    if feature_grad_mult != 1.0:
        feats = _GradMultiply.apply(feats, feature_grad_mult)
    return encoder_layers(sd, feats, dtype)


def embed(sd, wav: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """``TripletModel.forward`` (``nomad.py:224-231``): (B, N) or (B, 1, N) -> (B, 256)."""
    layers = ssl_layers(sd, wav, dtype)
    return head(layers[-1], sd["embedding_layer.1.weight"], sd["embedding_layer.1.bias"])


def embed_each(sd, waves: Sequence[torch.Tensor], dtype=torch.float32) -> torch.Tensor:
    """The reference's per-file, batch-1 loop (``nomad.py:166-189``): this is what
    "masked / variable length" batching has to equal."""
    with torch.no_grad():
        return torch.cat([embed(sd, w.reshape(1, -1), dtype) for w in waves], 0)


def lossnet_layers(sd, head_w, head_b, wav, dtype=torch.float32, feature_grad_mult: float = 1.0):
    """``LossNetLayers.forward`` (``nomad.py:243-258``): 12 layer outputs + head output."""
    layers = ssl_layers(sd, wav, dtype, feature_grad_mult)
    return layers + [head(layers[-1], head_w, head_b)]


def nomad_loss(ref_list, test_list) -> torch.Tensor:
    """``NomadLoss.forward`` (``nomad.py:267-282``): sum of 13 mean-L1 terms."""
    tot = 0.0
    for r, t in zip(ref_list, test_list):
        tot = tot + F.l1_loss(t, r)
    return tot


def nomad_forward(sd, head_w, head_b, estimate, clean, dtype=torch.float32, feature_grad_mult: float = 1.0):
    """``Nomad.forward`` (``nomad.py:142-146``)."""
    est = lossnet_layers(sd, head_w, head_b, estimate, dtype, feature_grad_mult)
    cl = lossnet_layers(sd, head_w, head_b, clean, dtype, feature_grad_mult)
    return nomad_loss(cl, est)


def cdist_mean(test: np.ndarray, nmr: np.ndarray):
    """``scipy.spatial.distance.cdist`` (Euclidean; float64 direct differences) +
    ``np.mean(axis=1)`` (``nomad.py:108,111``)."""
    a = np.asarray(test, dtype=np.float64)
    b = np.asarray(nmr, dtype=np.float64)
    dm = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    step = max(1, (1 << 22) // max(1, b.shape[0] * a.shape[1]))
    for i in range(0, a.shape[0], step):
        d = a[i:i + step, None, :] - b[None, :, :]
        dm[i:i + step] = np.sqrt(np.einsum("nmd,nmd->nm", d, d))
    return dm, dm.mean(axis=1)


def stem(path: str) -> str:
    """``x.split('/')[-1].split('.')[0]`` (``nomad.py:114,120``)."""
    return path.split("/")[-1].split(".")[0]


def flops_embed(n_samples: int) -> float:
    """Algorithmic forward FLOPs F(N) for one utterance."""
    t = int(n_samples)
    T = []
    for k, s in zip(CONV_KERNELS, CONV_STRIDES):
        t = (t - k) // s + 1
        T.append(t)
    T6 = T[6]
    return (2.0 * (5120 * T[0] + 786432 * (T[1] + T[2] + T[3] + T[4]) + 524288 * (T[5] + T[6]))
            + 786432.0 * T6 + 9437184.0 * T6
            + 12.0 * (4718592.0 * T6 + 9437184.0 * T6 + 3072.0 * T6 * T6) + 393216.0)


def resample_kernel_bank(orig_freq: int, new_freq: int):
    """torchaudio ``_get_sinc_resample_kernel`` restated in numpy float64 (sinc-Hann, width 6, rolloff 0.99),
    including the float32 rounding of ``-p / new`` that its int64-arange / int division produces.
    -> (float32 (new, K) bank, width, orig, new) with the rates reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    lpw, rolloff = 6, 0.99
    base = min(o, n) * rolloff
    width = math.ceil(lpw * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    tp = (np.arange(0, -n, -1, dtype=np.int64) / np.float32(n)).astype(np.float32).astype(np.float64)[:, None]
    t = np.clip((tp + idx) * base, -lpw, lpw)
    window = np.cos(t * math.pi / lpw / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return (k * window * (base / o)).astype(np.float32), width, o, n


def load_processing_pcm16(pcm: np.ndarray, sr: int, target_sr: int = 16000, trim: bool = False) -> np.ndarray:
    """``Nomad.load_processing`` (``nomad.py:192-212``) from decoded 16-bit PCM (n_frames, channels): scale, mean of
    the first two channels, torchaudio-default ``Resample``, optional 10 s trim.  -> float32 (N,)."""
    x = np.asarray(pcm, dtype=np.float32) / np.float32(32768.0)
    if x.ndim == 1:
        x = x[:, None]
    x = (x[:, 0] + x[:, 1]) / np.float32(2) if x.shape[1] > 1 else x[:, 0]
    if sr != target_sr:
        bank, width, o, n = resample_kernel_bank(sr, target_sr)
        length = x.shape[0]
        xp = np.concatenate([np.zeros(width, np.float32), x, np.zeros(width + o, np.float32)])
        groups = (xp.shape[0] - bank.shape[1]) // o + 1
        win = np.lib.stride_tricks.sliding_window_view(xp, bank.shape[1])[::o][:groups]
        y = (win.astype(np.float64) @ bank.T.astype(np.float64)).astype(np.float32).reshape(-1)
        x = y[: -(-n * length // o)]
    if trim and x.shape[0] > target_sr * 10:
        x = x[: target_sr * 10]
    return x
