"""Triplet fine-tuning step of the reference (``src/training/train_triplet.py:92-133``) on the B200 path.

``Training.train`` runs, per batch of (anchor, positive, negative) waveforms: three forwards of ``TripletModel``,
``nn.TripletMarginLoss(margin)``, ``loss.backward()`` and an Adam step over everything but the frozen conv feature
encoder (``freeze_convnet: True``), with lr 1e-5 for the pre-trained network and ``lr`` for the embedding head
(``:96-104``).  Here the three forwards, the loss and ALL parameter gradients come from one C-ABI call
(``nomad_b200_triplet_fwd_bwd``: tensor-core forward, dgrad chain, weight-gradient GEMMs); the optimiser state and the
update are host plumbing (``torch.optim.Adam`` on fp32 master tensors on the GPU), after which the kernel-ready
weights are rebuilt on the device from those tensors (``nomad_b200_refresh_weights``).

Deliberate difference from the reference: the network runs in evaluation mode (no dropout / LayerDrop: fairseq's
``Wav2Vec2Model`` applies dropout 0.1 and layerdrop 0.05 under ``model.train()``, which makes the reference's step
stochastic).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from . import _lib
from .engine import Engine, _ptr, _stream_ptr


def grad_segments():
    """[(name, offset, numel)] of the flat gradient buffer, as the library lays it out."""
    lib = _lib.load()
    out, i = [], 0
    name = C.create_string_buffer(128)
    off, cnt = C.c_int64(), C.c_int64()
    while lib.nomad_b200_triplet_grad_segment(i, name, 128, C.byref(off), C.byref(cnt)) == 0:
        out.append((name.value.decode(), int(off.value), int(cnt.value)))
        i += 1
    return out


_DEV_CACHE: Dict = {}


def _device_copy(engine: Engine, state_dict, key: str) -> torch.Tensor:
    """fp32 copy of a (host) checkpoint tensor on the engine's device, cached per source tensor version: the positional
    conv's weight_v is 19 MB, and a pageable H2D copy of it per step cost 1.5 ms."""
    t = state_dict[key]
    if t.is_cuda:
        return t.detach().to(torch.float32).clone()
    tag = (id(t), t._version, engine.device_index)
    hit = _DEV_CACHE.get(key)
    if hit is None or hit[0] != tag:
        _DEV_CACHE[key] = (tag, t.detach().to(engine.device, torch.float32))
    return _DEV_CACHE[key][1].clone()


SHAPES = {"qkv.weight": (2304, 768), "self_attn.out_proj.weight": (768, 768), "fc1.weight": (3072, 768),
          "fc2.weight": (768, 3072), "post_extract_proj.weight": (768, 512), "embedding_layer.1.weight": (256, 768)}


def triplet_loss_and_grads(engine: Engine, state_dict, anchor: torch.Tensor, positive: torch.Tensor, negative: torch.Tensor,
                           margin: float = 0.2) -> Tuple[torch.Tensor, "OrderedDict[str, torch.Tensor]"]:
    """loss (0-dim CUDA tensor) and {state_dict key: gradient} for every trainable tensor (conv encoder frozen).
    ``anchor`` / ``positive`` / ``negative``: (B, N) or (B, 1, N) waveforms of one common length."""
    lib = engine.lib
    sq = lambda t: (t.squeeze(1) if t.dim() == 3 else t).to(engine.device, torch.float32)
    wav = torch.cat([sq(anchor), sq(positive), sq(negative)]).contiguous()
    B, N = wav.shape[0] // 3, wav.shape[1]
    need = lib.nomad_b200_triplet_workspace_bytes(B, N)
    if need == 0:
        raise _lib.NomadB200Error(lib.nomad_b200_last_error().decode())
    wp, wbytes = engine._aligned(engine.workspace(need))
    n_grad = int(lib.nomad_b200_triplet_grad_floats())
    flat = torch.empty((n_grad,), dtype=torch.float32, device=engine.device)
    loss = torch.empty((), dtype=torch.float32, device=engine.device)
    scale = C.c_float()
    with torch.cuda.device(engine.device):
        _lib.check(lib.nomad_b200_triplet_fwd_bwd(engine.handle, _ptr(wav), B, N, C.c_float(margin), _ptr(loss), _ptr(flat),
                                                  C.byref(scale), wp, wbytes, _stream_ptr()), "nomad_b200_triplet_fwd_bwd")
    flat = flat / scale.value
    grads: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, off, cnt in grad_segments():
        g = flat[off:off + cnt]
        leaf = next((k for k in SHAPES if name.endswith(k)), None)
        if leaf is not None:
            g = g.view(SHAPES[leaf])
        if name.endswith("qkv.weight") or name.endswith("qkv.bias"):
            base = name[: -len("qkv.weight")] if name.endswith("weight") else name[: -len("qkv.bias")]
            kind = "weight" if name.endswith("weight") else "bias"
            q, k, v = g[:768], g[768:1536], g[1536:]
            grads[base + f"self_attn.q_proj.{kind}"] = q * 0.125   # the kernel's q rows carry head_dim^-0.5
            grads[base + f"self_attn.k_proj.{kind}"] = k
            grads[base + f"self_attn.v_proj.{kind}"] = v
        elif name.endswith("pos_conv.0.folded_weight"):
            # [g][n][tap][c] -> (out = g * 48 + n, in = c, tap); then through weight_norm(dim=2): w = g * v / ||v||
            dw = g.view(16, 48, 128, 48).permute(0, 1, 3, 2).reshape(768, 48, 128)
            v = _device_copy(engine, state_dict, "ssl_model.encoder.pos_conv.0.weight_v").requires_grad_(True)
            gg = _device_copy(engine, state_dict, "ssl_model.encoder.pos_conv.0.weight_g").requires_grad_(True)
            w = gg * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
            w.backward(dw)
            grads["ssl_model.encoder.pos_conv.0.weight_v"] = v.grad
            grads["ssl_model.encoder.pos_conv.0.weight_g"] = gg.grad
        else:
            grads[name] = g
    return loss, grads


class TripletTrainer:
    """``Training.train``'s inner step (``train_triplet.py:112-133``) with the optimiser of ``:92-107``."""

    def __init__(self, nomad, lr: float = 1e-4, margin: float = 0.2, lr_pretrained: float = 1e-5):
        self.nomad = nomad
        self.margin = margin
        dev = nomad.engine.device
        self.master = OrderedDict((k, v.detach().to(dev, torch.float32).clone()) for k, v in self._state_dict(nomad).items())
        head = ["embedding_layer.1.weight", "embedding_layer.1.bias"]
        train = [k for k in self.master if "feature_extractor" not in k and not k.endswith("mask_emb")]
        for k in train:
            self.master[k].requires_grad_(True)
        self.optim = torch.optim.Adam([{"params": [self.master[k] for k in train if k not in head], "lr": lr_pretrained},
                                       {"params": [self.master[k] for k in head]}], lr=lr, fused=True)

    @staticmethod
    def _state_dict(nomad):
        sd = getattr(nomad, "state_dict_ref", None)
        if sd is None:
            raise ValueError("TripletTrainer needs the Nomad to be built with keep_state_dict=True")
        return sd

    def step(self, anchor, positive, negative) -> float:
        loss, grads = triplet_loss_and_grads(self.nomad.engine, self.master, anchor, positive, negative, self.margin)
        self.optim.zero_grad(set_to_none=True)
        for k, g in grads.items():
            self.master[k].grad = g.reshape(self.master[k].shape).contiguous()
        self.optim.step()
        # rebuild the kernel-ready weights (16-bit copies, q|k|v fusion, transposes, LayerNorm folds, weight-norm fold) from
        # the updated master tensors, on the device
        self.nomad.engine.refresh_weights(self.master)
        return float(loss.item())
