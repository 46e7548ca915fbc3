"""Timing of the triplet fine-tuning step at the reference's training shape (src/config/train_triplet.yaml: train_bs 8,
clips trimmed to 10 s): three forwards + TripletMarginLoss + all parameter gradients in one C-ABI call."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nomad_b200.engine import Engine
from nomad_b200.triplet import triplet_loss_and_grads
from nomad_b200.weights import flops_embed, random_state_dict

sd = random_state_dict(1234)
eng = Engine(sd, 0)
out = []
for B, secs in ((8, 10), (8, 4), (32, 2)):
    N = 16000 * secs
    g = torch.Generator().manual_seed(0)
    A, P, Nn = (0.1 * torch.randn(B, N, generator=g) for _ in range(3))
    A, P, Nn = A.cuda(), P.cuda(), Nn.cuda()
    for _ in range(2):
        triplet_loss_and_grads(eng, sd, A, P, Nn)
    torch.cuda.synchronize()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib_loss, _ = triplet_loss_and_grads(eng, sd, A, P, Nn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    # forward of 3B utterances + dgrad + wgrad of everything after the conv encoder (conv: forward only)
    t = eng.num_frames(N)
    conv = 2.0 * (5120 * ((N - 10) // 5 + 1)) + 0  # conv0 only term of F(N); the rest of the conv stack below
    fwd = flops_embed(N)
    enc = 786432.0 * t + 9437184.0 * t + 12.0 * (4718592.0 * t + 9437184.0 * t + 3072.0 * t * t) + 393216.0
    flops = 3 * B * (fwd + 2.0 * enc)
    out.append({"B": B, "seconds": secs, "frames_per_utt": t, "ms_per_step": ms, "launches_per_step": (eng.launch_count() - n0) // 5,
                "tflops_algorithmic": flops / (ms / 1e3) / 1e12, "triplets_per_s": B / (ms / 1e3)})
print(json.dumps({"config": "triplet step: 3B forwards + TripletMarginLoss + all parameter gradients (conv encoder frozen), "
                            "flops = 3B * (F(N) + 2 * F_encoder(N))", "results": out}))
