"""BASELINE config 1 through the library's own graphed training closure (losses.GraphedTrainingLoss): a few
replays for an ncu launch list (which nodes make up the captured chain), and the replay time without ncu.

    python tools/probes/cfg1_closure_launches.py            # prints ms per replay
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/probes/cfg1_closure_launches.py 3
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import pdenet_b200 as P
from bench import build_model, synth_fields
from pdenet_b200 import losses

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, H, W, steps = 8, 100, 100, 1
model, _ = build_model(P, dev, torch.float64, H, W, 0, True)
u0 = synth_fields(0, B, 2, H, W, dev, torch.float64)
closure = losses.GraphedTrainingLoss(model, u0, torch.stack([u0] * steps))
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("replays")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    loss, grad = closure()
e1.record()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(json.dumps({"config": "cfg1 through losses.GraphedTrainingLoss", "replays": reps,
                  "ms_per_replay": e0.elapsed_time(e1) / reps, "loss": float(loss),
                  "grad_norm": float(grad.norm())}))
