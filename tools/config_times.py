"""Time the small BASELINE configs (1-3) fwd+bwd on one GPU: python tools/config_times.py
(the same measurement bench.py reports as its `small_configs` sub-record)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pdenet_b200 as P
from bench import small_configs_record
from pdenet_b200 import ops

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peaks = {"f64": ops.fma_peak(torch.float64, 2048) / 1e12, "f32": ops.fma_peak(torch.float32, 2048) / 1e12}
    for r in small_configs_record(P, dev, peaks):
        print(json.dumps(r))
