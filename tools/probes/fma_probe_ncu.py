"""Run the FMA operand-pattern probes once each (for an ncu capture):
    ncu --set full -k regex:probe python tools/probes/fma_probe_ncu.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from pdenet_b200 import ops

for v in (3, 5, 6):
    print(v, ops.fma_peak(torch.float64, 256, v) / 1e12)
