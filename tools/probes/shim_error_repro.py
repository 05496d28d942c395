"""Localise a crash on the error path of the TORCH_LIBRARY shim (each case in its own process)."""
import subprocess
import sys

CASES = {
    "A_before_guard": "T.m2k(torch.zeros(4,4,dtype=torch.float64,device=dev), [], None)",
    "B_after_guard": "T.m2k(torch.zeros(1,3,3,dtype=torch.float64,device=dev), [1.0,2.0], None)",
    "C_after_call": "T.m2k(torch.zeros(1,4,4,dtype=torch.float64,device=dev), [], None)",
    "D_python_req": "P.m2k(torch.zeros(1,4,4,dtype=torch.float64))",
}
PRE = ("import sys, faulthandler; faulthandler.enable(); sys.path.insert(0,'/root/repo'); import torch; import pdenet_b200 as P; "
       "from pdenet_b200 import _lib; dev=torch.device('cuda:0'); T=_lib.torch_ops(); torch.zeros(1,device=dev)\n")
for name, stmt in CASES.items():
    code = PRE + "try:\n    %s\n    print('no error')\nexcept Exception as e:\n    print('caught', type(e).__name__, str(e)[:70])\n" % stmt
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print(name, "rc", r.returncode, r.stdout.strip()[:120], "|", r.stderr.strip().splitlines()[0][:80] if r.stderr.strip() else "")
