"""Diagnostic: ONE rn_svd call on a graded 1432 x 512 block (profiling target; not part of the product).
    python tools/svd_one.py [m] [n] [cplx]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from renormalizer_b200 import _lib, ops
_lib.get()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1432
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cplx = len(sys.argv) > 3 and sys.argv[3] == "1"
rng = np.random.default_rng(0)
a = rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if cplx else 0)
u, s, vh = np.linalg.svd(a, full_matrices=False)
a = (u * np.exp(-0.05 * np.arange(len(s)))) @ vh
ad = torch.from_numpy(a).cuda()
ops.svd(ad)                      # warm-up (module load, allocator)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
u, s, vh = ops.svd(ad)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("sweeps", ops.svd.last_sweeps, "recon", float(((u * s) @ vh - ad).abs().max()))
