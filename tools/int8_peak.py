"""Measure the dense int8 tcgen05 rate of this GPU (rn_int8_peak) and write profiles/r02_int8_peak.json:
burst = best of 10 short launches, sustained = mean over ~3 s of back-to-back launches."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from renormalizer_b200 import _lib
lib = _lib.get()
st = _lib.stream_ptr()
t = ctypes.c_double(0)
lib.rn_int8_peak(st, 2000, ctypes.byref(t))           # warm up
burst = []
for _ in range(10):
    lib.rn_int8_peak(st, 20000, ctypes.byref(t))
    burst.append(t.value)
sus, t0 = [], time.time()
while time.time() - t0 < 3.0:
    lib.rn_int8_peak(st, 200000, ctypes.byref(t))
    sus.append(t.value)
out = {"int8_tops_burst": max(burst), "int8_tops_sustained": sum(sus[len(sus) // 2:]) / len(sus[len(sus) // 2:]),
       "samples_sustained": len(sus), "gpu": torch.cuda.get_device_name(0),
       "how": "rn_int8_peak: 148 CTAs x back-to-back tcgen05.mma.cta_group::1.kind::i8 128x128x32 on resident "
              "128B-swizzled shared-memory tiles, 4 TMEM accumulators round-robin, CUDA events; burst = best of 10 "
              "launches of 20000 x 4 MMAs, sustained = mean of the second half of 3 s of back-to-back launches"}
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
