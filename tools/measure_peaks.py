"""Measures the FP64 / TF32 roofline denominators on the box (MEASURED_PEAKS.json only has HBM and bf16):
cuBLAS DGEMM and TF32 GEMM 8192^3 through torch (best of 10), plus tools/fp64_peak (DMMA / DFMA register loops).
Writes gpurun_out/peaks_fp64.json; a copy is committed as profiles/peaks_fp64.json."""
import json
import os
import subprocess
import sys

import torch

out = {}
here = os.path.dirname(os.path.abspath(__file__))
try:
    out["micro"] = json.loads(subprocess.check_output([os.path.join(here, "fp64_peak")]).decode())
except Exception as e:  # noqa
    out["micro_error"] = str(e)
n = 8192
for name, dtype, tf32 in (("dgemm_tflops", torch.float64, False), ("tf32_tflops", torch.float32, True), ("fp32_tflops", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    best = 1e9
    for i in range(8):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); c = a @ b; e.record(); torch.cuda.synchronize()
        if i: best = min(best, s.elapsed_time(e))
    out[name] = 2 * n ** 3 / best / 1e9
    del a, b, c
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/peaks_fp64.json", "w"), indent=1)
print(json.dumps(out))
