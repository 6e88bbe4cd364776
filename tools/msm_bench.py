#!/usr/bin/env python3
"""Standalone MSM timing (device-resident scalars): python tools/msm_bench.py --log-n 20 [--reps 3]
Prints one JSON line with per-phase ms.  Env TP_MSM_CHUNK / TP_MSM_C / TP_MSM_LEVELS tune the kernels."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from typlonk_b200 import synthetic
from typlonk_b200.ffi import Context, fr_rand_stream
from typlonk_b200.kzg import Srs

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = Context(0, st.cuda_stream)
n = 1 << a.log_n
srs = Srs.from_secret(ctx, synthetic.tau(), n - 3)
base = torch.frombuffer(bytearray(fr_rand_stream(3, min(n, 1 << 16))), dtype=torch.uint8).to(dev)
sc = base.repeat(max(1, n >> 16)).view(n, 32).clone()
if n > (1 << 16):
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    mix = ((idx >> 16) * 2654435761) & 0xFFFFFFFF
    for b in range(4):
        sc[:, 8 + b] ^= ((mix >> (8 * b)) & 0xFF).to(torch.uint8)
torch.cuda.synchronize()
first = ctx.commit_dev(srs.handle, sc.data_ptr(), n)
ctx.prof_reset(); ctx.prof_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    out = ctx.commit_dev(srs.handle, sc.data_ptr(), n)
e1.record(); torch.cuda.synchronize()
assert out == first
prof = ctx.prof_get()
print(json.dumps({"log_n": a.log_n, "ms": e0.elapsed_time(e1) / a.reps,
                  "env": {k: os.environ.get(k) for k in ("TP_MSM_CHUNK", "TP_MSM_C", "TP_MSM_LEVELS")},
                  "phases": {k: round(v[0] / a.reps, 3) for k, v in prof.items() if k.startswith("msm")},
                  "digest": out[:8].hex()}))
