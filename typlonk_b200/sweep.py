"""BASELINE.json configs[3]: standalone KZG G1 MSM sweep (2^16 .. 2^26 points, plus the skewed and all-equal scalar
cases of SURVEY.md 8(d)) and Fr NTT / inverse NTT / coset NTT sweep (2^16 .. 2^24), device resident, timed with CUDA
events on the ctx stream.  `run` returns one dict per point and, given `out`, writes them as JSON lines."""
import json

from . import field as F, synthetic
from .ffi import fr_rand_stream
from .kzg import Srs


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


FR_MUL_WIDE_PRODUCTS = 112   # wide multiply-adds of one Fr Montgomery product (tools/gen_mont.py; profiles r1 Q)


def _time_msm(ctx, torch, srs, sc, n, reps):
    ctx.commit_dev(srs.handle, sc.data_ptr(), n)
    ctx.prof_reset(); ctx.prof_enable(True)
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(reps):
        ctx.commit_dev(srs.handle, sc.data_ptr(), n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = ctx.prof_get(); ctx.prof_enable(False)
    return ms, {k: round(v[0] / reps, 4) for k, v in prof.items() if k.startswith("msm")}


def run(ctx, torch, out=None, msm_logs=(16, 18, 20, 22, 24, 26), ntt_logs=(16, 18, 20, 22, 24), reps=3, skew_log=20):
    dev = torch.device("cuda", ctx.device)
    imad, imad_wide = ctx.measure_imad_peak()
    rows = []

    def emit(row):
        rows.append(row)
        if out is not None:
            out.write(json.dumps(row) + "\n")
    for lg in msm_logs:
        n = 1 << lg
        # one SRS per size: its fixed-base tables (window bits, levels) are planned for that length
        srs = Srs.from_secret(ctx, synthetic.tau(), n - 3)
        # uniform scalars: Fr::rand stream (seed 3) for the first 2^16, then a device-side mix
        base = torch.frombuffer(bytearray(fr_rand_stream(synthetic.SEED_MSM, 1 << 16)), dtype=torch.uint8).to(dev)
        sc = base.repeat(n >> 16).view(n, 32).clone()
        if n > (1 << 16):
            # make repeats distinct while staying < r: xor a counter into limb bytes 8..11
            idx = torch.arange(n, device=dev, dtype=torch.int64)
            mix = ((idx >> 16) * 2654435761) & 0xFFFFFFFF
            for b in range(4):
                sc[:, 8 + b] ^= ((mix >> (8 * b)) & 0xFF).to(torch.uint8)
        torch.cuda.synchronize()
        ms, phases = _time_msm(ctx, torch, srs, sc, n, reps)
        emit({"sweep": "msm", "scalars": "uniform", "log_n": lg, "ms": round(ms, 4), "mpts_per_s": round(n / ms / 1e3, 2),
              "hbm_gbs_algorithmic": round(128.0 * n / ms / 1e6, 2), "window_bits": ctx.get_stat("msm_window_bits"),
              "windows": ctx.get_stat("msm_windows"), "phases_ms": phases})
        if lg == skew_log:
            # scalars below 2^16 (every digit in the lowest window) and one repeated scalar (one bucket)
            # (canonical values; the Montgomery encoding the ABI carries is full-width, so it is made on the host)
            import random
            rnd = random.Random(16)
            small = torch.frombuffer(bytearray(F.fr_vec_to_bytes([rnd.randrange(1 << 16) for _ in range(n)])),
                                     dtype=torch.uint8).to(dev).view(n, 32)
            ms, phases = _time_msm(ctx, torch, srs, small, n, reps)
            emit({"sweep": "msm", "scalars": "below_2^16", "log_n": lg, "ms": round(ms, 4),
                  "mpts_per_s": round(n / ms / 1e3, 2), "phases_ms": phases})
            small[:] = small[0].clone()
            ms, phases = _time_msm(ctx, torch, srs, small, n, reps)
            emit({"sweep": "msm", "scalars": "all_equal", "log_n": lg, "ms": round(ms, 4),
                  "mpts_per_s": round(n / ms / 1e3, 2), "phases_ms": phases})
            del small
        del sc
        srs.handle.destroy()
        torch.cuda.empty_cache()
    for lg in ntt_logs:
        n = 1 << lg
        x = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
        x[:, 31] &= 0x3F
        torch.cuda.synchronize()
        for name, kw in (("ntt", {}), ("intt", {"inverse": True}), ("coset_ntt", {"coset_mont": F.fr_to_bytes(7)})):
            ctx.ntt_dev(x.data_ptr(), lg, **kw)
            e0, e1 = _events(torch)
            e0.record()
            for _ in range(reps):
                ctx.ntt_dev(x.data_ptr(), lg, **kw)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            butterflies = (n // 2) * lg
            emit({"sweep": name, "log_n": lg, "ms": round(ms, 4), "gbs_algorithmic": round(64.0 * n / ms / 1e6, 1),
                  "gbutterflies_per_s": round(butterflies / ms / 1e6, 2),
                  "wide_mul_frac": round(butterflies * FR_MUL_WIDE_PRODUCTS / (ms * 1e-3) / imad_wide, 4)})
        del x
    if out is not None:
        out.flush()
    return rows


if __name__ == "__main__":
    import argparse
    import sys

    import torch

    from .ffi import Context

    ap = argparse.ArgumentParser()
    ap.add_argument("--msm", default="16,18,20,22,24,26")
    ap.add_argument("--ntt", default="16,18,20,22,24")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    st = torch.cuda.Stream(device=torch.device("cuda", 0))
    torch.cuda.set_stream(st)  # torch events must sit on the stream the library launches on
    c = Context(0, st.cuda_stream)
    run(c, torch, sys.stdout, tuple(int(x) for x in a.msm.split(",") if x), tuple(int(x) for x in a.ntt.split(",") if x),
        a.reps)
