"""Router (+ plan) alone at BASELINE.json configs[1] size and larger: time per call from CUDA events over back-to-back
calls on rotating inputs (at 16,384 tokens this loop is bound by the host side of the Python wrapper -- run it under
`ncu --metrics gpu__time_duration.sum` for the kernel's own duration), algorithmic bytes 4192 B/token
against the measured HBM peak.
    python tools/bench_router.py [T ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6529.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    Ts = [int(a) for a in sys.argv[1:]] or [16384, 65536, 262144]
    dt = torch.bfloat16
    for T in Ts:
        gen = torch.Generator(device=dev).manual_seed(1)
        n_rot = max(2, min(8, (1 << 30) // (T * 4096)))
        xs = [torch.randn(T, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt) for _ in range(n_rot)]
        wg = (torch.randn(11, 2048, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt)
        ws = ops.Workspace(ops.LayerDims(), dt, T, dev)
        res = {}
        for name, fn in (("router_only", lambda x: ops.router(x, wg, ws)),
                         ("router_then_plan", lambda x: (ops.router(x, wg, ws), ops.plan(ws)))):
            for i in range(5):
                fn(xs[i % n_rot])
            torch.cuda.synchronize()
            n = 50
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(n):
                fn(xs[i % n_rot])
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) / n * 1e3
            res[name] = {"us": round(us, 2), "gbs": round(T * 4192 / us / 1e3, 1), "frac_of_hbm_peak": round(T * 4192 / us / 1e3 / peak, 3)}
        print(json.dumps({"T": T, "hbm_peak_gbs": peak, **res}), flush=True)


if __name__ == "__main__":
    main()
