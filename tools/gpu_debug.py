"""Stage-by-stage diagnostics on a GPU box (not a test; prints error maps to help debug kernels).

    python tools/gpu_debug.py [T]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import dcmoe_oracle as O  # noqa: E402  (checker only)
from oracle import route_oracle_c as R  # noqa: E402
from unimoe_audio_b200 import DCMoE, ops  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    print("device", torch.cuda.get_device_name(0), "T", T, flush=True)
    W = O.make_weights(seed=0, dtype=dt)
    with torch.device("meta"):
        m = DCMoE(dict(O.DEFAULT_CONFIG))
    m = m.to(dt).to_empty(device=dev).eval()
    m.load_state_dict({k: v.to(dev) for k, v in W.items()})
    x = torch.randn(1, T, 2048, generator=torch.Generator().manual_seed(7)).to(dt).to(dev)

    # --- CUDA-core path first (no tcgen05) ---
    m.ffn_impl = 1
    out_cc = m(x, None, None)
    torch.cuda.synchronize()
    ws = m.last_workspace
    h_cc, y_cc = ws.h.clone(), ws.y.clone()
    k2, m2, gw2, aux2 = R.route(out_cc[1].cpu())
    print("router: top_k eq", torch.equal(out_cc[2].cpu(), k2), "mask eq", torch.equal(out_cc[3].cpu(), m2),
          "gw eq", torch.equal(out_cc[4].cpu(), gw2), "aux", out_cc[5].item(), aux2.item(), flush=True)
    Tref = min(T, 512)
    ref = O.forward(x[:, :Tref].cpu(), W, None, logits=out_cc[1][:Tref].cpu())
    a, b = out_cc[0][0, :Tref].float().cpu(), ref.final_hidden_states[0].float()
    print("cuda-core layer vs oracle: max err %.3e rel fro %.3e (max |ref| %.3f)" %
          ((a - b).abs().max().item(), ((a - b).norm() / b.norm()).item(), b.abs().max().item()), flush=True)
    n_mt = ws.n_mtiles.item()
    mt = ws.mtiles[:n_mt].cpu()
    print("counts", ws.counts.cpu().tolist(), "seg", ws.seg_base.cpu().tolist(), "n_mtiles", n_mt, flush=True)

    # --- tcgen05 path ---
    m.ffn_impl = 0
    t0 = time.time()
    out_tc = m(x, None, None)
    torch.cuda.synchronize()
    print("tcgen05 forward done in %.3f s" % (time.time() - t0), flush=True)
    h_tc, y_tc = ws.h, ws.y
    used = int(ws.seg_base.cpu()[-1].item())
    valid = torch.zeros(ws.row_capacity, dtype=torch.bool)
    for i in range(n_mt):
        r0, rows = mt[i, 1].item(), mt[i, 3].item()
        valid[r0:r0 + rows] = True
    valid = valid.to(dev)
    for name, a_, b_ in (("h", h_tc, h_cc), ("y", y_tc, y_cc)):
        a, b = a_[valid].float(), b_[valid].float()
        err = (a - b).abs()
        print(f"{name}: max err {err.max().item():.3e}  rel fro {(err.norm() / b.norm()).item():.3e}  "
              f"nan {torch.isnan(a).sum().item()}  max|ref| {b.abs().max().item():.3f}", flush=True)
        if (err.norm() / b.norm()).item() > 2e-2:
            # error map: per 128-row tile x 64-col block
            rows = torch.nonzero(valid).flatten()
            e2 = torch.zeros(ws.row_capacity, a_.shape[1], device=dev)
            e2[valid] = err
            ntile = used // 128
            em = e2[:ntile * 128].reshape(ntile, 128, -1, 64).amax(dim=(1, 3)).cpu()
            torch.set_printoptions(linewidth=250, precision=2, sci_mode=False)
            print("per (m-tile, 64-col block) max err:", em[:, :].shape)
            print(em[: min(ntile, 12), : 48])
            # row pattern inside the first bad tile
            bad = torch.nonzero(em.amax(1) > 1e-2).flatten()
            if bad.numel():
                t = bad[0].item()
                print("first bad tile", t, "per-row max err (128 rows):")
                print(e2[t * 128:(t + 1) * 128].amax(1).cpu().reshape(8, 16))
                print("per-col max err first 256 cols:")
                print(e2[t * 128:(t + 1) * 128, :256].amax(0).cpu().reshape(-1, 32))
    a, b = out_tc[0].float(), out_cc[0].float()
    print("tcgen05 layer vs cuda-core layer: max err %.3e rel fro %.3e" %
          ((a - b).abs().max().item(), ((a - b).norm() / b.norm()).item()), flush=True)

    # --- timing of the stages (CUDA events) ---
    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    wg = m.gate.weight.detach()
    xf = x.reshape(T, 2048)
    lg, tk, mk, gw = ops.router(xf, wg, ws)
    ops.plan(ws)
    print("router  %.4f ms" % timeit(lambda: ops.router(xf, wg, ws)))
    print("plan    %.4f ms" % timeit(lambda: ops.plan(ws)))
    print("permute %.4f ms" % timeit(lambda: ops.permute(xf, mk, gw, ws)))
    print("ffn tc  %.4f ms" % timeit(lambda: ops.grouped_ffn(xf, m._w13, m._w2, ws, 0)))
    outb = torch.empty_like(xf)
    print("combine %.4f ms" % timeit(lambda: ops.combine(ws, outb)))
    ms = timeit(lambda: m(x, None, None))
    A = ws.counts.sum().item()
    flops = 45056 * T + (T + A) * 33816576
    print("layer   %.4f ms  -> %.0f tokens/s, %.1f TFLOP/s (A=%d, r=%.2f)" % (ms, T / ms * 1e3, flops / ms / 1e9, A, A / T))


if __name__ == "__main__":
    main()
