"""Decode-sized GEMM-1 / GEMM-2 device time (DCMOE_FFN_STREAM=0: the 128x256-tile kernel instead of the weight-streaming one), with the
weights of three layers rotated so that nothing is served from L2.
    python tools/bench_decode_gemm.py [T ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE, ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01)
    layers = []
    gen = torch.Generator(device=dev).manual_seed(0)
    for _ in range(3):
        with torch.device("meta"):
            m = DCMoE(cfg)
        m = m.to(dt).to_empty(device=dev).eval()
        with torch.no_grad():
            for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
                p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
        layers.append(m)
    Ts = [int(a) for a in sys.argv[1:]] or [2, 8, 32]
    reps = 12
    for T in Ts:
        x = torch.randn(T, 1, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
        for m in layers:
            m(x, None, None)
        ws = layers[0].last_workspace
        n_groups = int((ws.counts[:8] > 0).sum().item()) + 1 if hasattr(ws, "counts") else -1
        res = {}
        for phase in (1, 2):
            def run():
                for i in range(reps):
                    m = layers[i % 3]
                    ops.grouped_ffn(x.reshape(T, 2048), m._w13, m._w2, ws, 0, phase=phase)
            run()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            best = 1e9
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g.replay()
                e.record()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e) / reps * 1e3)
            res[phase] = best
        hit_mb = n_groups * 3 * 2048 * 2752 * 2 / 1e6
        print(f"stream={os.environ.get('DCMOE_FFN_STREAM', '1')} T={T:3d} "
              f"groups={n_groups}  gemm1 {res[1]:6.1f} us  gemm2 {res[2]:6.1f} us  sum {res[1] + res[2]:6.1f} us "
              f"(hit weights {hit_mb:.0f} MB -> {hit_mb / 6.5297:.1f} us at 6.53 TB/s)")


if __name__ == "__main__":
    main()
