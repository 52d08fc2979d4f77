"""Decoder-layer glue around the MoE block (reference model.py:239-242) at BASELINE.json configs[1] (8 x 2048 tokens,
bf16): PostAttentionMoE (dcmoe_rmsnorm + layer + residual fused in combine) against the same layer with the norm and
the residual add done by separate torch kernels; and at decode size (T = 2) under CUDA-graph replay.
    python tools/bench_glue.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE, PostAttentionMoE, ops  # noqa: E402


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01, rms_norm_eps=1e-6)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    gen = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    blk = PostAttentionMoE(cfg, mlp=m).to(dev).eval()
    blk.post_attention_layernorm.weight.data = (1 + 0.1 * torch.randn(2048, generator=gen, device=dev)).to(dt)
    w = blk.post_attention_layernorm.weight.detach()

    def torch_norm(h):
        hf = h.float()
        return w * (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + 1e-6)).to(dt)

    for B, S, n in ((8, 2048, 30), (2, 1, 200)):
        xs = [torch.randn(B, S, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt) for _ in range(6)]
        i = [0]

        def fused():
            i[0] += 1
            return blk(xs[i[0] % 6], None, None)

        def unfused():
            i[0] += 1
            h = xs[i[0] % 6]
            o = m(torch_norm(h), None, None)
            return h + o[0]

        def norm_only():
            i[0] += 1
            return ops.rmsnorm(xs[i[0] % 6], w, 1e-6, m.dims)

        def moe_only():
            i[0] += 1
            return m(xs[i[0] % 6], None, None)

        t_f, t_u, t_n, t_m = timed(fused, n), timed(unfused, n), timed(norm_only, n), timed(moe_only, n)
        T = B * S
        gbs = 2 * T * 2048 * 2 / (t_n * 1e-6) / 1e9
        print(f"T={T:6d}: PostAttentionMoE {t_f:8.1f} us | torch norm + layer + torch add {t_u:8.1f} us | layer alone {t_m:8.1f} us | "
              f"dcmoe_rmsnorm alone {t_n:6.1f} us ({gbs:.0f} GB/s)")


if __name__ == "__main__":
    main()
