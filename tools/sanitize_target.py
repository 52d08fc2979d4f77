"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): one large-path forward
(T = 300), decode-sized forwards (T = 2, 40), the glue block and a fixed top-k layer, checked against nothing --
the sanitizer is the checker.  python tools/sanitize_target.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE, PostAttentionMoE  # noqa: E402
dev = torch.device("cuda:0"); dt = torch.bfloat16
cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
           mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
           shared_intermediate_size=1376, router_jitter_noise=0.01, rms_norm_eps=1e-6)
gen = torch.Generator(device=dev).manual_seed(0)


def build(c):
    with torch.device("meta"):
        m = DCMoE(c)
    m = m.to(dt).to_empty(device=dev).eval()
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    return m


m = build(cfg)
for T in (300, 2, 40, 16384 if len(sys.argv) > 1 else 1000):
    x = torch.randn(1, T, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
    out = m(x, None, None)
    torch.cuda.synchronize()
    assert torch.isfinite(out[0].float()).all()
blk = PostAttentionMoE(cfg, mlp=m).to(dev).eval()
blk.post_attention_layernorm.weight.data = torch.ones(2048, device=dev, dtype=dt)
out = blk(torch.randn(1, 33, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt), None, None)
mk = build(dict(cfg, mlp_dynamic_top_p=0, mlp_dynamic_top_k=2))
out = mk(torch.randn(1, 50, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt), None, None)
m.ffn_impl = 2
out = m(torch.randn(1, 300, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt), None, None)
torch.cuda.synchronize()
print("sanitize target ok")
