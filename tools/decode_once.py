"""A few eager decode-sized forwards (profiling target for ncu).  python tools/decode_once.py [T] [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE  # noqa: E402
dev = torch.device("cuda:0"); dt = torch.bfloat16
cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
           mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
           shared_intermediate_size=1376, router_jitter_noise=0.01)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
layers = []
gen = torch.Generator(device=dev).manual_seed(0)
for _ in range(2):
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    layers.append(m)
x = torch.randn(T, 1, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt)
for i in range(iters):
    layers[i % 2](x, None, None)
torch.cuda.synchronize()
print("ok")
