"""Un-hooked forward loop at BASELINE.json configs[1] (8 x 2048 tokens, bf16): device time per step with no events
between the kernels (A/B for DCMOE_PDL and similar switches).  python tools/bench_plain.py [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import DCMoE  # noqa: E402
dev = torch.device("cuda:0"); dt = torch.bfloat16
cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
           mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
           shared_intermediate_size=1376, router_jitter_noise=0.01)
with torch.device("meta"):
    m = DCMoE(cfg)
m = m.to(dt).to_empty(device=dev).eval()
gen = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
        p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
xs = [torch.randn(8, 2048, 2048, generator=gen, device=dev, dtype=torch.float32).to(dt) for _ in range(4)]
out = None
for i in range(8):
    out = m(xs[i % 4], None, None)
torch.cuda.synchronize()
for rep in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        out = m(xs[i % 4], None, None)
    e.record()
    torch.cuda.synchronize()
    print(f"PDL={os.environ.get('DCMOE_PDL', '1')} rep {rep}: {s.elapsed_time(e) / steps * 1e3:8.1f} us/step")
