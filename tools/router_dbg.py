"""Per-role cycle counters of the TMA-fed router (DCMOE_ROUTER_DEBUG=1).  python tools/router_dbg.py [T]"""
import os, sys
os.environ["DCMOE_ROUTER_DEBUG"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import ops  # noqa: E402
dev = torch.device("cuda:0"); dt = torch.bfloat16
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dims = ops.LayerDims()
g = torch.Generator(device=dev).manual_seed(0)
wg = (torch.randn(11, 2048, generator=g, device=dev) * 0.02).to(dt)
ws = ops.Workspace(dims, dt, T, dev)
for i in range(4):
    x = torch.randn(T, 2048, generator=g, device=dev).to(dt)
    ops.router(x, wg, ws)
torch.cuda.synchronize()
