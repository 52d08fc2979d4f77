"""front_small phase counters, cold (L2 flushed) vs warm (called again at once).  python tools/front_dbg.py [T]"""
import os, sys
os.environ["DCMOE_ROUTER_DEBUG"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import ops  # noqa: E402
dev = torch.device("cuda:0"); dt = torch.bfloat16
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dims = ops.LayerDims()
g = torch.Generator(device=dev).manual_seed(0)
wg = (torch.randn(11, 2048, generator=g, device=dev) * 0.02).to(dt)
x = torch.randn(T, 2048, generator=g, device=dev).to(dt)
ws = ops.Workspace(dims, dt, T, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for rep in range(2):
    flush.fill_(rep); torch.cuda.synchronize()
    sys.stderr.write("cold: "); sys.stderr.flush()
    ops.front_small(x, wg, ws)
    sys.stderr.write("warm: "); sys.stderr.flush()
    ops.front_small(x, wg, ws)
torch.cuda.synchronize()
