"""How far are the drop-in's Top-P decisions from the reference AS USERS RUN IT (bf16 on CUDA)?

The router reproduces the rounding points of torch's CPU kernels (the arithmetic the fixtures were generated with,
DESIGN.md section 3).  The reference ships on CUDA in bf16 (reference utils/UniMoE_Audio_mod.py:44, :81-83), where ATen's
softmax / cumsum kernels accumulate differently, and 1-2 % of tokens sit within one bf16 ulp of the `>= top_p` edge.
This script runs the reference's five ATen calls (core.py:162-166: softmax, sort, cumsum, >=, sum) with torch ON THE GPU
and, for comparison, on the CPU, on the same bf16 logits the router sees, and reports the fraction of tokens whose
dynamic_top_k differs from the router's.  No /root/reference needed.
    python tools/topk_cuda_vs_canonical.py > profiles/r02_topk_cuda_vs_canonical.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimoe_audio_b200 import ops  # noqa: E402


def reference_top_k(logits9: torch.Tensor, top_p: float) -> torch.Tensor:
    scores = torch.softmax(logits9, dim=-1)                      # core.py:162
    s, _ = torch.sort(scores, dim=-1, descending=True)           # :163
    c = s.cumsum(dim=-1)                                         # :164
    return (~(c >= top_p)).sum(dim=-1) + 1                       # :165-166


def main():
    dev = torch.device("cuda:0")
    dims = ops.LayerDims()
    T = 262144
    out = {"tokens": T, "top_p": dims.top_p, "dtype": "bf16", "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "cases": {}}
    gen = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(T, 2048, generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    wg = (torch.randn(11, 2048, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    cases = {"gate projection of N(0,1) activations, N(0,0.02^2) gate (logit sigma ~0.9)": torch.nn.functional.linear(x, wg),
             "iid N(0, 0.3^2) logits (flat router)": (torch.randn(T, 11, generator=gen, device=dev) * 0.3).to(torch.bfloat16),
             "iid N(0, 2^2) logits (peaky router)": (torch.randn(T, 11, generator=gen, device=dev) * 2.0).to(torch.bfloat16)}
    for name, lg in cases.items():
        lg = lg.contiguous()
        ws = ops.Workspace(dims, torch.bfloat16, T, dev)
        _, ours, mask, _gw = ops.router(None, None, ws, logits_in=lg)
        k_cuda = reference_top_k(lg[:, :9], dims.top_p)
        k_cpu = reference_top_k(lg[:, :9].cpu(), dims.top_p)
        torch.cuda.synchronize()
        d_cuda = (ours != k_cuda)
        d_cpu = (ours.cpu() != k_cpu)
        diff = (k_cuda - ours)[d_cuda]
        out["cases"][name] = {
            "router_vs_torch_cpu_fraction_differing": float(d_cpu.float().mean()),
            "router_vs_torch_cuda_fraction_differing": float(d_cuda.float().mean()),
            "torch_cuda_selects_one_more": int((diff == 1).sum()), "torch_cuda_selects_one_fewer": int((diff == -1).sum()),
            "differences_larger_than_one": int((diff.abs() > 1).sum()),
            "mean_top_k_router": float(ours.float().mean()), "mean_top_k_torch_cuda": float(k_cuda.float().mean()),
            "torch_cpu_vs_torch_cuda_fraction_differing": float((k_cpu != k_cuda.cpu()).float().mean())}
    out["reading"] = ("the router is bit-identical to torch's CPU kernels (column 1 = 0); torch's CUDA kernels round the softmax / "
                      "running sum at other points, so a token whose cumulative probability lands within one bf16 ulp of top_p can "
                      "get one expert more or fewer there -- the same disagreement torch has with itself between its two devices "
                      "(last column)")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
