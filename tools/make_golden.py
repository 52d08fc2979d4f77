"""Generate tests/golden/*.npz by running the UNMODIFIED reference block.

Runs only in the build container (needs /root/reference; see oracle/ref_loader.py for how the
reference module is imported with zero edits).  The fixtures are committed; the GPU box never
reads /root/reference.

  python tools/make_golden.py            # rewrite every fixture

Fixtures
  route_<dtype>_<case>.npz   router pinned on a given logits tensor (the gate Linear of the
                             reference block is bypassed so both sides see identical logits):
                             logits, [attention_mask], dynamic_top_k, expert_mask, global_weight,
                             aux_loss   -- reference core.py:252-332, :361-389
  layer_<dtype>_c1.npz       BASELINE.json config 1 (1 x 512 tokens, utils/config.json dims,
                             N(0, 0.02^2) weights seed 0, x ~ N(0,1) seed 1235): the full 6-tuple
                             (final hidden states for every 4th token + whole-tensor checksums)
bf16 tensors are stored as float32 holding bf16-representable values.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16}


def _np(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        t = t.float()
    return t.detach().cpu().numpy()


def route_cases():
    g = lambda s: torch.Generator().manual_seed(s)  # noqa: E731
    T = 1024
    cases = {}
    cases["iid09"] = dict(logits=torch.randn(T, 11, generator=g(11)) * 0.9)
    cases["flat03"] = dict(logits=torch.randn(T, 11, generator=g(12)) * 0.3)
    cases["peaky20"] = dict(logits=torch.randn(T, 11, generator=g(13)) * 2.0)
    # exact ties and near-ties: logits quantised to 1/8 (stresses first-index tie-break + the
    # 2 % near-tie multipliers of core.py:105-119)
    cases["ties"] = dict(logits=torch.round(torch.randn(T, 11, generator=g(14)) * 8) / 8)
    # skewed router of BASELINE.json config 5: bias linspace(+2,-2) on the 9 dynamic logits
    skew = torch.randn(T, 11, generator=g(15)) * 0.9
    skew[:, :9] += torch.linspace(2.0, -2.0, 9)
    cases["skew"] = dict(logits=skew)
    # padding mask (core.py:286-288)
    cases["masked"] = dict(logits=torch.randn(T, 11, generator=g(16)) * 0.9,
                           attention_mask=(torch.rand(1, T, generator=g(17)) > 0.25))
    return cases


@torch.no_grad()
def run_reference_router(block, logits: torch.Tensor, attention_mask):
    """Run the reference forward with the gate replaced by a constant (identical logits on both
    sides); hidden states are zeros so the expert FFNs cost little and do not matter."""
    T = logits.shape[0]

    class _Const(torch.nn.Module):
        def forward(self, _x):
            return logits

    gate = block.gate
    block.gate = _Const()
    try:
        x = torch.zeros(1, T, block.hidden_dim, dtype=logits.dtype)
        out = block(x, attention_mask, None)
    finally:
        block.gate = gate
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for dname, dt in DTYPES.items():
        block = ref_loader.build_reference_block(dtype=dt, seed=0)
        for cname, case in route_cases().items():
            lg = case["logits"].to(dt)
            am = case.get("attention_mask")
            out = run_reference_router(block, lg, am)
            payload = dict(logits=_np(lg), dynamic_top_k=_np(out[2]), expert_mask=_np(out[3]),
                           global_weight=_np(out[4]), aux_loss=_np(out[5]))
            if am is not None:
                payload["attention_mask"] = _np(am)
            np.savez_compressed(os.path.join(OUT, f"route_{dname}_{cname}.npz"), **payload)
            print("wrote", f"route_{dname}_{cname}.npz", "mean k", out[2].float().mean().item())

        # ---- full layer, BASELINE.json config 1 ----
        x = torch.randn(1, 512, 2048, generator=torch.Generator().manual_seed(1235)).to(dt)
        out = block(x, None, None)
        final = out[0].float().reshape(512, 2048)
        np.savez_compressed(
            os.path.join(OUT, f"layer_{dname}_c1.npz"),
            weight_seed=np.int64(0), x_seed=np.int64(1235),
            final_rows=_np(final[::4]), final_sum=np.float64(final.double().sum().item()),
            final_abs_sum=np.float64(final.double().abs().sum().item()),
            full_router_logits=_np(out[1]), dynamic_top_k=_np(out[2]), expert_mask=_np(out[3]),
            global_weight=_np(out[4]), aux_loss=_np(out[5]))
        print("wrote", f"layer_{dname}_c1.npz")


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available; fixtures can only be regenerated in the build container")
    main()
