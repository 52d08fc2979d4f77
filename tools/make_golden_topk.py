"""Golden vectors for the fixed top-k branch of the router (reference utils/UniMoE_Audio_core.py:254-257,
``mlp_dynamic_top_p == 0`` -> every token selects ``mlp_dynamic_top_k`` dynamic experts), generated with the
UNMODIFIED reference block.  Writes tests/golden/routek_{fp32,bf16}_k{2,3}.npz and layerk_bf16_k2.npz.
    python tools/make_golden_topk.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tools.make_golden import _np, run_reference_router  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    g = lambda s: torch.Generator().manual_seed(s)  # noqa: E731
    T = 512
    logits = {"iid": torch.randn(T, 11, generator=g(41)) * 0.9,
              "ties": torch.round(torch.randn(T, 11, generator=g(42)) * 8) / 8}
    mask = torch.rand(1, T, generator=g(43)) > 0.25
    for dname, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        for k in (2, 3):
            cfg = dict(ref_loader.reference_text_config())
            cfg.update(mlp_dynamic_top_p=0, mlp_dynamic_top_k=k)
            block = ref_loader.build_reference_block(cfg, dtype=dt, seed=0)
            payload = {}
            for cname, lg in logits.items():
                lg = lg.to(dt)
                out = run_reference_router(block, lg, mask if cname == "ties" else None)
                payload.update({f"{cname}_logits": _np(lg), f"{cname}_dynamic_top_k": _np(out[2]),
                                f"{cname}_expert_mask": _np(out[3]), f"{cname}_global_weight": _np(out[4]),
                                f"{cname}_aux_loss": _np(out[5])})
                assert out[2].dtype == torch.int32
            payload["ties_attention_mask"] = _np(mask)
            np.savez_compressed(os.path.join(OUT, f"routek_{dname}_k{k}.npz"), **payload)
            print("wrote", f"routek_{dname}_k{k}.npz")
    # one full layer, bf16, k = 2
    cfg = dict(ref_loader.reference_text_config())
    cfg.update(mlp_dynamic_top_p=0, mlp_dynamic_top_k=2)
    block = ref_loader.build_reference_block(cfg, dtype=torch.bfloat16, seed=0)
    x = torch.randn(1, 256, 2048, generator=g(4242)).to(torch.bfloat16)
    with torch.no_grad():
        out = block(x, None, None)
    final = out[0].float().reshape(256, 2048)
    np.savez_compressed(os.path.join(OUT, "layerk_bf16_k2.npz"), weight_seed=np.int64(0), x_seed=np.int64(4242),
                        final_rows=_np(final[::4]), full_router_logits=_np(out[1]), dynamic_top_k=_np(out[2]),
                        expert_mask=_np(out[3]), global_weight=_np(out[4]), aux_loss=_np(out[5]))
    print("wrote layerk_bf16_k2.npz")


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available; fixtures can only be regenerated in the build container")
    main()
