"""Golden vectors for the decoder-layer glue around the MoE block (reference utils/UniMoE_Audio_model.py:239-242):

    residual = h;  h = post_attention_layernorm(h);  h, ... = mlp(h, mask, None);  h = residual + h

generated with the UNMODIFIED reference block (oracle/ref_loader.py) and the RMSNorm class the reference imports
(transformers Qwen2RMSNorm, model.py:54 / :207; transformers is the reference's pinned dependency and is installed in
the build container).  Writes tests/golden/glue_{fp32,bf16}.npz.
    python tools/make_golden_glue.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EPS = 1e-6          # utils/config.json text_config.rms_norm_eps
T = 128


def _np(t):
    return (t.float() if t.dtype == torch.bfloat16 else t).detach().cpu().numpy()


def main():
    from transformers.models.qwen2.modeling_qwen2 import Qwen2RMSNorm
    for dname, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        block = ref_loader.build_reference_block(dtype=dt, seed=0)
        norm = Qwen2RMSNorm(2048, eps=EPS)
        w = (1.0 + 0.1 * torch.randn(2048, generator=torch.Generator().manual_seed(77))).to(dt)
        with torch.no_grad():
            norm.weight.data = w.clone()
        norm = norm.to(dt)
        x = (torch.randn(1, T, 2048, generator=torch.Generator().manual_seed(4321)) * 1.7).to(dt)
        with torch.no_grad():
            n = norm(x)
            out = block(n, None, None)
            final = x + out[0]
        np.savez_compressed(os.path.join(OUT, f"glue_{dname}.npz"), weight_seed=np.int64(0), x_seed=np.int64(4321),
                            x_scale=np.float64(1.7), norm_weight_seed=np.int64(77), eps=np.float64(EPS),
                            normed_rows=_np(n.reshape(T, 2048)[::4]), final_rows=_np(final.reshape(T, 2048)[::4]),
                            full_router_logits=_np(out[1]), expert_mask=_np(out[3]))
        print("wrote", f"glue_{dname}.npz")


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available; fixtures can only be regenerated in the build container")
    main()
