"""Golden vectors for the branches of the block that utils/config.json leaves off but the V2 training recipe turns on
(UniMoEV2-Preview/script/training.sh:46-59), generated with the UNMODIFIED reference block:

  drop_<dtype>.npz        token_drop=True, drop_policy="probs" (utils/UniMoE_Audio_core.py:302-329, capacity :170-175):
                          router pinned on given logits, capacity_factor 1.0 and 2.0, with and without a padding mask
  drop_position_nan.npz   drop_policy="position" (core.py:321-323): the reference's result -- NaN weights for every token
                          past the capacity (the cumsum also runs over the shared experts' all-ones columns) -- which is
                          why the product rejects that policy
  auxw_<dtype>.npz        aux_balance_weight (core.py:380-385): int64 weights as the data collator builds them
                          (UniMoEV2-Preview/training/DataLoaders/qwen2vl_datasets.py:191-194) and float32 weights
  fp32gate_bf16.npz       training-mode forward with fp32_gate (core.py:240-249), input_jitter_noise = 0: full 6-tuple of
                          a bf16 block on 1 x 256 tokens
    python tools/make_golden_drop.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tools.make_golden import _np  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


@torch.no_grad()
def run_router(block, logits, attention_mask=None, aux_w=None):
    """The reference forward with the gate replaced by a constant; zero hidden states (the FFNs do not matter)."""
    T = logits.shape[0]

    class _Const(torch.nn.Module):
        def forward(self, _x):
            return logits

    gate = block.gate
    block.gate = _Const()
    try:
        x = torch.zeros(1, T, block.hidden_dim, dtype=logits.dtype)
        return block(x, attention_mask, aux_w)
    finally:
        block.gate = gate


def boundary_untied(logits, pre_mask, post_mask, n_dyn):
    """True when, in every dynamic column, the smallest kept logit is strictly larger than the largest dropped one."""
    for e in range(n_dyn):
        kept = logits[post_mask[:, e] != 0, e].float()
        dropped = logits[(pre_mask[:, e] != 0) & (post_mask[:, e] == 0), e].float()
        if kept.numel() and dropped.numel() and not (kept.min() > dropped.max()):
            return False
    return True


def main():
    g = lambda s: torch.Generator().manual_seed(s)  # noqa: E731
    T = 1024
    base = dict(ref_loader.reference_text_config())
    lg32 = torch.randn(T, 11, generator=g(71)) * 0.9
    am = torch.rand(1, T, generator=g(72)) > 0.25
    for dname, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        lg = lg32.to(dt)
        plain = ref_loader.build_reference_block(dict(base, token_drop=False), dtype=dt, seed=0)
        pre = run_router(plain, lg, None)[3]
        pre_m = run_router(plain, lg, am)[3]
        payload = {"logits": _np(lg), "attention_mask": _np(am), "pre_mask": _np(pre), "pre_mask_masked": _np(pre_m)}
        for cf in (1.0, 2.0):
            cfg = dict(base, token_drop=True, drop_policy="probs", capacity_factor=cf, min_capacity=8)
            block = ref_loader.build_reference_block(cfg, dtype=dt, seed=0)
            for tag, mask, pm in (("", None, pre), ("_masked", am, pre_m)):
                out = run_router(block, lg, mask)
                key = f"cf{int(cf)}{tag}"
                payload.update({f"{key}_dynamic_top_k": _np(out[2]), f"{key}_expert_mask": _np(out[3]),
                                f"{key}_global_weight": _np(out[4]), f"{key}_aux_loss": _np(out[5]),
                                f"{key}_untied": np.bool_(boundary_untied(lg, pm, out[3], 9))})
                print(dname, key, "dropped", int(pm[:, :9].sum() - out[3][:, :9].sum()), "of", int(pm[:, :9].sum()),
                      "untied boundary:", bool(payload[f"{key}_untied"]))
        np.savez_compressed(os.path.join(OUT, f"drop_{dname}.npz"), **payload)
        print("wrote", f"drop_{dname}.npz")

        # ---- aux_balance_weight ----
        wi = torch.ones(1, T, dtype=torch.int64)
        wi[torch.rand(1, T, generator=g(73)) > 0.6] = 10          # training.sh:59 --aux_balance_weight 10
        wi = wi * am.to(torch.int64)                               # UniMoEV2.py:1125: attention_mask * aux_balance_weight
        wf = torch.rand(1, T, generator=g(74)) + 0.25
        payload = {"logits": _np(lg), "attention_mask": _np(am), "w_int64": wi.numpy(), "w_fp32": wf.numpy()}
        for tag, w, mask in (("int", wi, am), ("float", wf, None)):
            out = run_router(plain, lg, mask, w)
            payload.update({f"{tag}_aux_loss": _np(out[5]), f"{tag}_expert_mask": _np(out[3]), f"{tag}_global_weight": _np(out[4])})
            print(dname, "aux_balance_weight", tag, float(out[5]), "(unweighted", float(run_router(plain, lg, mask)[5]), ")")
        np.savez_compressed(os.path.join(OUT, f"auxw_{dname}.npz"), **payload)
        print("wrote", f"auxw_{dname}.npz")

    # ---- position policy: what the reference returns ----
    cfg = dict(base, token_drop=True, drop_policy="position", capacity_factor=1.0, min_capacity=8)
    block = ref_loader.build_reference_block(cfg, dtype=torch.float32, seed=0)
    out = run_router(block, lg32, None)
    nan_rows = torch.isnan(out[4]).any(-1)
    np.savez_compressed(os.path.join(OUT, "drop_position_nan.npz"), logits=_np(lg32), expert_mask=_np(out[3]),
                        global_weight_is_nan=_np(nan_rows), capacity=np.int64(114))
    print("position policy: rows with NaN global weights:", int(nan_rows.sum()), "of", T, "first NaN row", int(torch.nonzero(nan_rows)[0]))

    # ---- training-mode forward with the fp32 gate ----
    cfg = dict(base, fp32_gate=True, input_jitter_noise=0.0)
    block = ref_loader.build_reference_block(cfg, dtype=torch.bfloat16, seed=0)
    block.train()
    x = torch.randn(1, 256, 2048, generator=g(4343)).to(torch.bfloat16)
    with torch.no_grad():
        out = block(x, None, None)
    assert out[1].dtype == torch.float32 and out[4].dtype == torch.bfloat16
    final = out[0].float().reshape(256, 2048)
    np.savez_compressed(os.path.join(OUT, "fp32gate_bf16.npz"), weight_seed=np.int64(0), x_seed=np.int64(4343),
                        final_rows=_np(final[::2]), full_router_logits=_np(out[1]), dynamic_top_k=_np(out[2]),
                        expert_mask=_np(out[3]), global_weight=_np(out[4]), aux_loss=_np(out[5]))
    print("wrote fp32gate_bf16.npz")


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available; fixtures can only be regenerated in the build container")
    main()
