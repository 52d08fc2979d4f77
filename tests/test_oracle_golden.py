"""Pin the oracle (oracle/route_oracle.c + oracle/dcmoe_oracle.py) to the fixtures produced by
the UNMODIFIED reference block (tools/make_golden.py).  CPU only.

Bars: dynamic_top_k / expert_mask / global_weight bit-exact (the C restatement reproduces the
reference's rounding points exactly in both dtypes); aux_loss rtol 1e-5; layer output within
1e-5 (fp32) / 1e-2 (bf16) of the reference's (observed: 1.2e-7 abs / bit-equal).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dcmoe_oracle as O
from oracle import route_oracle_c as R

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}
ROUTE_FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "route_*.npz")))


def test_fixtures_present():
    assert len(ROUTE_FILES) == 12


@pytest.mark.parametrize("path", ROUTE_FILES, ids=[os.path.basename(p)[:-4] for p in ROUTE_FILES])
def test_route_oracle_matches_reference_golden(path):
    g = np.load(path)
    dt = DT[os.path.basename(path).split("_")[1]]
    logits = torch.from_numpy(g["logits"]).to(dt)
    am = torch.from_numpy(g["attention_mask"]) if "attention_mask" in g.files else None
    top_k, mask, gw, aux = R.route(logits, am)
    assert top_k.dtype == torch.int64 and mask.dtype == torch.int32 and gw.dtype == dt
    assert np.array_equal(top_k.numpy(), g["dynamic_top_k"])
    assert np.array_equal(mask.numpy(), g["expert_mask"])
    assert np.array_equal(gw.float().numpy(), g["global_weight"])          # bit-exact, both dtypes
    np.testing.assert_allclose(aux.item(), float(g["aux_loss"]), rtol=1e-5)
    if am is not None:  # padded tokens route only to the shared experts (core.py:286-291)
        pad = ~am.reshape(-1)
        assert (mask[pad, :9] == 0).all() and (mask[pad, 9:] == 1).all()
        assert (gw[pad, :9] == 0).all()


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_layer_oracle_matches_reference_golden(dname, golden_dir):
    g = np.load(os.path.join(golden_dir, f"layer_{dname}_c1.npz"))
    dt = DT[dname]
    W = O.make_weights(seed=int(g["weight_seed"]), dtype=dt)
    x = torch.randn(1, 512, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(dt)
    out = O.forward(x, W)
    assert np.array_equal(out.full_router_logits.float().numpy(), g["full_router_logits"])
    assert np.array_equal(out.dynamic_top_k.numpy(), g["dynamic_top_k"])
    assert np.array_equal(out.expert_mask.numpy(), g["expert_mask"])
    assert np.array_equal(out.global_weight.float().numpy(), g["global_weight"])
    np.testing.assert_allclose(out.aux_loss.item(), float(g["aux_loss"]), rtol=1e-5)
    final = out.final_hidden_states.float().reshape(512, 2048)
    rtol = 1e-5 if dname == "fp32" else 1e-2
    scale = float(np.abs(g["final_rows"]).max())
    np.testing.assert_allclose(final[::4].numpy(), g["final_rows"], rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(final.double().sum().item(), float(g["final_sum"]), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(final.double().abs().sum().item(), float(g["final_abs_sum"]), rtol=1e-4)
    # output dtypes of the 6-tuple (SURVEY.md 8a note 6)
    t = out.as_tuple()
    assert [o.dtype for o in t] == [dt, dt, torch.int64, torch.int32, dt, torch.float32]
    assert t[0].shape == (1, 512, 2048) and t[5].dim() == 0


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_glue_oracle_matches_reference_golden(dname, golden_dir):
    """rmsnorm + MoE + residual (model.py:239-242) against the reference block + transformers' Qwen2RMSNorm."""
    g = np.load(os.path.join(golden_dir, f"glue_{dname}.npz"))
    dt = DT[dname]
    W = O.make_weights(seed=int(g["weight_seed"]), dtype=dt)
    x = (torch.randn(1, 128, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))) * float(g["x_scale"])).to(dt)
    nw = (1.0 + 0.1 * torch.randn(2048, generator=torch.Generator().manual_seed(int(g["norm_weight_seed"])))).to(dt)
    n = O.rmsnorm(x, nw, float(g["eps"]))
    assert n.dtype == dt
    assert np.array_equal(n.float().reshape(128, 2048)[::4].numpy(), g["normed_rows"])      # same torch ops: bit-exact
    final, o = O.glue_forward(x, nw, W, eps=float(g["eps"]))
    assert np.array_equal(o.full_router_logits.float().numpy(), g["full_router_logits"])
    assert np.array_equal(o.expert_mask.numpy(), g["expert_mask"])
    rtol = 1e-5 if dname == "fp32" else 1e-2
    scale = float(np.abs(g["final_rows"]).max())
    np.testing.assert_allclose(final.float().reshape(128, 2048)[::4].numpy(), g["final_rows"], rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
@pytest.mark.parametrize("k", [2, 3])
def test_fixed_topk_route_oracle_matches_reference_golden(dname, k, golden_dir):
    """mlp_dynamic_top_p == 0 (core.py:254-257): every token selects mlp_dynamic_top_k experts; top_k is int32."""
    g = np.load(os.path.join(golden_dir, f"routek_{dname}_k{k}.npz"))
    for case in ("iid", "ties"):
        lg = torch.from_numpy(g[f"{case}_logits"]).to(DT[dname])
        am = torch.from_numpy(g["ties_attention_mask"]) if case == "ties" else None
        top_k, mask, gw, aux = R.route(lg, am, top_p=0.0, fixed_top_k=k)
        assert top_k.dtype == torch.int32 and np.array_equal(top_k.numpy(), g[f"{case}_dynamic_top_k"])
        assert np.array_equal(mask.numpy(), g[f"{case}_expert_mask"])
        assert np.array_equal(gw.float().numpy(), g[f"{case}_global_weight"])
        np.testing.assert_allclose(aux.item(), float(g[f"{case}_aux_loss"]), rtol=1e-5)


def test_fixed_topk_layer_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "layerk_bf16_k2.npz"))
    dt = torch.bfloat16
    W = O.make_weights(seed=int(g["weight_seed"]), dtype=dt)
    x = torch.randn(1, 256, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(dt)
    out = O.forward(x, W, cfg=dict(mlp_dynamic_top_p=0, mlp_dynamic_top_k=2))
    assert np.array_equal(out.full_router_logits.float().numpy(), g["full_router_logits"])
    assert out.dynamic_top_k.dtype == torch.int32 and np.array_equal(out.dynamic_top_k.numpy(), g["dynamic_top_k"])
    assert np.array_equal(out.expert_mask.numpy(), g["expert_mask"])
    assert np.array_equal(out.global_weight.float().numpy(), g["global_weight"])
    final = out.final_hidden_states.float().reshape(256, 2048)
    scale = float(np.abs(g["final_rows"]).max())
    np.testing.assert_allclose(final[::4].numpy(), g["final_rows"], rtol=1e-2, atol=1e-2 * scale)


def test_route_edge_cases():
    # empty input, single token, all-equal logits (exact 9-way tie -> lowest indices win)
    top_k, mask, gw, aux = R.route(torch.zeros(0, 11))
    assert top_k.numel() == 0 and mask.shape == (0, 11)
    top_k, mask, gw, _ = R.route(torch.zeros(1, 11))
    # uniform p = 1/9: prefixes 1/9..: first prefix >= 0.7 is the 7th -> k = 7
    assert top_k.item() == 7
    assert mask[0].tolist() == [1] * 7 + [0, 0] + [1, 1]
    # one dominant logit -> k = 1, weight 1 on it before global scaling
    lg = torch.full((1, 11), -10.0)
    lg[0, 3] = 10.0
    top_k, mask, gw, _ = R.route(lg)
    assert top_k.item() == 1 and mask[0].tolist() == [0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 1]
    assert abs(gw[0].sum().item() - 1.0) < 1e-5
    # null expert (index 8) selectable, never dispatched: mask col 8 set, weight kept in gw
    lg = torch.full((1, 11), -10.0)
    lg[0, 8] = 10.0
    top_k, mask, gw, _ = R.route(lg)
    assert mask[0, 8] == 1 and mask[0, :8].sum() == 0
