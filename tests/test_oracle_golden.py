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


# ------------------------------------------------------------------ the real reference block -> DCMoE (build container only)
@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).reference_available(),
                    reason="the reference tree (/root/reference) only exists in the build container")
def test_real_reference_block_loads_into_dcmoe_strict():
    """the drop-in claims of INTEGRATION.md, checked against the UNMODIFIED reference class:
    identical state-dict keys and shapes (load_state_dict strict=True both ways), ``DCMoE.from_reference`` reads the
    block's configuration back, and the reference model's ``_init_weights`` isinstance hook (model.py:275-278) works on
    the replacement once the symbol is rebound."""
    from oracle import ref_loader
    from unimoe_audio_b200 import DCMoE, UniMoEAudioSparseMoeBlock

    block = ref_loader.build_reference_block(dtype=torch.bfloat16, seed=4)
    cfg = ref_loader.reference_text_config()
    ours = DCMoE(cfg).to(torch.bfloat16)
    sd = block.state_dict()
    assert list(sd.keys()) == list(ours.state_dict().keys())
    assert all(sd[k].shape == v.shape and sd[k].dtype == v.dtype for k, v in ours.state_dict().items())
    res = ours.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    block.load_state_dict(ours.state_dict(), strict=True)           # and back
    assert all(torch.equal(sd[k], v) for k, v in ours.state_dict().items())
    # from_reference: every constructor scalar the reference stores is read back (core.py:204-234)
    twin = DCMoE.from_reference(block)
    for attr in ("hidden_dim", "mlp_dynamic_expert_num", "mlp_dynamic_real_expert_num", "mlp_dynamic_null_expert_num",
                 "mlp_dynamic_top_p", "mlp_dynamic_top_k", "mlp_fixed_expert_num", "num_experts", "router_jitter_noise",
                 "token_drop", "drop_policy", "capacity_factor", "min_capacity", "fp32_gate", "avg_hidden_states_last"):
        assert getattr(twin, attr) == getattr(block, attr), attr
    assert twin.dims.dynamic_intermediate_size == cfg["dynamic_intermediate_size"]
    assert twin.dims.shared_intermediate_size == cfg["shared_intermediate_size"]
    assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), sd.values()))
    # the rebind of INTEGRATION.md section 1: model.py:275 does isinstance(module, UniMoEAudioSparseMoeBlock)
    assert UniMoEAudioSparseMoeBlock is DCMoE and isinstance(twin, UniMoEAudioSparseMoeBlock)
    twin.gate.weight.data.normal_(mean=0.0, std=0.02)                # what _init_weights does to the block (model.py:276-278)
    # sub-module paths the reference touches (core.py:356, :510)
    assert twin.dynamic_real_moe.deepspeed_moe.ep_group is None
    assert len(twin.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts) == cfg["mlp_dynamic_expert_num"]


# ------------------------------------------------------------------ token drop / aux_balance_weight / fp32 gate
# (fixtures of tools/make_golden_drop.py: the branches the V2 training recipe turns on, UniMoEV2-Preview/script/training.sh:46-59)
DROP_CASES = [(cf, tag) for cf in (1, 2) for tag in ("", "_masked")]


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
@pytest.mark.parametrize("cf,tag", DROP_CASES)
def test_token_drop_oracle_matches_reference_golden(dname, cf, tag, golden_dir):
    """drop_policy "probs" (core.py:302-329).  fp32: the boundary logits are untied, so the kept set is unique and the
    whole result must be bit-equal.  bf16: boundaries are tied (torch.topk's pick among equals is implementation
    defined), so the selection is checked as a property -- per expert the kept COUNT and the multiset of kept logit
    VALUES equal the reference's -- and the mask + renormalise step is pinned bit-exactly on the reference's own kept set."""
    g = np.load(os.path.join(golden_dir, f"drop_{dname}.npz"))
    dt = DT[dname]
    logits = torch.from_numpy(g["logits"]).to(dt)
    am = torch.from_numpy(g["attention_mask"]) if tag else None
    key = f"cf{cf}{tag}"
    cfg = dict(token_drop=True, drop_policy="probs", capacity_factor=float(cf), min_capacity=8)
    top_k, mask, gw, aux = O.route(logits, None if am is None else am.reshape(-1), cfg)
    ref_mask = torch.from_numpy(g[f"{key}_expert_mask"])
    pre = torch.from_numpy(g["pre_mask_masked" if tag else "pre_mask"])
    cap = O.expert_capacity(logits.shape[0], 9, float(cf), 8)
    assert cap == {1: 114, 2: 228}[cf]
    assert np.array_equal(top_k.numpy(), g[f"{key}_dynamic_top_k"])
    np.testing.assert_allclose(aux.item(), float(g[f"{key}_aux_loss"]), rtol=1e-5)      # computed BEFORE the drop
    for e in range(9):
        assert int(mask[:, e].sum()) == int(ref_mask[:, e].sum()) == min(cap, int(pre[:, e].sum()))
        a = np.sort(logits[mask[:, e] != 0, e].float().numpy())
        b = np.sort(logits[ref_mask[:, e] != 0, e].float().numpy())
        assert np.array_equal(a, b)
    assert (mask[:, 9:] == 1).all()
    if bool(g[f"{key}_untied"]):
        assert dname == "fp32"
        assert np.array_equal(mask.numpy(), ref_mask.numpy())
        assert np.array_equal(gw.float().numpy(), g[f"{key}_global_weight"])
    # the renormalise step on the reference's kept set (exact in both dtypes)
    keep = ref_mask.to(torch.uint8).clone()
    keep[:, 9:] = 1
    if am is not None:          # a padded token's columns are cleared by the padding mask, not by the capacity
        keep[~am.reshape(-1), :9] = 1
    _tk, m2, gw2, _ = R.route(logits, None if am is None else am.reshape(-1), keep=keep)
    assert np.array_equal(m2.numpy(), ref_mask.numpy())
    assert np.array_equal(gw2.float().numpy(), g[f"{key}_global_weight"])


def test_token_drop_position_policy_is_nan_in_the_reference(golden_dir):
    """Why drop_policy="position" (core.py:321-323) is rejected: the reference clears the shared experts' columns too,
    so most tokens past the capacity end with NaN weights.  Nothing to be compatible with."""
    g = np.load(os.path.join(golden_dir, "drop_position_nan.npz"))
    nan = g["global_weight_is_nan"]
    assert nan.sum() > 0.5 * nan.shape[0] and not nan[: int(g["capacity"])].any()
    assert (g["expert_mask"][int(g["capacity"]):, 9:] == 0).all()       # shared experts dropped for every later token
    with pytest.raises(ValueError):
        O.route(torch.from_numpy(g["logits"]), None, dict(token_drop=True, drop_policy="position"))


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_aux_balance_weight_oracle_matches_reference_golden(dname, golden_dir):
    g = np.load(os.path.join(golden_dir, f"auxw_{dname}.npz"))
    dt = DT[dname]
    logits = torch.from_numpy(g["logits"]).to(dt)
    am = torch.from_numpy(g["attention_mask"]).reshape(-1)
    tol = 1e-5 if dname == "fp32" else 2e-3
    _, mask, gw, aux = O.route(logits, am, None, torch.from_numpy(g["w_int64"]))
    np.testing.assert_allclose(aux.item(), float(g["int_aux_loss"]), rtol=tol)
    assert np.array_equal(mask.numpy(), g["int_expert_mask"]) and np.array_equal(gw.float().numpy(), g["int_global_weight"])
    _, mask, gw, aux = O.route(logits, None, None, torch.from_numpy(g["w_fp32"]))
    np.testing.assert_allclose(aux.item(), float(g["float_aux_loss"]), rtol=tol)
    assert np.array_equal(mask.numpy(), g["float_expert_mask"])


def test_fp32_gate_training_forward_oracle_matches_reference_golden(golden_dir):
    """Training-mode forward with fp32_gate (core.py:240-249): fp32 logits and routing arithmetic on a bf16 layer."""
    g = np.load(os.path.join(golden_dir, "fp32gate_bf16.npz"))
    W = O.make_weights(seed=int(g["weight_seed"]), dtype=torch.bfloat16)
    x = torch.randn(1, 256, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(torch.bfloat16)
    out = O.forward(x, W, fp32_gate=True)
    assert out.full_router_logits.dtype == torch.float32 and out.global_weight.dtype == torch.bfloat16
    np.testing.assert_allclose(out.full_router_logits.numpy(), g["full_router_logits"], rtol=0, atol=2e-6)
    # routing pinned on the reference's own fp32 logits
    out = O.forward(x, W, fp32_gate=True, logits=torch.from_numpy(g["full_router_logits"]))
    assert np.array_equal(out.dynamic_top_k.numpy(), g["dynamic_top_k"])
    assert np.array_equal(out.expert_mask.numpy(), g["expert_mask"])
    assert np.array_equal(out.global_weight.float().numpy(), g["global_weight"])
    np.testing.assert_allclose(out.aux_loss.item(), float(g["aux_loss"]), rtol=1e-5)
    final = out.final_hidden_states.float().reshape(256, 2048)
    scale = float(np.abs(g["final_rows"]).max())
    np.testing.assert_allclose(final[::2].numpy(), g["final_rows"], rtol=1e-2, atol=1e-2 * scale)


# ------------------------------------------------------------------ the router's float-pair exponential (oracle/exp_fast.h)
def test_exp_fast_statement_is_correctly_rounded_on_a_strided_sweep(tmp_path):
    """tools/verify_exp_fast.c over every 1024th float of (-80, 0) (the full sweep, 1.1e9 inputs, is recorded in
    profiles/r02_exp_fast_exhaustive.txt): no accepted value may differ from the long-double exponential, and the previous
    definition (float)exp((double)x) must agree with it everywhere."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "verify_exp_fast")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, os.path.join(root, "tools", "verify_exp_fast.c"), "-lm"], check=True)
    res = subprocess.run([exe, "1024"], capture_output=True, text=True, check=True)
    m = re.search(r"inputs (\d+)\s+accepted-but-wrong (\d+)\s+fallbacks (\d+).*correctly rounded: (\d+)", res.stdout)
    assert m, res.stdout
    n, wrong, fb, dbl = (int(v) for v in m.groups())
    assert n > 1_000_000 and wrong == 0 and dbl == 0 and fb < n // 10_000


def test_exp_fast_agrees_with_the_double_precision_definition():
    import ctypes
    lib = R.lib()
    rng = np.random.default_rng(5)
    a = (rng.standard_normal(20000) * 1.5).astype(np.float32)
    b = (rng.standard_normal(20000) * 1.5).astype(np.float32)
    xs = -np.abs(torch.from_numpy(a).to(torch.bfloat16).float().numpy() - torch.from_numpy(b).to(torch.bfloat16).float().numpy())
    fb = ctypes.c_int(0)
    n_fb = 0
    for x in xs.tolist() + [0.0, -1e-30, -79.9, -80.0, -100.0]:
        v = lib.dcmoe_oracle_exp_fast(x, ctypes.byref(fb))
        n_fb += fb.value
        if not fb.value:
            assert np.float32(v).tobytes() == np.float32(lib.dcmoe_oracle_exp_cr(x)).tobytes(), x
    assert n_fb <= 5
