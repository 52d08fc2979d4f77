"""GPU parity of the branches the V2 training recipe turns on (UniMoEV2-Preview/script/training.sh:46-59):
token_drop / drop_policy "probs" (reference utils/UniMoE_Audio_core.py:302-329, capacity :170-175), aux_balance_weight
(:380-385) and the training-mode forward with the fp32 gate and the input jitter (:240-249).  Fixtures come from the
UNMODIFIED reference block (tools/make_golden_drop.py); everything runs through the C ABI.

Bars: integer outputs (dynamic_top_k, expert_mask, the kept set) bit-exact -- with the tie rule "lower token index" where
the reference leaves ties to torch.topk; global_weight bit-exact given identical logits; aux 1e-5 (fp32) / 2e-3 (bf16);
layer output rtol 1e-2 (bf16).
"""
import os

import numpy as np
import pytest
import torch

from oracle import dcmoe_oracle as O
from oracle import route_oracle_c as R

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


_MODULES = {}


def _module(dt, dev, **cfg):
    from unimoe_audio_b200 import DCMoE
    key = (dt, tuple(sorted(cfg.items())))
    if key not in _MODULES:
        W = O.make_weights(seed=0, dtype=dt)
        with torch.device("meta"):
            m = DCMoE(dict(O.DEFAULT_CONFIG, **cfg))
        m = m.to(dt).to_empty(device=dev)
        m.load_state_dict({k: v.to(dev) for k, v in W.items()})
        _MODULES[key] = (m.eval(), W)
    return _MODULES[key]


def _zeros(T, dt, dev):
    return torch.zeros(1, T, 2048, dtype=dt, device=dev)


# ------------------------------------------------------------------ token drop
@pytest.mark.parametrize("dname", ["fp32", "bf16"])
@pytest.mark.parametrize("cf,tag", [(1, ""), (1, "_masked"), (2, ""), (2, "_masked")])
def test_token_drop_matches_reference_golden(dname, cf, tag, dev):
    g = np.load(os.path.join(GOLD, f"drop_{dname}.npz"))
    dt = DT[dname]
    logits = torch.from_numpy(g["logits"]).to(dt)
    T = logits.shape[0]
    am = torch.from_numpy(g["attention_mask"]) if tag else None
    key = f"cf{cf}{tag}"
    m, _ = _module(dt, dev, token_drop=True, drop_policy="probs", capacity_factor=float(cf), min_capacity=8)
    out = m(_zeros(T, dt, dev), None if am is None else am.to(dev), None, router_logits=logits.to(dev))
    torch.cuda.synchronize()
    ref_mask = torch.from_numpy(g[f"{key}_expert_mask"])
    mask = out[3].cpu()
    assert out[3].dtype == torch.int32 and out[2].dtype == torch.int64
    assert np.array_equal(out[2].cpu().numpy(), g[f"{key}_dynamic_top_k"])
    np.testing.assert_allclose(out[5].item(), float(g[f"{key}_aux_loss"]), rtol=1e-5 if dname == "fp32" else 2e-3)
    if bool(g[f"{key}_untied"]):        # fp32: unique kept set -> everything bit-equal to the reference
        assert np.array_equal(mask.numpy(), ref_mask.numpy())
        assert np.array_equal(out[4].float().cpu().numpy(), g[f"{key}_global_weight"])
    else:                               # bf16: tied boundaries -> same counts and the same multiset of kept logit values
        for e in range(9):
            assert int(mask[:, e].sum()) == int(ref_mask[:, e].sum())
            assert np.array_equal(np.sort(logits[mask[:, e] != 0, e].float().numpy()),
                                  np.sort(logits[ref_mask[:, e] != 0, e].float().numpy()))
    # and against the oracle (same tie rule): bit-equal in both dtypes
    cfg = dict(token_drop=True, drop_policy="probs", capacity_factor=float(cf), min_capacity=8)
    k2, m2, gw2, aux2 = O.route(logits, None if am is None else am.reshape(-1), cfg)
    assert torch.equal(mask, m2) and torch.equal(out[4].cpu(), gw2) and torch.equal(out[2].cpu(), k2)
    # the plan the FFN ran on is the one of the post-drop mask
    assert np.array_equal(m.last_workspace.counts.cpu().numpy(), m2[:, :8].sum(0).numpy())


@pytest.mark.parametrize("dname,T,cf", [("bf16", 16384, 1.0), ("fp32", 4099, 0.5), ("bf16", 777, 3.0), ("bf16", 40, 1.0),
                                        ("bf16", 65536, 2.0)])
def test_token_drop_select_bit_exact_vs_oracle_at_scale(dname, T, cf, dev):
    """The radix select on its own: keep mask == the oracle's stable-sort restatement, for quantised logits (many exact
    ties at every boundary) and continuous ones."""
    from unimoe_audio_b200 import ops
    dt = DT[dname]
    gen = torch.Generator().manual_seed(T)
    for quant in (False, True):
        lg = torch.randn(T, 11, generator=gen) * 0.9
        if quant:
            lg = torch.round(lg * 4) / 4
        lg = lg.to(dt)
        _k, mask, _gw, _a = R.route(lg)
        cap = O.expert_capacity(T, 9, cf, 8)
        ws = ops.Workspace(ops.LayerDims(), dt, T, dev)
        assert ops.expert_capacity(ops.LayerDims(), T, cf, 8) == cap
        keep = ops.drop_select(lg.to(dev), mask.to(dev), cap, ws)
        torch.cuda.synchronize()
        ref = O.drop_keep_mask(lg, mask, 9, cap)
        assert torch.equal(keep.cpu(), ref)
        for e in range(9):
            assert int(keep[:, e].sum()) == min(cap, int(mask[:, e].sum()))


def test_token_drop_full_layer_matches_oracle(dev):
    dt = torch.bfloat16
    cfg = dict(token_drop=True, drop_policy="probs", capacity_factor=1.0, min_capacity=8)
    m, W = _module(dt, dev, **cfg)
    x = torch.randn(2, 300, 2048, generator=torch.Generator().manual_seed(5)).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    ref = O.forward(x, W, None, cfg, logits=out[1].cpu())
    assert torch.equal(out[2].cpu(), ref.dynamic_top_k) and torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    assert int(ref.expert_mask[:, :9].sum()) < int(R.route(out[1].cpu())[1][:, :9].sum())       # something was dropped
    a, b = out[0].float().cpu().reshape(-1, 2048), ref.final_hidden_states.float().reshape(-1, 2048)
    assert ((a - b).abs() <= 1e-2 * b.abs() + 1e-2 * b.abs().max()).all()
    assert ((a - b).norm() / b.norm()).item() < 6e-3
    np.testing.assert_allclose(out[5].item(), ref.aux_loss.item(), rtol=2e-3)


def test_token_drop_with_zero_capacity_drops_every_routed_expert(dev):
    """capacity_factor = 0, min_capacity = 0 -> torch.topk(k = 0): every routed column is cleared, the layer is the two
    shared experts weighted by their 2-way softmax (core.py:306-316, :188-192)."""
    dt = torch.bfloat16
    cfg = dict(token_drop=True, drop_policy="probs", capacity_factor=0.0, min_capacity=0)
    m, W = _module(dt, dev, **cfg)
    x = torch.randn(1, 100, 2048, generator=torch.Generator().manual_seed(8)).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    assert int(out[3][:, :9].sum()) == 0 and bool((out[3][:, 9:] == 1).all())
    assert bool((out[4][:, :9] == 0).all())
    ref = O.forward(x, W, None, cfg, logits=out[1].cpu())
    assert torch.equal(out[3].cpu(), ref.expert_mask) and torch.equal(out[4].cpu(), ref.global_weight)
    a, b = out[0].float().cpu().reshape(-1, 2048), ref.final_hidden_states.float().reshape(-1, 2048)
    assert ((a - b).abs() <= 1e-2 * b.abs() + 1e-2 * b.abs().max()).all()


def test_token_drop_prints_the_reference_message_when_asked(dev, capsys):
    """drop_token_num_print (core.py:316-319): "drop N tokens from total M tokens", N and M counted over the 9 dynamic columns."""
    dt = torch.bfloat16
    cfg = dict(token_drop=True, drop_policy="probs", capacity_factor=1.0, min_capacity=8, drop_token_num_print=True)
    m, W = _module(dt, dev, **cfg)
    x = torch.randn(1, 512, 2048, generator=torch.Generator().manual_seed(9)).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    pre = R.route(out[1].cpu())[1][:, :9].sum().item()
    post = out[3][:, :9].sum().item()
    assert f"drop {pre - post} tokens from total {pre} tokens" in capsys.readouterr().out


def test_token_drop_policy_errors(dev):
    from unimoe_audio_b200 import DCMoE
    with pytest.raises(NotImplementedError):
        DCMoE(dict(O.DEFAULT_CONFIG, token_drop=True, drop_policy="position"))
    with torch.device("meta"):
        m = DCMoE(dict(O.DEFAULT_CONFIG, token_drop=True, drop_policy="nonsense"))
    m = m.to(torch.bfloat16).to_empty(device=dev).eval()
    with pytest.raises(ValueError, match="Invalid drop_policy"):       # core.py:325
        m(torch.zeros(1, 4, 2048, dtype=torch.bfloat16, device=dev), None, None)


# ------------------------------------------------------------------ aux_balance_weight
@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_aux_balance_weight_matches_reference_golden(dname, dev):
    g = np.load(os.path.join(GOLD, f"auxw_{dname}.npz"))
    dt = DT[dname]
    logits = torch.from_numpy(g["logits"]).to(dt)
    T = logits.shape[0]
    am = torch.from_numpy(g["attention_mask"])
    m, _ = _module(dt, dev)
    tol = 1e-5 if dname == "fp32" else 2e-3
    out = m(_zeros(T, dt, dev), am.to(dev), torch.from_numpy(g["w_int64"]).to(dev), router_logits=logits.to(dev))
    np.testing.assert_allclose(out[5].item(), float(g["int_aux_loss"]), rtol=tol)
    assert out[5].dtype == torch.float32 and out[5].dim() == 0
    assert np.array_equal(out[3].cpu().numpy(), g["int_expert_mask"])
    assert np.array_equal(out[4].float().cpu().numpy(), g["int_global_weight"])
    out = m(_zeros(T, dt, dev), None, torch.from_numpy(g["w_fp32"]).to(dev), router_logits=logits.to(dev))
    np.testing.assert_allclose(out[5].item(), float(g["float_aux_loss"]), rtol=tol)
    # same call without weights: the plain loss (the weighted branch must not leak into it)
    plain = m(_zeros(T, dt, dev), None, None, router_logits=logits.to(dev))
    assert abs(plain[5].item() - out[5].item()) > 1e-3
    # ... and with all-ones weights the two branches agree
    ones = m(_zeros(T, dt, dev), None, torch.ones(1, T, dtype=torch.int64, device=dev), router_logits=logits.to(dev))
    np.testing.assert_allclose(ones[5].item(), plain[5].item(), rtol=tol)


def test_aux_balance_weight_at_scale_vs_oracle(dev):
    dt = torch.bfloat16
    T = 16384
    gen = torch.Generator().manual_seed(9)
    logits = (torch.randn(T, 11, generator=gen) * 0.9).to(dt)
    w = torch.ones(8, 2048, dtype=torch.int64)
    w[torch.rand(8, 2048, generator=gen) > 0.5] = 10
    m, _ = _module(dt, dev)
    out = m(torch.zeros(8, 2048, 2048, dtype=dt, device=dev), None, w.to(dev), router_logits=logits.to(dev))
    ref = O.route(logits, None, None, w)
    np.testing.assert_allclose(out[5].item(), ref[3].item(), rtol=2e-3)


# ------------------------------------------------------------------ training-mode forward
def test_fp32_gate_training_forward_matches_reference_golden(dev):
    g = np.load(os.path.join(GOLD, "fp32gate_bf16.npz"))
    dt = torch.bfloat16
    m, W = _module(dt, dev, fp32_gate=True, input_jitter_noise=0.0)
    x = torch.randn(1, 256, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(dt)
    m.train()
    try:
        out = m(x.to(dev), None, None)
        torch.cuda.synchronize()
        assert out[1].dtype == torch.float32 and out[4].dtype == dt and out[0].dtype == dt
        # fp32 gate: logits agree with the reference's CPU fp32 projection to summation-order noise
        np.testing.assert_allclose(out[1].cpu().numpy(), g["full_router_logits"], rtol=0, atol=3e-6)
        # routing is a function of the logits: bit-exact against the oracle on the GPU's own logits ...
        ref = O.forward(x, W, fp32_gate=True, logits=out[1].cpu())
        assert torch.equal(out[2].cpu(), ref.dynamic_top_k) and torch.equal(out[3].cpu(), ref.expert_mask)
        assert torch.equal(out[4].cpu(), ref.global_weight)
        # ... and on the reference's logits it reproduces the reference's decisions
        out2 = m(x.to(dev), None, None, router_logits=torch.from_numpy(g["full_router_logits"]).to(dev))
        assert np.array_equal(out2[2].cpu().numpy(), g["dynamic_top_k"])
        assert np.array_equal(out2[3].cpu().numpy(), g["expert_mask"])
        assert np.array_equal(out2[4].float().cpu().numpy(), g["global_weight"])
        np.testing.assert_allclose(out2[5].item(), float(g["aux_loss"]), rtol=1e-5)
        final = out2[0].float().cpu().reshape(256, 2048)
        scale = float(np.abs(g["final_rows"]).max())
        np.testing.assert_allclose(final[::2].numpy(), g["final_rows"], rtol=1e-2, atol=1e-2 * scale)
    finally:
        m.eval()


@pytest.mark.parametrize("fp32_gate", [True, False])
def test_training_forward_input_jitter_consumes_torch_rng_like_the_reference(fp32_gate, dev):
    """core.py:243-244: `hidden_states *= torch.empty_like(hidden_states).uniform_(1 - e, 1 + e)`.  The product draws the
    noise with the same torch call, so with the same CUDA seed it equals 'noise applied by hand, jitter off'."""
    dt = torch.bfloat16
    mj, W = _module(dt, dev, fp32_gate=fp32_gate, input_jitter_noise=0.01)
    m0, _ = _module(dt, dev, fp32_gate=fp32_gate, input_jitter_noise=0.0)
    x = torch.randn(1, 128, 2048, generator=torch.Generator().manual_seed(3)).to(dt).to(dev)
    mj.train(), m0.train()
    try:
        torch.manual_seed(1234)
        xin = x.clone()
        a = mj(xin, None, None)
        torch.manual_seed(1234)
        if fp32_gate:       # the jitter only reaches the gate (a float copy, core.py:241-244); the input is untouched
            assert torch.equal(xin, x)
            xg = x.float()
            xg *= torch.empty_like(xg).uniform_(0.99, 1.01)
            logits = torch.nn.functional.linear(xg.reshape(-1, 2048), mj.gate.weight.detach().float())
            np.testing.assert_allclose(a[1].cpu().numpy(), logits.cpu().numpy(), rtol=0, atol=3e-6)
            b = m0(x, None, None, router_logits=a[1])
        else:               # in place on the caller's tensor, experts see the jittered rows
            xj = x.clone()
            xj *= torch.empty_like(xj).uniform_(0.99, 1.01)
            assert torch.equal(xin, xj) and not torch.equal(xin, x)
            b = m0(xj, None, None)
        for u, v in zip(a, b):
            assert torch.equal(u, v)
    finally:
        mj.eval(), m0.eval()


def test_training_forward_with_differentiable_router_is_rejected(dev):
    m, _ = _module(torch.bfloat16, dev, ignore_differentiable_router=False)
    m.train()
    try:
        with pytest.raises(NotImplementedError):
            m(torch.zeros(1, 4, 2048, dtype=torch.bfloat16, device=dev), None, None)
    finally:
        m.eval()
