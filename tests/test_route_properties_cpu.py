"""Size-independent properties of the routing rule, checked on the CPU oracle with hypothesis-generated logits
(SURVEY.md 8a notes 1-8).  The GPU router is bit-compared with this oracle in tests/test_gpu_parity.py, so the
properties carry over."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import route_oracle_c as R

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}


def _logits(seed, T, scale, quant, dt):
    g = torch.Generator().manual_seed(seed)
    lg = torch.randn(T, 11, generator=g) * scale
    if quant:
        lg = torch.round(lg * quant) / quant        # exact ties and near ties
    return lg.to(dt)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), T=st.integers(1, 96), scale=st.sampled_from([0.05, 0.3, 0.9, 3.0, 20.0]),
       quant=st.sampled_from([0, 4, 64]), dname=st.sampled_from(["fp32", "bf16"]))
def test_routing_invariants(seed, T, scale, quant, dname):
    dt = DT[dname]
    lg = _logits(seed, T, scale, quant, dt)
    top_k, mask, gw, aux = R.route(lg)
    m, k = mask.numpy(), top_k.numpy()
    l9 = lg.float().numpy()[:, :9]
    assert ((k >= 1) & (k <= 10)).all()                                  # raw count (10: never reached the threshold)
    sel = m[:, :9].sum(1)
    assert (sel == np.where(k <= 9, k, 0)).all()                         # exactly k_t dynamic experts (note 1)
    assert (m[:, 9:] == 1).all()                                         # shared experts always on
    for t in range(T):                                                   # top-k by value, ties -> lowest index (note 2)
        order = sorted(range(9), key=lambda j: (-l9[t, j], j))
        assert set(np.nonzero(m[t, :9])[0]) == set(order[: sel[t]])
    g = gw.float().numpy()
    assert (g >= 0).all() and (g[:, :9][m[:, :9] == 0] == 0).all()       # weights only on selected experts
    tol = 3e-2 if dname == "bf16" else 1e-5
    assert np.allclose(g.sum(1), 1.0, atol=tol)                          # dynamic share + shared share = 1 (note 4)
    assert np.isfinite(aux.item()) and aux.item() >= 0


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), T=st.integers(2, 64), dname=st.sampled_from(["fp32", "bf16"]))
def test_routing_is_per_token_and_mask_only_zeroes_routed_columns(seed, T, dname):
    dt = DT[dname]
    lg = _logits(seed, T, 0.9, 0, dt)
    top_k, mask, gw, _ = R.route(lg)
    perm = torch.randperm(T, generator=torch.Generator().manual_seed(seed))
    top_k2, mask2, gw2, _ = R.route(lg[perm])                            # tokens are independent
    assert torch.equal(top_k2, top_k[perm]) and torch.equal(mask2, mask[perm]) and torch.equal(gw2, gw[perm])
    am = (torch.rand(T, generator=torch.Generator().manual_seed(seed + 1)) > 0.5)
    top_k3, mask3, gw3, _ = R.route(lg, am)                              # padding mask (note 8)
    assert torch.equal(top_k3, top_k)
    pad = ~am
    assert (mask3[pad, :9] == 0).all() and (mask3[pad, 9:] == 1).all() and (gw3[pad, :9] == 0).all()
    assert torch.equal(mask3[am], mask[am]) and torch.equal(gw3[am], gw[am])


@pytest.mark.parametrize("k", [1, 2, 5, 9])
def test_fixed_topk_selects_exactly_k(k):
    lg = _logits(7, 200, 0.9, 0, torch.float32)
    top_k, mask, gw, _ = R.route(lg, top_p=0.0, fixed_top_k=k)
    assert top_k.dtype == torch.int32 and (top_k == k).all()
    assert (mask[:, :9].sum(1) == k).all()
