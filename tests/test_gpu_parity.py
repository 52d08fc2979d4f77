"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C ABI
(unimoe_audio_b200.ops -> libdcmoe_b200.so); the oracle (oracle/) is only the checker.

Bars (BASELINE.json north_star):
  * dynamic_top_k / expert_mask / per-expert counts / canonical permutation: BIT-EXACT given identical
    router logits -- and here global_weight too, because the kernel restates the oracle's arithmetic;
  * layer outputs: rtol 1e-2 in bf16, 1e-5 in fp32 (with an absolute floor of rtol * max|ref| -- the
    outputs are sums of ~10 signed terms, so elements near zero carry the error of their largest term);
  * aux_loss rtol 1e-5 in fp32 (2e-3 in bf16, where the reference rounds the mean probability to bf16).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dcmoe_oracle as O
from oracle import route_oracle_c as R

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROUTE_FILES = sorted(glob.glob(os.path.join(GOLD, "route_*.npz")))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _dims():
    from unimoe_audio_b200.ops import LayerDims
    return LayerDims()


def _route_gpu(logits, dev, attention_mask=None):
    from unimoe_audio_b200 import ops
    ws = ops.Workspace(_dims(), logits.dtype, logits.shape[0], dev, row_capacity=0)
    lg, top_k, mask, gw = ops.router(None, None, ws, logits_in=logits.to(dev).contiguous(), attention_mask=attention_mask)
    ops.plan(ws)
    torch.cuda.synchronize()
    return lg, top_k, mask, gw, ws


# ------------------------------------------------------------------ router
@pytest.mark.parametrize("path", ROUTE_FILES, ids=[os.path.basename(p)[:-4] for p in ROUTE_FILES])
def test_router_matches_reference_golden_bit_exact(path, dev):
    g = np.load(path)
    dt = DT[os.path.basename(path).split("_")[1]]
    logits = torch.from_numpy(g["logits"]).to(dt)
    am = torch.from_numpy(g["attention_mask"]) if "attention_mask" in g.files else None
    lg, top_k, mask, gw, ws = _route_gpu(logits, dev, am)
    assert top_k.dtype == torch.int64 and mask.dtype == torch.int32 and gw.dtype == dt and lg.dtype == dt
    assert np.array_equal(lg.float().cpu().numpy(), g["logits"])
    assert np.array_equal(top_k.cpu().numpy(), g["dynamic_top_k"])
    assert np.array_equal(mask.cpu().numpy(), g["expert_mask"])
    assert np.array_equal(gw.float().cpu().numpy(), g["global_weight"])
    np.testing.assert_allclose(ws.aux_loss.item(), float(g["aux_loss"]), rtol=1e-5 if dt == torch.float32 else 2e-3)
    assert np.array_equal(ws.counts.cpu().numpy(), g["expert_mask"][:, :8].sum(0))


@pytest.mark.parametrize("dname,T,scale", [("bf16", 16384, 0.9), ("fp32", 16384, 0.9), ("bf16", 8191, 0.3),
                                           ("fp32", 4099, 2.0), ("bf16", 65536, 1.3)])
def test_router_bit_exact_vs_oracle_at_scale(dname, T, scale, dev):
    dt = DT[dname]
    logits = (torch.randn(T, 11, generator=torch.Generator().manual_seed(T)) * scale).to(dt)
    lg, top_k, mask, gw, ws = _route_gpu(logits, dev)
    k2, m2, gw2, aux2 = R.route(logits)
    assert torch.equal(top_k.cpu(), k2)
    assert torch.equal(mask.cpu(), m2)
    assert torch.equal(gw.cpu(), gw2)
    np.testing.assert_allclose(ws.aux_loss.item(), aux2.item(), rtol=1e-5 if dt == torch.float32 else 2e-3)


@pytest.mark.parametrize("dname", ["bf16", "fp32"])
def test_gate_projection_matches_torch(dname, dev):
    from unimoe_audio_b200 import ops
    dt = DT[dname]
    T = 1000  # ragged: not a multiple of 16
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(T, 2048, generator=gen).to(dt).to(dev)
    wg = (torch.randn(11, 2048, generator=gen) * 0.02).to(dt).to(dev)
    ws = ops.Workspace(_dims(), dt, T, dev)
    lg, *_ = ops.router(x, wg, ws)
    ref = (x.double() @ wg.double().T)
    err = (lg.double() - ref).abs().max().item()
    assert err <= (2e-2 if dt == torch.bfloat16 else 2e-5), err      # |logit| ~ 1: bf16 ulp 2^-8 at 1..2
    # and torch's own kernel agrees to the same level
    ref_t = torch.nn.functional.linear(x, wg).double()
    assert (lg.double() - ref_t).abs().max().item() <= (2e-2 if dt == torch.bfloat16 else 2e-5)


# ------------------------------------------------------------------ plan + permute
@pytest.mark.parametrize("dname,T", [("bf16", 2048 + 37), ("fp32", 515), ("bf16", 1)])
def test_permutation_is_the_canonical_stable_one(dname, T, dev):
    from unimoe_audio_b200 import ops
    dt = DT[dname]
    gen = torch.Generator().manual_seed(11 + T)
    logits = (torch.randn(T, 11, generator=gen) * 0.9).to(dt)
    x = torch.randn(T, 2048, generator=gen).to(dt).to(dev)
    lg, top_k, mask, gw, ws = _route_gpu(logits, dev)
    ops.permute(x, mask, gw, ws)
    torch.cuda.synchronize()
    o_k, o_mask, o_gw, _ = R.route(logits)
    perm = O.canonical_permutation(o_mask, 8)
    counts = ws.counts.cpu()
    seg = ws.seg_base.cpu()
    assert counts.tolist() == [p.numel() for p in perm]
    t_pad = ws.t_pad
    assert seg[0].item() == t_pad
    row_token = ws.row_token.cpu()
    slot_of = ws.slot_of.cpu()
    xp = ws.x_packed.cpu()
    xs = x.cpu()
    for e in range(8):
        s, c = seg[e].item(), counts[e].item()
        assert s % 128 == 0
        assert torch.equal(row_token[s:s + c].long(), perm[e])                 # permutation indices, bit exact
        assert torch.equal(xp[s - t_pad:s - t_pad + c], xs[perm[e]])            # gathered rows, bit exact
        assert torch.equal(slot_of[perm[e], e].long(), torch.arange(s, s + c))  # inverse map
        scale = ws.row_scale[s:s + c].cpu()
        assert torch.equal(scale[:, 0], o_gw[perm[e], e].float()) and torch.equal(scale[:, 1], scale[:, 0])
    assert (slot_of[o_mask[:, :8] == 0] == -1).all()
    assert torch.equal(ws.row_scale[:T].cpu(), o_gw[:, 9:11].float())
    # tile table covers exactly the used row space
    n_mt = ws.n_mtiles.item()
    mt = ws.mtiles[:n_mt].cpu()
    assert n_mt == t_pad // 128 + sum((c + 127) // 128 for c in counts.tolist())
    assert mt[:, 3].sum().item() == T + counts.sum().item()
    assert (mt[: t_pad // 128, 2] == 8).all()


# ------------------------------------------------------------------ full layer
_MODULES = {}


def _module(dt, dev, seed=0, ffn_impl=None):
    from unimoe_audio_b200 import DCMoE
    key = (dt, seed)
    if key not in _MODULES:
        W = O.make_weights(seed=seed, dtype=dt)
        with torch.device("meta"):
            m = DCMoE(dict(O.DEFAULT_CONFIG))
        m = m.to(dt).to_empty(device=dev)
        m.load_state_dict({k: v.to(dev) for k, v in W.items()})
        _MODULES[key] = (m.eval(), W)
    m, W = _MODULES[key]
    m.ffn_impl = ffn_impl
    return m, W


class _env:
    """Temporarily set a process environment variable (the library reads its tuning switches with getenv per call)."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = os.environ.get(self.name)
        os.environ[self.name] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop(self.name, None)
        else:
            os.environ[self.name] = self.old


def _check_layer(out, ref, dt):
    rtol = 1e-5 if dt == torch.float32 else 1e-2
    a, b = out.float().cpu(), ref.float()
    scale = b.abs().max().item()
    err = (a - b).abs()
    bound = rtol * b.abs() + rtol * scale
    assert (err <= bound).all(), f"max err {err.max().item():.3e} (scale {scale:.3e})"
    rel_fro = (a - b).norm().item() / b.norm().item()
    assert rel_fro <= (3e-6 if dt == torch.float32 else 6e-3), rel_fro


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_layer_config1_matches_reference_golden(dname, dev):
    """BASELINE.json config 1 (1 x 512 tokens) against the fixture produced by the unmodified reference."""
    g = np.load(os.path.join(GOLD, f"layer_{dname}_c1.npz"))
    dt = DT[dname]
    m, W = _module(dt, dev, seed=int(g["weight_seed"]))
    x = torch.randn(1, 512, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(dt)
    # identical router logits on both sides: feed the reference's logits
    ref_logits = torch.from_numpy(g["full_router_logits"]).to(dt).to(dev)
    out = m(x.to(dev), None, None, router_logits=ref_logits)
    torch.cuda.synchronize()
    assert [o.dtype for o in out] == [dt, dt, torch.int64, torch.int32, dt, torch.float32]
    assert out[0].shape == (1, 512, 2048) and out[5].dim() == 0
    assert np.array_equal(out[2].cpu().numpy(), g["dynamic_top_k"])
    assert np.array_equal(out[3].cpu().numpy(), g["expert_mask"])
    assert np.array_equal(out[4].float().cpu().numpy(), g["global_weight"])
    np.testing.assert_allclose(out[5].item(), float(g["aux_loss"]), rtol=1e-5 if dname == "fp32" else 2e-3)
    final = out[0].float().cpu().reshape(512, 2048)
    _check_layer(final[::4], torch.from_numpy(g["final_rows"]), dt)
    # and with our own gate projection: logits within tolerance of the reference's
    out2 = m(x.to(dev), None, None)
    lg_err = (out2[1].float().cpu() - torch.from_numpy(g["full_router_logits"])).abs().max().item()
    assert lg_err <= (2e-2 if dname == "bf16" else 2e-5), lg_err


@pytest.mark.parametrize("dname,B,S,masked", [("bf16", 2, 300, False), ("bf16", 1, 1, False), ("bf16", 3, 171, True),
                                              ("fp32", 1, 130, True), ("fp32", 2, 64, False)])
def test_layer_matches_oracle_ragged_and_masked(dname, B, S, masked, dev):
    dt = DT[dname]
    m, W = _module(dt, dev, seed=1)
    gen = torch.Generator().manual_seed(100 + B * S)
    x = torch.randn(B, S, 2048, generator=gen).to(dt)
    am = (torch.rand(B, S, generator=gen) > 0.3) if masked else None
    out = m(x.to(dev), am.to(dev) if am is not None else None, None)
    torch.cuda.synchronize()
    ref = O.forward(x, W, am, logits=out[1].cpu())          # identical logits on both sides
    assert torch.equal(out[2].cpu(), ref.dynamic_top_k)
    assert torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    np.testing.assert_allclose(out[5].item(), ref.aux_loss.item(), rtol=1e-5 if dname == "fp32" else 2e-3)
    _check_layer(out[0].reshape(B * S, 2048), ref.final_hidden_states.reshape(B * S, 2048), dt)
    # our gate vs torch's
    lg_ref = torch.nn.functional.linear(x.reshape(-1, 2048), W[O.GATE].to(dt)).float()
    assert (out[1].float().cpu() - lg_ref).abs().max().item() <= (2e-2 if dname == "bf16" else 2e-5)


def test_layer_all_tokens_masked_and_empty_input(dev):
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=1)
    x = torch.randn(1, 40, 2048, generator=torch.Generator().manual_seed(5)).to(dt)
    am = torch.zeros(1, 40, dtype=torch.bool)
    out = m(x.to(dev), am.to(dev), None)
    torch.cuda.synchronize()
    assert (out[3][:, :9] == 0).all() and (out[3][:, 9:] == 1).all()   # only shared experts (core.py:286-291)
    ref = O.forward(x, W, am, logits=out[1].cpu())
    _check_layer(out[0].reshape(40, 2048), ref.final_hidden_states.reshape(40, 2048), dt)
    out0 = m(torch.zeros(1, 0, 2048, dtype=dt, device=dev), None, None)
    assert out0[0].shape == (1, 0, 2048) and out0[3].shape == (0, 11)


def test_tcgen05_ffn_agrees_with_cuda_core_ffn(dev):
    """Two independent implementations of the grouped FFN on the same packed rows (bf16)."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=2)
    x = torch.randn(4, 640, 2048, generator=torch.Generator().manual_seed(9)).to(dt).to(dev)
    out_tc = m(x, None, None)
    y_tc = m.last_workspace.y.clone()
    m.ffn_impl = 1
    out_cc = m(x, None, None)
    torch.cuda.synchronize()
    assert torch.equal(out_tc[3], out_cc[3])
    a, b = out_tc[0].float(), out_cc[0].float()
    assert (a - b).abs().max().item() <= 2e-2 * b.abs().max().item()
    assert ((a - b).norm() / b.norm()).item() < 4e-3


def test_layer_config2_size_properties_and_full_parity(dev):
    """BASELINE.json config 2 (8 x 2048 tokens, bf16): size-independent properties + oracle parity on ALL 16,384 rows."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=0)
    B, S = 8, 2048
    x = torch.randn(B, S, 2048, generator=torch.Generator().manual_seed(1236)).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    final, logits, top_k, mask, gw, aux = out
    T = B * S
    ws = m.last_workspace
    # routing decisions bit-exact vs the oracle for all 16384 tokens (identical logits)
    k2, m2, gw2, aux2 = R.route(logits.cpu())
    assert torch.equal(top_k.cpu(), k2) and torch.equal(mask.cpu(), m2) and torch.equal(gw.cpu(), gw2)
    np.testing.assert_allclose(aux.item(), aux2.item(), rtol=2e-3)
    # histogram / prefix-sum invariants
    counts = ws.counts.cpu()
    assert torch.equal(counts.long(), m2[:, :8].sum(0))
    assert (mask[:, :9].sum(1).cpu() == top_k.cpu()).all()          # k_t experts selected per token (incl. null)
    slot = ws.slot_of.cpu()
    used = slot[slot >= 0]
    assert used.numel() == counts.sum().item() and used.unique().numel() == used.numel()   # a permutation
    rt = ws.row_token.cpu()
    tok_ids = torch.arange(T).unsqueeze(1).expand(T, 8)[slot >= 0]
    assert torch.equal(rt[used.long()].long(), tok_ids)             # permute o inverse = identity
    assert torch.isfinite(final.float()).all()
    # determinism: a second forward is bitwise identical (no atomics anywhere)
    out_b = m(x.to(dev), None, None)
    assert torch.equal(out_b[0], final) and torch.equal(out_b[5], aux)
    # oracle parity of the FFN output on every row (the oracle gathers rows per expert: seconds at this size)
    ref = O.forward(x, W, None, logits=logits.cpu())
    assert torch.equal(mask.cpu(), ref.expert_mask) and torch.equal(counts.long(), ref.counts)
    _check_layer(final.reshape(T, 2048), ref.final_hidden_states.reshape(T, 2048), dt)


def test_bf16_layer_is_at_least_as_accurate_as_the_reference_arithmetic(dev):
    """Both the reference's bf16 path and ours are rounded versions of the same real-valued layer.  Against an fp32
    evaluation (same bf16 weights, same routing) our error must not exceed the reference arithmetic's own error."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=4)
    x = torch.randn(1, 384, 2048, generator=torch.Generator().manual_seed(21)).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    logits = out[1].cpu()
    ref_bf16 = O.forward(x, W, None, logits=logits).final_hidden_states.float().reshape(-1, 2048)
    # fp32 "truth": identical routing decisions and weights (taken from the bf16 run), fp32 arithmetic in the FFNs
    W32 = {k: v.float() for k, v in W.items()}
    gw = out[4].cpu().float()
    mask = out[3].cpu()
    xf = x.float().reshape(-1, 2048)
    truth = torch.zeros_like(xf)
    for e in range(8):
        idx = torch.nonzero(mask[:, e], as_tuple=True)[0]
        if idx.numel():
            y = O._ffn(xf[idx], W32[O.ROUTED.format(e=e, proj="gate_proj")], W32[O.ROUTED.format(e=e, proj="up_proj")],
                       W32[O.ROUTED.format(e=e, proj="down_proj")])
            truth.index_add_(0, idx, gw[idx, e, None] * y)
    for e in range(2):
        y = O._ffn(xf, W32[O.SHARED.format(e=e, proj="gate_proj")], W32[O.SHARED.format(e=e, proj="up_proj")],
                   W32[O.SHARED.format(e=e, proj="down_proj")])
        truth += gw[:, 9 + e, None] * y
    ours = out[0].float().cpu().reshape(-1, 2048)
    err_ours = ((ours - truth).norm() / truth.norm()).item()
    err_ref = ((ref_bf16 - truth).norm() / truth.norm()).item()
    assert err_ours <= 1.05 * err_ref + 1e-4, (err_ours, err_ref)


@pytest.mark.parametrize("B,S", [(2, 1), (4, 8), (1, 200)])
def test_cuda_graph_replay_matches_eager(B, S, dev):
    """Decode-sized calls replayed from a CUDA graph (the forward is launch-only) equal the eager forward bitwise."""
    from unimoe_audio_b200.host import GraphedDCMoE
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=1)
    g = GraphedDCMoE(m, B, S, dt, device=dev)
    gen = torch.Generator().manual_seed(B * 100 + S)
    for it in range(3):                                   # replays with different inputs, same graph
        x = torch.randn(B, S, 2048, generator=gen).to(dt).to(dev)
        got = [t.clone() for t in g(x)]
        ref = m(x, None, None)
        torch.cuda.synchronize()
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
    o = O.forward(x.cpu(), W, None, logits=ref[1].cpu())
    assert torch.equal(ref[3].cpu(), o.expert_mask)
    _check_layer(ref[0].reshape(B * S, 2048), o.final_hidden_states.reshape(B * S, 2048), dt)


@pytest.mark.parametrize("T,masked", [(1, False), (2, False), (16, False), (17, True), (32, False), (33, False), (64, True)])
def test_decode_sized_weight_streaming_ffn_matches_large_tiles(T, masked, dev):
    """T <= 64 runs the weight-streaming tcgen05 GEMMs (ffn_tcgen05_stream.cu: 16-column granules divided evenly over
    the SMs, weights as the MMA M operand, register-direct stores).  With one GEMM-2 accumulator the K order and fp32
    accumulation are those of the 128x256 tiles: the layer output equals the CTA-pair kernel's (always large tiles)
    bit for bit.  The default (four GEMM-2 accumulators) differs by fp32 reassociation only.  Both match the oracle."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=4)
    g = torch.Generator().manual_seed(500 + T)
    x = torch.randn(T, 1, 2048, generator=g).to(dt).to(dev)
    mask = None
    if masked:
        mask = (torch.rand(T, 1, generator=g) > 0.3).to(torch.int64).to(dev)
    m.ffn_impl = 0
    out_small = [t.clone() for t in m(x, mask, None)]
    with _env("DCMOE_FFN_STREAM_KSPLIT", "0"):          # GEMM-2 with one accumulator: same summation order
        out_exact = [t.clone() for t in m(x, mask, None)]
    with _env("DCMOE_FFN_STREAM", "0"):                 # the 128 x 256 tile kernel at this size
        out_large = m(x, mask, None)
    torch.cuda.synchronize()
    m.ffn_impl = None
    assert torch.equal(out_small[3], out_large[3])
    assert torch.equal(out_exact[0], out_large[0])
    # default: four GEMM-2 accumulators summed in the epilogue -> fp32 reassociation only
    a, b = out_small[0].float(), out_large[0].float()
    assert (a - b).abs().max().item() <= 2.0 ** -7 * b.abs().max().item()
    assert (a != b).float().mean().item() < 0.05
    ref = O.forward(x.cpu(), W, None if mask is None else mask.cpu(), logits=out_small[1].cpu())
    assert torch.equal(out_small[3].cpu(), ref.expert_mask)
    _check_layer(out_small[0].reshape(T, 2048), ref.final_hidden_states.reshape(T, 2048), dt)


def test_stack_of_layers_chained_config3_shape(dev):
    """BASELINE.json config 3 in miniature: several DCMoE layers with independent weights, activations chained
    (RMS-normalised between layers as the decoder does before the MoE, model.py:239-241), one shared workspace.
    Every layer is checked against the oracle on the activations it actually saw."""
    dt = torch.bfloat16
    T = 768
    x = torch.randn(1, T, 2048, generator=torch.Generator().manual_seed(77)).to(dt).to(dev)
    hidden = x
    for layer_seed in (0, 1, 2):
        m, W = _module(dt, dev, seed=layer_seed)
        normed = (hidden.float() * torch.rsqrt(hidden.float().pow(2).mean(-1, keepdim=True) + 1e-6)).to(dt)
        out = m(normed, None, None)
        torch.cuda.synchronize()
        ref = O.forward(normed.cpu(), W, None, logits=out[1].cpu())
        assert torch.equal(out[3].cpu(), ref.expert_mask) and torch.equal(out[2].cpu(), ref.dynamic_top_k)
        _check_layer(out[0].reshape(T, 2048), ref.final_hidden_states.reshape(T, 2048), dt)
        hidden = hidden + out[0]                                    # residual (model.py:242)


@pytest.mark.parametrize("top_p", [0.5, 0.95])
def test_topp_sweep_with_skewed_router_config5_shape(top_p, dev):
    """BASELINE.json config 5 in miniature: Top-P sweep with a skewed router (bias linspace(+2, -2) on the 9
    dynamic logits -> hot expert 0), single GPU and 4 virtual expert-parallel ranks, against the oracle."""
    from unimoe_audio_b200 import DCMoE
    from unimoe_audio_b200.ep import LocalRanks
    dt = torch.bfloat16
    cfg = dict(O.DEFAULT_CONFIG, mlp_dynamic_top_p=top_p)
    W = O.make_weights(seed=5, dtype=dt)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    m.load_state_dict({k: v.to(dev) for k, v in W.items()})
    T = 1024
    gen = torch.Generator().manual_seed(int(top_p * 100))
    x = torch.randn(1, T, 2048, generator=gen).to(dt)
    logits = torch.randn(T, 11, generator=gen) * 0.9
    logits[:, :9] += torch.linspace(2.0, -2.0, 9)
    logits = logits.to(dt)
    out = m(x.to(dev), None, None, router_logits=logits.to(dev))
    torch.cuda.synchronize()
    ref = O.forward(x, W, None, cfg=cfg, logits=logits)
    assert torch.equal(out[2].cpu(), ref.dynamic_top_k) and torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    counts = m.last_workspace.counts.cpu().long()
    assert torch.equal(counts, ref.counts.long())
    assert counts[0] > 2 * counts[7]                                  # the skew really loads expert 0
    _check_layer(out[0].reshape(T, 2048), ref.final_hidden_states.reshape(T, 2048), dt)
    # 4 expert-parallel virtual ranks on the same tokens (router fed the same logits slices)
    lr = LocalRanks(m, 4)
    xs = [x[:, r * 256:(r + 1) * 256].to(dev).contiguous() for r in range(4)]
    for r, ep in enumerate(lr.ranks):
        ep._forced_logits = logits[r * 256:(r + 1) * 256].to(dev).contiguous()
    outs = lr.forward(xs)
    torch.cuda.synchronize()
    got = torch.cat([o[0][0] for o in outs])
    assert torch.equal(got, out[0][0])


def test_fused_residual_add(dev):
    """`residual=` fuses the decoder layer's `residual + mlp(...)` (model.py:242) into the combine pass."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=1)
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(2, 200, 2048, generator=gen).to(dt).to(dev)
    res = torch.randn(2, 200, 2048, generator=gen).to(dt).to(dev)
    plain = m(x, None, None)
    fused = m(x, None, None, residual=res)
    torch.cuda.synchronize()
    expect = (res.float() + plain[0].float())
    # fused adds in fp32 before the single rounding: at least as close to the fp32 sum as rounding the layer first
    err_fused = (fused[0].float() - expect).abs().max().item()
    err_plain = ((res + plain[0]).float() - expect).abs().max().item()
    assert err_fused <= err_plain + 1e-6 and err_fused <= 2e-2
    assert torch.equal(fused[3], plain[3])


@pytest.mark.parametrize("max_ctas", [0, 100, 37])
def test_weight_streaming_ffn_on_a_capped_grid(max_ctas, dev):
    """dcmoe_grouped_ffn impl = 3 (weight-streaming tcgen05 GEMMs) called through the C ABI with a grid cap: the
    granule split changes with the CTA count, h and y must not (bit-equal to the large-tile kernel); a cap too
    small for the widest segment is refused."""
    from unimoe_audio_b200 import ops
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=2)
    x = torch.randn(9, 1, 2048, generator=torch.Generator().manual_seed(77)).to(dt).to(dev)
    with _env("DCMOE_FFN_STREAM", "0"):                 # reference: the 128 x 256 tile kernel
        m(x, None, None)
    torch.cuda.synchronize()
    ws = m.last_workspace
    h_ref, y_ref = ws.h.clone(), ws.y.clone()
    ws.h.zero_(); ws.y.zero_()
    phase_bits = max_ctas << 8
    if max_ctas == 37:      # 37 CTAs / 9 groups = 4 per group -> 172 / 4 = 43 granules > 16: refused
        with pytest.raises(RuntimeError):
            ops.grouped_ffn(x.reshape(9, 2048), m._w13, m._w2, ws, 3, phase=phase_bits)
        return
    with _env("DCMOE_FFN_STREAM_KSPLIT", "0"):
        ops.grouped_ffn(x.reshape(9, 2048), m._w13, m._w2, ws, 3, phase=phase_bits)
    torch.cuda.synchronize()
    m.ffn_impl = None
    mt = ws.mtiles[: int(ws.n_mtiles.item())].cpu()
    for a_row, out_row, group, n in mt.tolist():
        assert torch.equal(ws.h[out_row:out_row + n], h_ref[out_row:out_row + n])
        assert torch.equal(ws.y[out_row:out_row + n], y_ref[out_row:out_row + n])


@pytest.mark.parametrize("T,masked", [(1, False), (2, False), (17, True), (64, False), (33, True)])
def test_fused_decode_front_end_equals_three_kernel_path(T, masked, dev):
    """dcmoe_front_small (router + plan + permute in one launch, T <= 64) must reproduce the three-kernel path bit
    for bit: routing outputs, counts, segment bases, tile table, slots, scales, gathered rows, aux."""
    from unimoe_audio_b200 import ops
    dt = torch.bfloat16
    gen = torch.Generator().manual_seed(900 + T)
    x = torch.randn(T, 2048, generator=gen).to(dt).to(dev)
    wg = (torch.randn(11, 2048, generator=gen) * 0.02).to(dt).to(dev)
    am = (torch.rand(T, generator=gen) > 0.3).to(dev) if masked else None
    ws_a = ops.Workspace(_dims(), dt, T, dev)
    ws_b = ops.Workspace(_dims(), dt, T, dev)
    rb = ops.front_small(x, wg, ws_b, attention_mask=am)
    # the gate projection splits K differently in the two kernels (fp32 partial sums in another order), so the
    # logits may differ by one bf16 ulp: compare them with a tolerance and run the three-kernel path on the front
    # end's logits -- everything downstream must then be identical
    r0 = ops.router(x, wg, ws_a, attention_mask=am)
    assert (r0[0].float() - rb[0].float()).abs().max().item() <= 2e-2
    ra = ops.router(None, None, ws_a, logits_in=rb[0], attention_mask=am)
    ops.plan(ws_a)
    ops.permute(x, ra[2], ra[3], ws_a)
    torch.cuda.synchronize()
    for a, b in zip(ra, rb):
        assert torch.equal(a, b)
    assert torch.equal(ws_a.counts, ws_b.counts) and torch.equal(ws_a.seg_base, ws_b.seg_base)
    n = ws_a.n_mtiles.item()
    assert n == ws_b.n_mtiles.item() and torch.equal(ws_a.mtiles[:n], ws_b.mtiles[:n])
    assert torch.equal(ws_a.slot_of, ws_b.slot_of)
    used = int(ws_a.seg_base[-1].item())
    valid = torch.zeros(used, dtype=torch.bool, device=dev)
    mt = ws_a.mtiles[:n]
    for i in range(n):
        valid[mt[i, 1].item(): mt[i, 1].item() + mt[i, 3].item()] = True
    assert torch.equal(ws_a.row_token[:used][valid], ws_b.row_token[:used][valid])
    assert torch.equal(ws_a.row_scale[:used][valid], ws_b.row_scale[:used][valid])
    t_pad = ws_a.t_pad
    routed = valid[t_pad:]
    assert torch.equal(ws_a.x_packed[: used - t_pad][routed], ws_b.x_packed[: used - t_pad][routed])
    assert torch.equal(ws_a.aux_loss, ws_b.aux_loss) or abs(ws_a.aux_loss.item() - ws_b.aux_loss.item()) <= 1e-6 * abs(ws_a.aux_loss.item())


@pytest.mark.parametrize("dname", ["bf16", "fp32"])
@pytest.mark.parametrize("T", [8, 700])
def test_non_reference_dimensions_generic_kernels(dname, T, dev):
    """A smaller layer (hidden 512, 4 routed + 1 null + 2 shared experts, I_d 256): exercises the run-time expert-count
    paths of the router (templates <0, 0>), other tile counts in the GEMMs and the C oracle's general-n arithmetic."""
    from unimoe_audio_b200 import DCMoE
    dt = DT[dname]
    cfg = dict(O.DEFAULT_CONFIG, hidden_size=512, mlp_dynamic_expert_num=4, mlp_dynamic_null_expert_num=1,
               mlp_fixed_expert_num=2, dynamic_intermediate_size=256, shared_intermediate_size=128, mlp_dynamic_top_p=0.6)
    W = O.make_weights(cfg, seed=9, dtype=dt)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    m.load_state_dict({k: v.to(dev) for k, v in W.items()})
    gen = torch.Generator().manual_seed(T)
    x = torch.randn(1, T, 512, generator=gen).to(dt)
    am = torch.rand(1, T, generator=gen) > 0.2
    out = m(x.to(dev), am.to(dev), None)
    torch.cuda.synchronize()
    ref = O.forward(x, W, am, cfg=cfg, logits=out[1].cpu())
    assert out[1].shape == (T, 7) and out[3].shape == (T, 7)
    assert torch.equal(out[2].cpu(), ref.dynamic_top_k)
    assert torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    np.testing.assert_allclose(out[5].item(), ref.aux_loss.item(), rtol=1e-5 if dname == "fp32" else 2e-3)
    _check_layer(out[0].reshape(T, 512), ref.final_hidden_states.reshape(T, 512), dt)
    lg_ref = torch.nn.functional.linear(x.reshape(-1, 512), W[O.GATE].to(dt)).float()
    assert (out[1].float().cpu() - lg_ref).abs().max().item() <= (2e-2 if dname == "bf16" else 2e-5)


@pytest.mark.parametrize("dims", [
    dict(hidden_size=256, mlp_dynamic_expert_num=2, mlp_dynamic_null_expert_num=0, mlp_fixed_expert_num=1,
         dynamic_intermediate_size=64, shared_intermediate_size=64, mlp_dynamic_top_p=0.7),
    dict(hidden_size=768, mlp_dynamic_expert_num=6, mlp_dynamic_null_expert_num=2, mlp_fixed_expert_num=2,
         dynamic_intermediate_size=192, shared_intermediate_size=96, mlp_dynamic_top_p=0.8),
    dict(hidden_size=1024, mlp_dynamic_expert_num=12, mlp_dynamic_null_expert_num=2, mlp_fixed_expert_num=2,
         dynamic_intermediate_size=384, shared_intermediate_size=192, mlp_dynamic_top_p=0.5),
], ids=["h256-2+0+1", "h768-6+2+2", "h1024-12+2+2"])
@pytest.mark.parametrize("T", [5, 300])
def test_odd_configurations(dims, T, dev):
    """Edge configurations of the constructor contract: no null expert, a single shared expert, 16 router columns,
    intermediate sizes that end in a half tile -- bf16, single GPU and (when divisible) 2 virtual EP ranks."""
    from unimoe_audio_b200 import DCMoE
    from unimoe_audio_b200.ep import LocalRanks
    dt = torch.bfloat16
    cfg = dict(O.DEFAULT_CONFIG, **dims)
    H = cfg["hidden_size"]
    W = O.make_weights(cfg, seed=13, dtype=dt)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    m.load_state_dict({k: v.to(dev) for k, v in W.items()})
    gen = torch.Generator().manual_seed(T + H)
    x = torch.randn(1, T, H, generator=gen).to(dt)
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    ref = O.forward(x, W, None, cfg=cfg, logits=out[1].cpu())
    assert torch.equal(out[2].cpu(), ref.dynamic_top_k) and torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    _check_layer(out[0].reshape(T, H), ref.final_hidden_states.reshape(T, H), dt)
    lg_ref = torch.nn.functional.linear(x.reshape(-1, H), W[O.GATE].to(dt)).float()
    assert (out[1].float().cpu() - lg_ref).abs().max().item() <= 2e-2
    if T >= 16:
        half = T // 2
        xs = [x[:, :half].to(dev).contiguous(), x[:, half:].to(dev).contiguous()]
        outs = LocalRanks(m, 2).forward(xs)
        torch.cuda.synchronize()
        got = torch.cat([o[0][0] for o in outs])
        # the EP ranks run the large-T router on their half; logits may differ from the one-GPU call by a bf16 ulp
        # (different K split), so compare against the oracle on the logits the ranks actually produced
        lg = torch.cat([o[1] for o in outs]).cpu()
        ref2 = O.forward(x, W, None, cfg=cfg, logits=lg)
        assert torch.equal(torch.cat([o[3] for o in outs]).cpu(), ref2.expert_mask)
        _check_layer(got, ref2.final_hidden_states.reshape(T, H), dt)


@pytest.mark.parametrize("dname,T,H", [("bf16", 16384, 2048), ("fp32", 4096, 2048), ("bf16", 1, 2048), ("bf16", 77, 4096),
                                       ("fp32", 33, 768), ("bf16", 5, 256)])
def test_rmsnorm_matches_oracle(dname, T, H, dev):
    """dcmoe_rmsnorm (decoder-layer post_attention_layernorm, model.py:240) against the oracle restatement of
    Qwen2RMSNorm.  The sum of squares is reduced in a different order than torch's, so rsqrt may differ in the last
    fp32 bit: fp32 outputs within 2e-6 relative; bf16 outputs equal except rare (< 0.2 %) flips of a rounding."""
    from unimoe_audio_b200 import ops
    dt = torch.bfloat16 if dname == "bf16" else torch.float32
    dims = ops.LayerDims(hidden_size=H)
    g = torch.Generator().manual_seed(1000 + T)
    x = (torch.randn(T, H, generator=g) * 2.3).to(dt)
    w = (1.0 + 0.2 * torch.randn(H, generator=g)).to(dt)
    ref = O.rmsnorm(x, w, 1e-6)
    out = ops.rmsnorm(x.to(dev), w.to(dev), 1e-6, dims).cpu()
    assert out.dtype == dt and out.shape == x.shape
    if dt == torch.float32:
        assert torch.allclose(out, ref, rtol=2e-6, atol=0)
    else:
        a, b = out.float(), ref.float()
        bad = a != b
        assert bad.float().mean().item() < 2e-3
        # a one-ulp flip of the first rounding D(x * inv) passes through D(weight * .): at most two bf16 ulps
        assert ((a - b).abs() <= b.abs() * 2.0 ** -6 + 1e-30).all()


@pytest.mark.parametrize("dname", ["fp32", "bf16"])
def test_post_attention_moe_matches_reference_golden(dname, dev):
    """PostAttentionMoE = rmsnorm + DCMoE + fused residual (model.py:239-242) on the fixture produced by the
    unmodified reference block and transformers' Qwen2RMSNorm (tools/make_golden_glue.py)."""
    from unimoe_audio_b200 import PostAttentionMoE
    g = np.load(os.path.join(GOLD, f"glue_{dname}.npz"))
    dt = torch.bfloat16 if dname == "bf16" else torch.float32
    m, W = _module(dt, dev, seed=int(g["weight_seed"]))
    blk = PostAttentionMoE({"rms_norm_eps": float(g["eps"])}, mlp=m).to(dev).eval()
    nw = (1.0 + 0.1 * torch.randn(2048, generator=torch.Generator().manual_seed(int(g["norm_weight_seed"])))).to(dt)
    blk.post_attention_layernorm.weight.data = nw.to(dev)
    x = (torch.randn(1, 128, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))) * float(g["x_scale"])).to(dt)
    assert sorted(k for k in blk.state_dict() if not k.startswith("mlp.")) == ["post_attention_layernorm.weight"]
    out = blk(x.to(dev), None, None)
    torch.cuda.synchronize()
    # the MoE saw rmsnorm(x): its logits are a function of the normed rows
    final = out[0].float().cpu().reshape(128, 2048)
    rtol = 1e-5 if dname == "fp32" else 1e-2
    scale = float(np.abs(g["final_rows"]).max())
    ref_rows = torch.from_numpy(g["final_rows"])
    err = (final[::4] - ref_rows).abs()
    assert (err <= rtol * ref_rows.abs() + rtol * scale).all(), err.max().item()
    # routing decisions on the GPU's own logits are exact (oracle); against the golden they may differ only where
    # a one-ulp logit difference flips a near-tie, which these seeds do not contain
    o = O.forward(O.rmsnorm(x, nw, float(g["eps"])), W, None, logits=out[1].cpu())
    assert torch.equal(out[3].cpu(), o.expert_mask)
    assert (out[3].cpu().numpy() != g["expert_mask"]).mean() < 5e-3


# ------------------------------------------------------------------ fixed top-k routing (mlp_dynamic_top_p == 0)
@pytest.mark.parametrize("dname", ["fp32", "bf16"])
@pytest.mark.parametrize("k", [2, 3])
def test_fixed_topk_router_matches_reference_golden_bit_exact(dname, k, dev):
    """core.py:254-257: with mlp_dynamic_top_p == 0 every token selects mlp_dynamic_top_k dynamic experts."""
    from unimoe_audio_b200 import ops
    g = np.load(os.path.join(GOLD, f"routek_{dname}_k{k}.npz"))
    dt = DT[dname]
    dims = ops.LayerDims(top_p=0.0, fixed_top_k=k)
    for case in ("iid", "ties"):
        logits = torch.from_numpy(g[f"{case}_logits"]).to(dt)
        am = torch.from_numpy(g["ties_attention_mask"]).to(dev) if case == "ties" else None
        ws = ops.Workspace(dims, dt, logits.shape[0], dev)
        lg, top_k, mask, gw = ops.router(None, None, ws, logits_in=logits.to(dev).contiguous(), attention_mask=am)
        ops.plan(ws)
        torch.cuda.synchronize()
        assert np.array_equal(top_k.cpu().numpy(), g[f"{case}_dynamic_top_k"])
        assert np.array_equal(mask.cpu().numpy(), g[f"{case}_expert_mask"])
        assert np.array_equal(gw.float().cpu().numpy(), g[f"{case}_global_weight"])
        np.testing.assert_allclose(ws.aux_loss.item(), float(g[f"{case}_aux_loss"]), rtol=1e-5 if dt == torch.float32 else 2e-3)


@pytest.mark.parametrize("T", [256, 40])
def test_fixed_topk_layer_matches_reference_golden_and_oracle(T, dev):
    """Whole layer in fixed top-k mode (bf16, k = 2): the T = 256 case is the reference fixture; T = 40 takes the
    decode-sized kernels (fused front end + weight-streaming GEMMs) and is checked against the oracle."""
    from unimoe_audio_b200 import DCMoE
    g = np.load(os.path.join(GOLD, "layerk_bf16_k2.npz"))
    dt = torch.bfloat16
    W = O.make_weights(seed=int(g["weight_seed"]), dtype=dt)
    cfg = dict(O.DEFAULT_CONFIG, mlp_dynamic_top_p=0, mlp_dynamic_top_k=2)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev)
    m.load_state_dict({kk: v.to(dev) for kk, v in W.items()})
    m.eval()
    x = torch.randn(1, 256, 2048, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(dt)[:, :T]
    out = m(x.to(dev), None, None)
    torch.cuda.synchronize()
    assert out[2].dtype == torch.int32 and (out[2] == 2).all()
    ref = O.forward(x, W, cfg=cfg, logits=out[1].cpu())
    assert torch.equal(out[3].cpu(), ref.expert_mask)
    assert torch.equal(out[4].cpu(), ref.global_weight)
    _check_layer(out[0].reshape(T, 2048), ref.final_hidden_states.reshape(T, 2048), dt)
    if T == 256:
        assert (out[3].cpu().numpy() != g["expert_mask"]).mean() < 5e-3      # GPU logits may differ by one bf16 ulp
        scale = float(np.abs(g["final_rows"]).max())
        err = (out[0].float().cpu().reshape(256, 2048)[::4] - torch.from_numpy(g["final_rows"])).abs()
        assert (err <= 2e-2 * scale).float().mean() > 0.995


def test_decode_sized_kernels_random_sweep(dev):
    """Every token count 1..64 once (random masks on a third of them, peaky and flat routers so that 1..9 weight
    groups are hit): the decode-sized path (fused front end / gather kernel / weight-streaming GEMMs) must equal
    the large-tile path bit for bit."""
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=6)
    g = torch.Generator().manual_seed(2024)
    groups_seen = set()
    for T in range(1, 65):
        scale = (0.3, 1.0, 6.0)[T % 3]          # flat -> many experts per token, peaky -> one
        x = (torch.randn(T, 1, 2048, generator=g) * scale).to(dt).to(dev)
        mask = (torch.rand(T, 1, generator=g) > 0.3).to(torch.int64).to(dev) if T % 3 == 0 else None
        m.ffn_impl = 0
        m.use_front_small = True
        with _env("DCMOE_FFN_STREAM_KSPLIT", "0"):
            a = [t.clone() for t in m(x, mask, None)]
        groups_seen.add(int(m.last_workspace.n_mtiles.item()))
        m.use_front_small = False
        with _env("DCMOE_FFN_STREAM", "0"):             # three-kernel front end + the 128 x 256 tile kernel
            b = m(x, mask, None, router_logits=a[1])
        torch.cuda.synchronize()
        m.ffn_impl = None
        m.use_front_small = True
        assert torch.equal(a[3], b[3]) and torch.equal(a[2], b[2]) and torch.equal(a[4], b[4]), T
        assert torch.equal(a[0], b[0]), T
        assert abs(a[5].item() - b[5].item()) <= 1e-6 * max(1.0, abs(b[5].item())), T
    assert len(groups_seen) >= 5, groups_seen


def test_row_capacity_hint_overflow_is_bounded_and_detected(dev):
    """a row_capacity below the worst case is a supported setting.  Rows that do not fit are DROPPED (slot -1,
    never written past the buffers), plan.overflow is raised on the host by the next forward / check_overflow(), and a
    sufficient factor reproduces the worst-case workspace's result bit for bit."""
    from unimoe_audio_b200 import _lib
    dt = torch.bfloat16
    m, W = _module(dt, dev, seed=0)
    x = torch.randn(2, 700, 2048, generator=torch.Generator().manual_seed(77)).to(dt).to(dev)
    T = 1400
    ref = [t.clone() for t in m(x, None, None)]
    try:
        m.row_capacity_factor = 5.5                 # mean r ~ 3.6 at top_p 0.7: fits
        cap = m.effective_row_capacity(T)
        assert 0 < cap < 1408 + 8 * T + 1024
        out = m(x, None, None)
        m.check_overflow()
        assert m.last_workspace.row_capacity == cap and m.last_workspace.reduced
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
        m.row_capacity_factor = 1.0                 # far too small: most routed rows are dropped
        out = m(x, None, None)
        torch.cuda.synchronize()
        ws = m.last_workspace
        limit = ws.row_capacity // 128 * 128
        slot = ws.slot_of.cpu()
        assert int(slot.max()) < limit and (slot == -1).sum() > (ref[3][:, :8] == 0).sum().cpu()
        assert ws.overflowed and torch.isfinite(out[0].float()).all()
        with pytest.raises(_lib.DcmoeError, match="overflow"):
            m.check_overflow()
        for i in (1, 2, 3, 4):                      # the routing outputs do not depend on the workspace
            assert torch.equal(out[i], ref[i])
    finally:
        m.row_capacity_factor = None
    out = m(x, None, None)
    assert torch.equal(out[0], ref[0])


def test_layer_on_a_device_that_is_not_current(dev):
    """launches follow the INPUT's device (as PyTorch ops do), not the process's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from unimoe_audio_b200 import DCMoE
    dt = torch.bfloat16
    W = O.make_weights(seed=0, dtype=dt)
    d1 = torch.device("cuda:1")
    with torch.device("meta"):
        m1 = DCMoE(dict(O.DEFAULT_CONFIG))
    m1 = m1.to(dt).to_empty(device=d1).eval()
    m1.load_state_dict({k: v.to(d1) for k, v in W.items()})
    m0, _ = _module(dt, dev, seed=0)
    x = torch.randn(1, 300, 2048, generator=torch.Generator().manual_seed(3)).to(dt)
    assert torch.cuda.current_device() == 0
    o1 = m1(x.to(d1), None, None)                   # current device stays cuda:0
    o0 = m0(x.to(dev), None, None)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.cuda.current_device() == 0
    for a, b in zip(o1, o0):
        assert a.device == d1 and torch.equal(a.cpu(), b.cpu())


def test_router_exponential_is_the_correctly_rounded_one(dev):
    """csrc/exp_fast.cuh (float-pair evaluation + Ziv's rounding test, no FP64 on the accepted path) against the
    definition (float)exp((double)x) on the device and against the C statement of the same arithmetic: bit-equal on
    4M random differences of bf16 logits, on a dense sweep of (-80, 0], on the fallback triggers the exhaustive CPU run
    found, and outside the fast domain."""
    import ctypes

    from unimoe_audio_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(11)
    a = (torch.randn(4_000_000, generator=gen) * 1.5).to(torch.bfloat16).float()
    b = (torch.randn(4_000_000, generator=gen) * 1.5).to(torch.bfloat16).float()
    xs = [-(a - b).abs(), -torch.rand(2_000_000, generator=gen) * 80.0, -torch.logspace(-40, 1.9, 500_000),
          torch.tensor([0.0, -0.0, -80.0, -79.99999, -87.0, -104.0, -200.0, 1.0, 3.5, float("-inf"), float("nan"), -1e-45, -1.17e-38])]
    x = torch.cat(xs).to(dev).contiguous()
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.dcmoe_test_exp(x.data_ptr(), y0.data_ptr(), x.numel(), 0, st), "dcmoe_test_exp")
    _lib.check(lib.dcmoe_test_exp(x.data_ptr(), y1.data_ptr(), x.numel(), 1, st), "dcmoe_test_exp")
    torch.cuda.synchronize()
    same = (y0.view(torch.int32) == y1.view(torch.int32)) | (torch.isnan(y0) & torch.isnan(y1))
    assert bool(same.all()), x[~same][:8]
    # the C statement (oracle/exp_fast.h) on a sample: same accept/fallback decisions are not observable, the values are
    fb = ctypes.c_int(0)
    xc, yc = x.cpu(), y0.cpu()
    n_fb = 0
    for i in range(0, xc.numel(), 4001):
        v = R.lib().dcmoe_oracle_exp_fast(float(xc[i]), ctypes.byref(fb))
        n_fb += fb.value
        if not fb.value:
            assert np.float32(v).tobytes() == np.float32(yc[i]).tobytes(), float(xc[i])
        assert np.float32(R.lib().dcmoe_oracle_exp_cr(float(xc[i]))).tobytes() == np.float32(yc[i]).tobytes() or np.isnan(yc[i])
    assert n_fb < 50

