"""Expert-parallel kernels on ONE GPU: R virtual ranks in one process (unimoe_audio_b200.ep.LocalRanks) run the
same ep_plan / ep_dispatch / grouped FFN / ep_combine kernels as the multi-process path, with peer pointers that
happen to be local.  EP output must equal the single-GPU output on the concatenated batch (the
maths is row independent) -- here bitwise, because rows land in the same canonical order."""
import pytest
import torch

from oracle import dcmoe_oracle as O
from oracle import route_oracle_c as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from unimoe_audio_b200 import DCMoE
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    W = O.make_weights(seed=3, dtype=dt)
    with torch.device("meta"):
        m = DCMoE(dict(O.DEFAULT_CONFIG))
    m = m.to(dt).to_empty(device=dev).eval()
    m.load_state_dict({k: v.to(dev) for k, v in W.items()})
    return m, W, dev, dt


@pytest.mark.parametrize("split", [False, True], ids=["serial", "overlap-schedule"])
@pytest.mark.parametrize("world,tokens", [(2, [300, 300]), (4, [257, 16, 1, 130]), (8, [64] * 8), (2, [1000, 24])])
def test_local_ranks_match_single_gpu(setup, world, tokens, split):
    from unimoe_audio_b200.ep import LocalRanks, ep_layout
    m, W, dev, dt = setup
    gen = torch.Generator().manual_seed(sum(tokens) + world)
    xs = [torch.randn(1, t, 2048, generator=gen).to(dt).to(dev) for t in tokens]
    lr = LocalRanks(m, world, split=split)
    outs = lr.forward(xs)
    torch.cuda.synchronize()
    x_all = torch.cat(xs, dim=1)
    ref = m(x_all, None, None)
    torch.cuda.synchronize()
    off = 0
    for r, t in enumerate(tokens):
        o = outs[r]
        assert torch.equal(o[1], ref[1][off:off + t])            # logits
        assert torch.equal(o[2], ref[2][off:off + t])            # dynamic_top_k
        assert torch.equal(o[3], ref[3][off:off + t])            # expert_mask
        assert torch.equal(o[4], ref[4][off:off + t])            # global_weight
        assert torch.equal(o[0][0], ref[0][0, off:off + t]), f"rank {r} output differs"
        off += t
    # device ep_plan == host mirror; global per-expert counts are exact
    ac = lr.all_counts.cpu().tolist()
    total = ref[3][:, :8].sum(0).cpu().tolist()
    assert [sum(ac[r][e] for r in range(world)) for e in range(8)] == total
    n_loc = 8 // world
    for r, ep in enumerate(lr.ranks):
        dest_base, dest_tpad, seg, tot = ep_layout(ac, r, 8)
        meta = ep.ws.ep_meta.cpu().tolist()
        assert meta[:8] == dest_base and meta[16:24] == dest_tpad
        assert ep.ws.seg_base.cpu().tolist()[: n_loc + 1] == seg
        assert ep.ws.counts.cpu().tolist()[:n_loc] == tot


@pytest.mark.parametrize("world,tokens", [(2, [300, 300]), (4, [257, 96, 65, 130]), (8, [80] * 8)])
def test_weight_gather_path_matches_single_gpu(setup, world, tokens):
    """Weight-gather expert parallelism on virtual ranks: every rank copies all ranks' packs into its staging pack
    (dcmoe_ep_fetch_weights) and runs the single-GPU forward on its own rows -> each rank's 6-tuple (except the aux loss,
    which is per rank) is bit-equal to the single-GPU layer on the concatenated batch.  Called twice with different
    inputs: the second call uses the other staging slot and must not see stale weights or rows.  (More than 64 tokens
    per rank -- the path is meant for thousands: at T <= 64 the layer picks the decode-sized GEMMs, whose GEMM-2 sums
    four accumulators, i.e. equal up to fp32 reassociation only.)"""
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    lr = LocalRanks(m, world)
    for rep in range(2):
        gen = torch.Generator().manual_seed(sum(tokens) + world + 1000 * rep)
        xs = [torch.randn(1, t, 2048, generator=gen).to(dt).to(dev) for t in tokens]
        outs = lr.gather_forward(xs)
        torch.cuda.synchronize()
        ref = m(torch.cat(xs, dim=1), None, None)
        torch.cuda.synchronize()
        off = 0
        for r, t in enumerate(tokens):
            o = outs[r]
            for i in (1, 2, 3, 4):
                assert torch.equal(o[i], ref[i][off:off + t]), (rep, r, i)
            assert torch.equal(o[0][0], ref[0][0, off:off + t]), f"call {rep} rank {r} output differs"
            off += t
    # the staging pack of rank 0 now equals the single-GPU pack, group by group
    m.pack_weights()
    ctx = lr.ranks[0].ctx
    assert any(torch.equal(w13, m._w13) and torch.equal(w2, m._w2) for w13, w2 in ctx.stage)


def test_dispatch_workspace_is_reused_across_token_counts(setup):
    """a later call with MORE tokens than the first one must not reuse buffers sized for the first call
    (the fp32 partial sums of the overlapped combine used to be allocated once)."""
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    lr = LocalRanks(m, 2, split=True)
    for tokens in ([256, 256], [1024, 1000], [100, 90]):
        gen = torch.Generator().manual_seed(sum(tokens))
        xs = [torch.randn(1, t, 2048, generator=gen).to(dt).to(dev) for t in tokens]
        outs = lr.forward(xs)
        torch.cuda.synchronize()
        ref = m(torch.cat(xs, dim=1), None, None)
        off = 0
        for r, t in enumerate(tokens):
            assert torch.equal(outs[r][0][0], ref[0][0, off:off + t]), (tokens, r)
            assert outs[r][0].shape[1] == t and lr.ranks[r].ws.partial.shape[0] >= t
            off += t


def test_local_ranks_oracle_parity(setup):
    """EP result against the CPU oracle on the concatenated batch (identical logits)."""
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    gen = torch.Generator().manual_seed(77)
    xs = [torch.randn(1, 96, 2048, generator=gen).to(dt).to(dev) for _ in range(4)]
    outs = LocalRanks(m, 4).forward(xs)
    torch.cuda.synchronize()
    x_all = torch.cat([x.cpu() for x in xs], dim=1)
    logits = torch.cat([o[1] for o in outs]).cpu()
    ref = O.forward(x_all, W, None, logits=logits)
    got = torch.cat([o[0][0] for o in outs]).float().cpu()
    exp = ref.final_hidden_states[0].float()
    assert torch.equal(torch.cat([o[3] for o in outs]).cpu(), ref.expert_mask)
    assert ((got - exp).norm() / exp.norm()).item() < 6e-3
    assert (got - exp).abs().max().item() <= 1e-2 * exp.abs().max().item() + 1e-2 * exp.abs().max().item()


def test_multi_gpu_ep_if_available(setup):
    """Real multi-process EP over NCCL + cudaIpc peer memory; runs only where >= 2 GPUs are visible."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(root, "tools", "ep_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "EP_CHECK_OK" in res.stdout


@pytest.mark.parametrize("world,T,masked", [(2, 1, False), (2, 8, True), (4, 2, False), (4, 16, False), (8, 1, False),
                                            (8, 8, True), (8, 5, False)])
def test_decode_sized_expert_parallel_path_matches_single_gpu(setup, world, T, masked):
    """world * T <= 64: replicated routing of the gathered tokens, every virtual rank streams only its experts' weights
    (+ the shared pair), the combine gathers the routed rows from the owners' y.  Same kernels and row space as a
    single-GPU call on the gathered tokens -> every rank's slice of the 6-tuple is bit-equal to it."""
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    gen = torch.Generator().manual_seed(100 * world + T)
    xs = [(torch.randn(1, T, 2048, generator=gen) * (0.5 + r % 3)).to(dt).to(dev) for r in range(world)]
    ams = [(torch.rand(1, T, generator=gen) > 0.3).to(torch.int64).to(dev) for _ in range(world)] if masked else None
    lr = LocalRanks(m, world)
    assert lr.ranks[0].decode_applicable(T, dt)
    outs = lr.decode_forward(xs, ams)
    torch.cuda.synchronize()
    x_all = torch.cat(xs, dim=1)
    am_all = None if ams is None else torch.cat(ams, dim=1)
    ref = m(x_all, am_all, None)
    torch.cuda.synchronize()
    for r in range(world):
        o, sl = outs[r], slice(r * T, (r + 1) * T)
        for i in (1, 2, 3, 4):
            assert torch.equal(o[i], ref[i][sl]), (r, i)
        assert torch.equal(o[0][0], ref[0][0, sl]), f"rank {r} output differs"
        assert torch.equal(o[5], ref[5])
    # every rank only wrote the y rows of its own experts and of the shared pair
    ws0 = lr.ranks[0]._dws
    mt = ws0.mtiles[: int(ws0.n_mtiles.item())].cpu().tolist()
    assert {g for _a, _o, g, _n in mt} >= {8}


@pytest.mark.parametrize("world,T", [(2, 2), (8, 8), (4, 5)])
def test_decode_sized_calls_with_replicated_experts_match_single_gpu(setup, world, T):
    """Decode policy "replicate" (the default of ExpertParallelDCMoE): every rank keeps a resident copy of all experts'
    packs (fetched once) and runs its own tokens with no exchange -- rows bit-equal to the single-GPU forward."""
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    gen = torch.Generator().manual_seed(7 * world + T)
    lr = LocalRanks(m, world)
    for _ in range(2):          # the second call reuses the resident packs
        xs = [(torch.randn(1, T, 2048, generator=gen) * (0.5 + r % 3)).to(dt).to(dev) for r in range(world)]
        outs = lr.resident_forward(xs)
        torch.cuda.synchronize()
        for r in range(world):
            ref = m(xs[r], None, None)
            for i in range(6):
                assert torch.equal(outs[r][i], ref[i]), (r, i)
    assert all(ep._resident is not None for ep in lr.ranks)


@pytest.mark.parametrize("world", [2, 4])
def test_training_recipe_branches_ride_on_the_weight_gather_path(setup, world):
    """token_drop / aux_balance_weight under expert parallelism (V2 recipe: ep_size > 1 with token_drop, training.sh:55-59):
    per-rank computations on the rank's own tokens (core.py:293-329 runs before the exchange), so every rank's result
    must equal the single-GPU layer called on that rank's tokens alone -- all six fields, the aux loss included."""
    from unimoe_audio_b200 import DCMoE
    from unimoe_audio_b200.ep import LocalRanks
    m, W, dev, dt = setup
    with torch.device("meta"):
        md = DCMoE(dict(O.DEFAULT_CONFIG, token_drop=True, drop_policy="probs", capacity_factor=1.0, min_capacity=8))
    md = md.to(dt).to_empty(device=dev).eval()
    md.load_state_dict({k: v.to(dev) for k, v in W.items()})
    lr = LocalRanks(md, world)
    gen = torch.Generator().manual_seed(31 + world)
    tokens = [300, 200, 150, 90][:world]
    xs = [torch.randn(1, t, 2048, generator=gen).to(dt).to(dev) for t in tokens]
    ws = [torch.randint(1, 4, (1, t), generator=gen).to(dev) for t in tokens]
    for ep in lr.ranks:
        ep.pack_local_weights()
        ep.context(dt, dev)
    for ep in lr.ranks:
        ep.set_peer_weights([q._wbuf["w13"].ptr for q in lr.ranks], [q._wbuf["w2"].ptr for q in lr.ranks])
    for r, ep in enumerate(lr.ranks):
        out = ep.gather_forward(xs[r], None, ws[r])
        torch.cuda.synchronize()
        ref = md(xs[r], None, ws[r])
        torch.cuda.synchronize()
        for i in range(6):
            assert torch.equal(out[i], ref[i]), (r, i)
        plain = R.route(ref[1].cpu())[1]
        assert int(ref[3][:, :9].sum()) < int(plain[:, :9].sum())         # the capacity did drop something
