"""Host-side logic of the expert-parallel path on CPU: the layout arithmetic (mirror of ep_plan_kernel) and a
world_size-2 gloo run of the count exchange that drives it."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unimoe_audio_b200.ep import ep_layout


def _random_counts(world, seed):
    g = torch.Generator().manual_seed(seed)
    T = torch.randint(1, 3000, (world,), generator=g).tolist()
    return [[int(torch.randint(0, T[r] + 1, (1,), generator=g)) for _ in range(8)] + [T[r]] for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_ep_layout_segments_partition_the_owner_row_space(world):
    n_real, n_loc = 8, 8 // world
    ac = _random_counts(world, 100 + world)
    layouts = [ep_layout(ac, r, n_real) for r in range(world)]
    for e in range(n_real):
        owner = e // n_loc
        _, _, seg, tot = layouts[owner]
        l = e - owner * n_loc
        # the ranks' destination ranges tile [seg[l], seg[l] + total_e) in rank order, without gaps or overlap
        cur = seg[l]
        for r in range(world):
            base = layouts[r][0][e]
            assert base == cur
            assert layouts[r][1][e] == (ac[owner][n_real] + 127) // 128 * 128
            cur += ac[r][e]
        assert cur == seg[l] + tot[l] and cur <= seg[l + 1] and seg[l] % 128 == 0
    for r in range(world):
        assert layouts[r][2][0] == (ac[r][n_real] + 127) // 128 * 128     # routed rows start after the shared rows


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ac = _random_counts(world, 7)
    mine = torch.tensor(ac[rank], dtype=torch.int32)
    gathered = [torch.empty(9, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, mine)                         # the one data-path collective besides the barriers
    table = torch.stack(gathered).tolist()
    flag = torch.zeros(1, dtype=torch.int32)
    dist.all_reduce(flag)                                   # barrier stand-in
    ret[rank] = (table, ep_layout(table, rank, 8))
    dist.destroy_process_group()


def test_gloo_world2_count_exchange_gives_consistent_layouts():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29517, ret), nprocs=world, join=True)
    t0, l0 = ret[0]
    t1, l1 = ret[1]
    assert t0 == t1 == _random_counts(world, 7)
    # rank 1's rows of every expert start right after rank 0's
    for e in range(8):
        assert l1[0][e] == l0[0][e] + t0[0][e]
        assert l0[1][e] == l1[1][e]


def test_choose_path_thresholds():
    """Host logic that picks the expert-parallel path of a call: decode (replicated tokens) for world * T <= 64 in bf16,
    weight gather from gather_min_tokens on, the token dispatch in between; the mode switch forces one of the two."""
    from unimoe_audio_b200.ep import choose_path
    bf, f32 = torch.bfloat16, torch.float32
    assert choose_path(8, 8, bf) == "decode" and choose_path(9, 8, bf) == "dispatch"
    assert choose_path(8, 8, f32) == "dispatch"                       # the decode kernels are bf16 only
    assert choose_path(8, 8, bf, decode_ok=False) == "dispatch"
    assert choose_path(8191, 4, bf) == "dispatch" and choose_path(8192, 4, bf) == "gather"
    assert choose_path(100, 2, bf, mode="gather") == "gather" and choose_path(1 << 20, 2, bf, mode="dispatch") == "dispatch"
    assert choose_path(32, 2, bf, mode="gather") == "decode"          # decode-sized calls keep their own path
    assert choose_path(0, 2, bf) == "dispatch"
    # every configs[3] point (64 x 4096 tokens over 2 / 4 / 8 ranks) is a weight-gather call
    assert all(choose_path(64 * 4096 // r, r, bf) == "gather" for r in (2, 4, 8))


def test_decode_policy_and_extended_branches_choose_their_paths(monkeypatch):
    """Host logic of ExpertParallelDCMoE that needs no GPU: the decode policy read from DCMOE_EP_DECODE (default: a resident
    replica of the remote experts, "exchange" = replicated tokens, "0" = off), and the launch count bench.py reports."""
    from unimoe_audio_b200 import DCMoE
    from unimoe_audio_b200.ep import ExpertParallelDCMoE
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01)
    with torch.device("meta"):
        m = DCMoE(cfg)
    for env, policy, on in ((None, "replicate", True), ("exchange", "exchange", True), ("1", "exchange", True), ("0", "off", False)):
        if env is None:
            monkeypatch.delenv("DCMOE_EP_DECODE", raising=False)
        else:
            monkeypatch.setenv("DCMOE_EP_DECODE", env)
        ep = ExpertParallelDCMoE(m, group=None, rank=1, world=4)
        assert ep.decode_policy == policy and ep.decode_mode is on and ep.n_loc == 2
    with pytest.raises(ValueError):
        ExpertParallelDCMoE(m, group=None, rank=0, world=3)                 # 8 experts over 3 ranks (core.py:505)
