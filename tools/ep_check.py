"""Multi-process expert-parallel check (launch with torch.distributed.run, one rank per GPU):
the EP output over cudaIpc peer memory must equal the single-GPU DCMoE output on the concatenated batch, bit for bit, on
all three paths (token dispatch, weight gather, decode-sized).  Every iteration uses DIFFERENT inputs, so stale peer
data from an earlier iteration (a missed barrier, a cached peer line) cannot compare equal."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unimoe_audio_b200 import DCMoE  # noqa: E402
from unimoe_audio_b200.ep import ExpertParallelDCMoE  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dt = torch.bfloat16
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01)
    with torch.device("meta"):
        m = DCMoE(cfg)
    m = m.to(dt).to_empty(device=dev).eval()
    gen = torch.Generator(device=dev).manual_seed(0)          # same weights on every rank
    with torch.no_grad():
        for _, p in sorted(m.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32) * 0.02).to(dt))
    g2 = torch.Generator(device=dev).manual_seed(42)          # identical on all ranks

    def batch(T_loc):
        x_all = torch.randn(1, sum(T_loc), 2048, generator=g2, device=dev, dtype=torch.float32).to(dt)
        off = sum(T_loc[:rank])
        return x_all, off, x_all[:, off:off + T_loc[rank]].contiguous()

    def same(out, ref, off, n):
        return (torch.equal(out[0][0], ref[0][0, off:off + n]) and
                all(torch.equal(out[i], ref[i][off:off + n]) for i in (1, 2, 3, 4)))

    ep = ExpertParallelDCMoE(m, dist.group.WORLD)
    ep.check_lockstep = True
    results = {}

    # ---- token dispatch: ragged per-rank token counts, the token counts and the inputs change from call to call ----
    ok = True
    for it, base in enumerate((700, 300, 1500, 700)):
        T_loc = [base + 37 * r for r in range(world)]
        x_all, off, x_mine = batch(T_loc)
        out = ep(x_mine, None, None)
        assert ep.last_path == "dispatch"
        torch.cuda.synchronize()
        ref = m(x_all, None, None)
        torch.cuda.synchronize()
        ok = ok and same(out, ref, off, T_loc[rank])
    results["dispatch"] = ok

    # ---- the same with the NCCL collectives instead of the peer-memory barriers ----
    ok = True
    ep.ctx.use_flags = False
    for it, base in enumerate((500, 900)):
        T_loc = [base + 11 * r for r in range(world)]
        x_all, off, x_mine = batch(T_loc)
        out = ep(x_mine, None, None)
        torch.cuda.synchronize()
        ref = m(x_all, None, None)
        ok = ok and same(out, ref, off, T_loc[rank])
    ep.ctx.use_flags = True
    results["dispatch_nccl"] = ok

    # ---- weight gather: forced at a small size, then above the automatic threshold; back-to-back calls without a
    # host synchronisation in between alternate the two staging slots ----
    ok = True
    ep.mode = "gather"
    ep.check_lockstep = False          # (its host all-gather would serialise the back-to-back calls)
    outs, refs = [], []
    for it, base in enumerate((400, 650, 512)):
        T_loc = [base + 5 * r for r in range(world)]
        x_all, off, x_mine = batch(T_loc)
        outs.append((ep(x_mine, None, None), off, T_loc[rank]))
        assert ep.last_path == "gather"
        refs.append(x_all)
    torch.cuda.synchronize()
    for (out, off, n), x_all in zip(outs, refs):
        ref = m(x_all, None, None)
        ok = ok and same(out, ref, off, n)
    ep.mode = "auto"
    T_loc = [ep.gather_min_tokens + 128 * r for r in range(world)]
    x_all, off, x_mine = batch(T_loc)
    out = ep(x_mine, None, None)
    assert ep.last_path == "gather"
    torch.cuda.synchronize()
    ref = m(x_all, None, None)
    ok = ok and same(out, ref, off, T_loc[rank])
    results["gather"] = ok

    # ---- the reference's own way to ask for expert parallelism: config.ep_size (core.py:505-520) -- the module then
    # holds only this rank's routed experts and DCMoE.forward runs the expert-parallel path by itself ----
    n_loc = 8 // world
    with torch.device("meta"):
        m2 = DCMoE(dict(cfg, ep_size=world))
    m2 = m2.to(dt).to_empty(device=dev).eval()
    sd = m.state_dict()
    pre = "dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts."
    local_sd = {k: v for k, v in sd.items() if not k.startswith(pre)}
    for l in range(n_loc):
        for proj in ("gate_proj", "up_proj", "down_proj"):
            local_sd[f"{pre}{l}.{proj}.weight"] = sd[f"{pre}{rank * n_loc + l}.{proj}.weight"]
    m2.load_state_dict(local_sd)
    assert len(m2.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts) == n_loc
    ok = True
    for base, mode in ((600, "dispatch"), (640, "gather")):
        T_loc = [base + 3 * r for r in range(world)]
        x_all, off, x_mine = batch(T_loc)
        if m2._ep is not None:
            m2._ep.mode = mode
        out2 = m2(x_mine, None, None)
        assert m2._ep.last_path == mode
        torch.cuda.synchronize()
        ref = m(x_all, None, None)
        ok = ok and same(out2, ref, off, T_loc[rank])
    m2._ep.mode = "auto"
    results["ep_size_config"] = ok

    # ---- avg_hidden_states_last (core.py:355-356): all-reduce AVG of the final hidden states over the group ----
    x_all, off, x_eq = batch([256] * world)
    base_out = m2(x_eq, None, None)[0].clone()
    dist.all_reduce(base_out)
    base_out.div_(world)
    m2.avg_hidden_states_last = True
    out3 = m2(x_eq, None, None)
    m2.avg_hidden_states_last = False
    torch.cuda.synchronize()
    results["avg_hidden_states_last"] = torch.equal(out3[0], base_out)

    # ---- decode-sized calls, policy "exchange" (replicated routing of the gathered tokens): bit-equal to one GPU, inputs
    # change every call ----
    ok = True
    ep.decode_policy = "exchange"
    for Td in (1, 4, 64 // world, 2, 2, 2):
        xd_all, off, xd = batch([Td] * world)
        assert ep.decode_applicable(Td, dt)
        od = ep(xd, None, None)
        assert ep.last_path == "decode"
        torch.cuda.synchronize()
        rd = m(xd_all, None, None)
        torch.cuda.synchronize()
        ok = ok and same(od, rd, off, Td) and torch.equal(od[5], rd[5])
    # with a padding mask, and replayed back to back without host synchronisation
    outs = []
    for it in range(4):
        xd_all, off, xd = batch([3] * world)
        am_all = (torch.rand(1, 3 * world, generator=g2, device=dev) > 0.3).to(torch.int64)
        outs.append((ep(xd, am_all[:, off:off + 3].contiguous(), None), xd_all, am_all, off))
    torch.cuda.synchronize()
    for od, xd_all, am_all, off in outs:
        rd = m(xd_all, am_all, None)
        ok = ok and same(od, rd, off, 3)
    results["decode"] = ok

    # ---- decode-sized calls, policy "replicate" (the default): resident copy of every expert's pack, no exchange per
    # call; rows equal the single-GPU forward of the concatenated batch, the aux loss is the one of the local tokens ----
    ep.decode_policy = "replicate"
    ok = True
    for Td in (2, 64 // world, 1, 2):
        xd_all, off, xd = batch([Td] * world)
        od = ep(xd, None, None)
        assert ep.last_path == "decode" and ep._resident is not None
        torch.cuda.synchronize()
        rd = m(xd_all, None, None)
        rl = m(xd, None, None)
        torch.cuda.synchronize()
        ok = ok and same(od, rd, off, Td) and torch.equal(od[5], rl[5])
    results["decode_replicated"] = ok

    all_ok = all(results.values())
    print(f"rank {rank}: " + " ".join(f"{k}={'ok' if v else 'FAILED'}" for k, v in results.items()), flush=True)
    flag = torch.tensor([1 if all_ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("EP_CHECK_OK" if flag.item() == 1 else "EP_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
