"""Multi-process expert-parallel check (launch with torch.distributed.run, one rank per GPU):
EP output over NCCL + cudaIpc peer memory must equal the single-GPU DCMoE output on the concatenated batch."""
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
    T_loc = [700 + 37 * r for r in range(world)]              # ragged per-rank token counts
    g2 = torch.Generator(device=dev).manual_seed(42)
    x_all = torch.randn(1, sum(T_loc), 2048, generator=g2, device=dev, dtype=torch.float32).to(dt)   # identical on all ranks
    off = sum(T_loc[:rank])
    x_mine = x_all[:, off:off + T_loc[rank]].contiguous()
    ep = ExpertParallelDCMoE(m, dist.group.WORLD)
    for it in range(3):                                       # repeated calls reuse the peer buffers
        out = ep(x_mine, None, None)
    torch.cuda.synchronize()
    ref = m(x_all, None, None)
    torch.cuda.synchronize()
    ok = torch.equal(out[0][0], ref[0][0, off:off + T_loc[rank]]) and torch.equal(out[3], ref[3][off:off + T_loc[rank]])
    err = (out[0][0].float() - ref[0][0, off:off + T_loc[rank]].float()).abs().max().item()
    # the reference's own way to ask for expert parallelism: config.ep_size (core.py:505-520) -- the module then
    # holds only this rank's routed experts and DCMoE.forward runs the expert-parallel path by itself
    n_loc = 8 // world
    with torch.device("meta"):
        m2 = DCMoE(dict(cfg, ep_size=world))
    m2 = m2.to(dt).to_empty(device=dev).eval()
    sd = m.state_dict()
    pre = "dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts."
    local = {k: v for k, v in sd.items() if not k.startswith(pre)}
    for l in range(n_loc):
        for proj in ("gate_proj", "up_proj", "down_proj"):
            local[f"{pre}{l}.{proj}.weight"] = sd[f"{pre}{rank * n_loc + l}.{proj}.weight"]
    m2.load_state_dict(local)
    assert len(m2.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts) == n_loc
    out2 = m2(x_mine, None, None)
    torch.cuda.synchronize()
    ok2 = all(torch.equal(a, b) for a, b in zip(out2, out))
    # avg_hidden_states_last (core.py:355-356): all-reduce AVG of the final hidden states over the group
    x_eq = x_all[:, rank * 256:(rank + 1) * 256].contiguous()
    base = m2(x_eq, None, None)[0].clone()
    dist.all_reduce(base)
    base.div_(world)
    m2.avg_hidden_states_last = True
    out3 = m2(x_eq, None, None)
    m2.avg_hidden_states_last = False
    torch.cuda.synchronize()
    ok3 = torch.equal(out3[0], base)
    print(f"rank {rank}: ep_size-configured module equal={ok2} avg_hidden_states_last equal={ok3}", flush=True)
    # decode-sized calls take the replicated-routing path (ExpertParallelDCMoE.decode_forward): bit-equal to one GPU
    ok4 = True
    for Td in (1, 4):
        xd_all = x_all[:, 5000 - Td * world:5000] if x_all.shape[1] >= 5000 else x_all[:, :Td * world]
        xd = xd_all[:, rank * Td:(rank + 1) * Td].contiguous()
        assert ep.decode_applicable(Td, dt)
        for _ in range(3):
            od = ep(xd, None, None)
        torch.cuda.synchronize()
        rd = m(xd_all.contiguous(), None, None)
        torch.cuda.synchronize()
        ok4 = ok4 and torch.equal(od[0][0], rd[0][0, rank * Td:(rank + 1) * Td]) and torch.equal(od[3], rd[3][rank * Td:(rank + 1) * Td])
    print(f"rank {rank}: decode-sized expert-parallel path equal={ok4}", flush=True)
    ok = ok and ok2 and ok3 and ok4
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: equal={ok} max_abs_err={err:.3e}", flush=True)
    if rank == 0:
        print("EP_CHECK_OK" if flag.item() == 1 else "EP_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
