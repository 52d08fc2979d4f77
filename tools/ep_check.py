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
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: equal={ok} max_abs_err={err:.3e}", flush=True)
    if rank == 0:
        print("EP_CHECK_OK" if flag.item() == 1 else "EP_CHECK_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
