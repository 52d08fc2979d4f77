"""unimoe_audio_b200.checkpoint on the GPU: layers built straight from DeepSpeed / safetensors checkpoints (one expert
packed at a time) must behave exactly like a DCMoE that loaded the same state dict, single GPU and expert parallel."""
import pytest
import torch

from unimoe_audio_b200 import DCMoE, checkpoint as C
from unimoe_audio_b200.ep import ExpertParallelDCMoE

pytestmark = pytest.mark.gpu

H, ID, IS, NR, NF, L = 256, 128, 64, 8, 2, 3
CFG = dict(hidden_size=H, mlp_dynamic_expert_num=NR, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
           mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=NF, dynamic_intermediate_size=ID, shared_intermediate_size=IS,
           router_jitter_noise=0.01, rms_norm_eps=1e-6)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def world(tmp_path_factory, dev):
    from safetensors.torch import save_file
    dt = torch.bfloat16
    ref = DCMoE(CFG).to(dt)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for _, p in sorted(ref.named_parameters(), key=lambda kv: kv[0]):
            p.copy_((torch.randn(p.shape, generator=g) * 0.05).to(dt))
    sd = {f"model.layers.{L}.mlp.{k}": v.clone() for k, v in ref.state_dict().items()}
    sd[f"model.layers.{L}.post_attention_layernorm.weight"] = (1 + 0.1 * torch.randn(H, generator=g)).to(dt)
    root = tmp_path_factory.mktemp("ckpt")
    hf, ds = root / "hf", root / "ds"
    hf.mkdir(); ds.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(hf / "model.safetensors"))
    module = {k: v for k, v in sd.items() if "deepspeed_experts" not in k}
    torch.save({"module": module}, ds / "mp_rank_00_model_states.pt")
    for e in range(NR):
        torch.save({k: v for k, v in sd.items() if f"deepspeed_experts.{e}." in k}, ds / f"layer_{L}_expert_{e}_mp_rank_00_model_states.pt")
    return dict(ref=ref.to(dev).eval(), hf=str(hf), ds=str(ds), norm=sd[f"model.layers.{L}.post_attention_layernorm.weight"])


@pytest.mark.parametrize("kind", ["hf", "ds"])
@pytest.mark.parametrize("T", [7, 300])
def test_layer_from_checkpoint_equals_layer_from_state_dict(kind, T, world, dev):
    src = C.SafetensorsSource(world["hf"]) if kind == "hf" else C.DeepSpeedSource(world["ds"])
    m = C.load_dcmoe(src, L, CFG, torch.bfloat16, dev)
    assert all(p.numel() == 0 for n, p in m.named_parameters() if "proj" in n)       # only the packed copy exists
    x = torch.randn(1, T, H, generator=torch.Generator().manual_seed(T)).to(torch.bfloat16).to(dev)
    a, b = m(x, None, None), world["ref"](x, None, None)
    torch.cuda.synchronize()
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    blk = C.load_dcmoe(src, L, CFG, torch.bfloat16, dev, with_norm=True)
    assert torch.equal(blk.post_attention_layernorm.weight.cpu(), world["norm"])
    out = blk(x, None, None)
    assert torch.isfinite(out[0].float()).all()


@pytest.mark.parametrize("ep_size", [2, 8])
def test_expert_parallel_packs_from_checkpoint(ep_size, world, dev):
    src = C.DeepSpeedSource(world["ds"])
    for rank in range(ep_size):
        m, w13, w2 = C.load_dcmoe_ep(src, L, CFG, torch.bfloat16, dev, rank, ep_size)
        want = ExpertParallelDCMoE(world["ref"], None, rank=rank, world=ep_size)
        want.pack_local_weights()
        assert torch.equal(w13, want._w13) and torch.equal(w2, want._w2)
        got = ExpertParallelDCMoE(m, None, rank=rank, world=ep_size)
        got.set_packed_local_weights(w13, w2)
        got.pack_local_weights()                      # no-op: packs are already set
        assert got._w13 is w13
        assert torch.equal(m.gate.weight, world["ref"].gate.weight)
