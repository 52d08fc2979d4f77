"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol that
include/dcmoe_b200.h declares, validates configurations, and refuses to compute without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from unimoe_audio_b200 import DCMoE, _lib, ops
from unimoe_audio_b200.ops import LayerDims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dcmoe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcmoe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert "dcmoe_router" in declared and "dcmoe_grouped_ffn" in declared and len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dcmoe_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.dcmoe_abi_version() == _lib.ABI_VERSION == 3


def test_query_sizes_reference_config():
    dims = LayerDims()
    sz, lay = ops.query_sizes(dims, torch.bfloat16, 8 * 2048)
    assert sz.n_blocks == 1024 and sz.t_pad == 16384
    assert sz.row_capacity == 16384 + 8 * 16384 + 8 * 128
    assert sz.max_mtiles == sz.row_capacity // 128
    assert lay.total == sz.plan_bytes and lay.mtiles % 16 == 0
    sz0, _ = ops.query_sizes(dims, torch.float32, 0)
    assert sz0.n_blocks == 0 and sz0.t_pad == 0
    sz1, _ = ops.query_sizes(dims, torch.float32, 1)
    assert sz1.n_blocks == 1 and sz1.t_pad == 128


@pytest.mark.parametrize("bad", [
    dict(n_fix=3, shared_intermediate_size=1376),            # shared pack must equal the routed size
    dict(hidden_size=2000),                                   # H % 256
    dict(top_p=0.0),                                          # fixed top-k mode needs fixed_top_k >= 1
    dict(top_p=1.5),                                          # top_p outside [0, 1]
    dict(n_real=14, n_null=1, n_fix=2),                       # more than 16 router columns
])
def test_invalid_configs_are_rejected_with_a_message(bad):
    dims = LayerDims(**bad)
    with pytest.raises(_lib.DcmoeError) as ei:
        ops.query_sizes(dims, torch.bfloat16, 128)
    assert len(str(ei.value)) > 20


def test_module_mirrors_reference_interface():
    cfg = dict(hidden_size=2048, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=2752,
               shared_intermediate_size=1376, router_jitter_noise=0.01, ep_size=1, token_drop=False)
    with torch.device("meta"):
        m = DCMoE(cfg)
    keys = set(m.state_dict().keys())
    assert "gate.weight" in keys
    assert "fixed_real_moe.1.down_proj.weight" in keys
    assert "dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.7.up_proj.weight" in keys
    assert len(keys) == 1 + 2 * 3 + 8 * 3
    assert m.num_experts == 11 and m.mlp_dynamic_expert_num == 9
    assert m.dynamic_real_moe.deepspeed_moe.ep_group is None
    with torch.device("meta"):
        md = DCMoE(dict(cfg, token_drop=True))                         # drop_policy defaults to "probs" (core.py:230)
    assert md.token_drop and md.drop_policy == "probs"
    with pytest.raises(NotImplementedError, match="NaN"):             # the reference's "position" policy is broken
        DCMoE(dict(cfg, token_drop=True, drop_policy="position"))
    with pytest.raises(ValueError):
        DCMoE(dict(cfg, mlp_dynamic_top_p=0))                      # fixed top-k routing needs mlp_dynamic_top_k >= 1
    mk = DCMoE(dict(cfg, mlp_dynamic_top_p=0, mlp_dynamic_top_k=2))   # core.py:256-257
    assert mk.dims.fixed_top_k == 2 and mk.dims.c_config(torch.bfloat16).fixed_top_k == 2
    assert m.dims.c_config(torch.bfloat16).fixed_top_k == 0


def test_no_cpu_fallback():
    cfg = dict(hidden_size=256, mlp_dynamic_expert_num=8, mlp_dynamic_null_expert_num=1, mlp_dynamic_top_p=0.7,
               mlp_dynamic_top_k=0.0, mlp_fixed_expert_num=2, dynamic_intermediate_size=128,
               shared_intermediate_size=64, router_jitter_noise=0.01)
    m = DCMoE(cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 256), None, None)
    if not torch.cuda.is_available():
        lib = _lib.load()
        c = m.dims.c_config(torch.float32)
        rc = lib.dcmoe_combine(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, c, None, ctypes.c_void_p(16), None)
        assert rc == -2 and b"no CPU fallback" in lib.dcmoe_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unimoe_audio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "route_oracle" not in text.replace("oracle/route_oracle.c", ""), f


@pytest.mark.parametrize("gpg", [172, 128, 16, 40])          # I_d / 16, H / 16 of the reference config; small generic sizes
@pytest.mark.parametrize("grid", [148, 132, 100, 37, 9])
def test_weight_streaming_partition_covers_every_granule_once(gpg, grid):
    """The CTA -> (m-tile, granule range) map of the decode-sized GEMMs (host mirror of cta_segment): every granule of
    every hit weight group is computed by exactly one CTA, the CTAs of a group differ by at most one granule, and no
    segment is wider than the bound the launcher sizes its ring stages with."""
    for n_m in range(1, 10):
        if n_m > grid:
            continue
        seg = ops.stream_segments(n_m, gpg, grid)
        assert len(seg) == grid
        cover = {m: [0] * gpg for m in range(n_m)}
        sizes = {m: [] for m in range(n_m)}
        for m, g0, ng in seg:
            if ng == 0:
                continue
            assert 0 <= m < n_m and 0 <= g0 and g0 + ng <= gpg
            for g in range(g0, g0 + ng):
                cover[m][g] += 1
            sizes[m].append(ng)
        assert all(c == 1 for m in cover for c in cover[m])
        ctas = [sum(1 for mm, _g, _n in seg if mm == m) for m in range(n_m)]
        assert max(ctas) - min(ctas) <= 1                              # CTAs dealt evenly to the groups
        for m in range(n_m):
            assert max(sizes[m]) - min(sizes[m]) <= 1
        bound = -(-gpg // (grid // 9)) if grid >= 9 else gpg            # launcher: ceil(gpg / floor(grid / G)), G = 9
        assert max(max(v) for v in sizes.values()) <= max(bound, 1)


def test_expert_capacity_matches_the_reference_formula():
    """core.py:170-175 + :306-308, against the reference's own tensor arithmetic restated with torch on the CPU."""
    dims = LayerDims()
    for T, cf, mn in [(1024, 1.0, 8), (1024, 2.0, 8), (16384, 3.0, 8), (16384, 6.0, 8), (7, 1.0, 8), (40, 1.0, 8),
                      (262144, 1.25, 4), (999, 0.37, 8), (9, 1.0, 0)]:
        cap = torch.ceil((T / 9) * torch.tensor(cf)).to(torch.int64)
        if cap < torch.tensor(mn):
            cap = torch.tensor(mn).to(torch.int64)
        want = min(int(cap), T)
        assert ops.expert_capacity(dims, T, cf, mn) == want, (T, cf, mn)
