"""Loader for the UNMODIFIED reference DCMoE block (test infrastructure only).

This file is part of ``oracle/`` -- it is a *checker*, never a product path.  Only
``tests/``, ``tools/make_golden.py``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline leg may import from ``oracle/``.

It imports ``/root/reference/utils/UniMoE_Audio_core.py`` with ZERO edits, following the recipe
of SURVEY.md section 8(c):

  1. a stub ``deepspeed`` package exposing only the symbols core.py:7-11 imports
     (deepspeed 0.15.1 is not installed here; its ``einsum`` is ``torch.einsum``);
  2. a shim ``utils.UniMoE_Audio_utils`` module whose ``compress_matrix`` /
     ``decompress_matrix`` are produced by exec-ing the verbatim source span
     utils/UniMoE_Audio_utils.py:436-523 read from the reference tree at run time
     (nothing is copied into this repo);
  3. the identity ``_AllToAll.forward`` the reference installs at
     utils/UniMoE_Audio_utils.py:332-335,:429.

``/root/reference`` only exists in the build container, so everything here is used to
(a) validate ``oracle/dcmoe_oracle.py`` and (b) generate the committed fixtures under
``tests/golden/`` (``tools/make_golden.py``).  Nothing that runs on the GPU box calls this.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
from types import SimpleNamespace

import torch

REFERENCE_ROOT = os.environ.get("DCMOE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "UniMoE_Audio_core.py"))


def _install_deepspeed_stub() -> None:
    if "deepspeed" in sys.modules and getattr(sys.modules["deepspeed"], "__dcmoe_stub__", False):
        return
    ds = types.ModuleType("deepspeed")
    ds.__dcmoe_stub__ = True
    ds.__path__ = []  # mark as package

    comm = types.ModuleType("deepspeed.comm")

    class ReduceOp:  # noqa: D401 - mirrors deepspeed.comm.ReduceOp names only
        MAX = "max"
        AVG = "avg"
        SUM = "sum"

    def all_reduce(*_a, **_k):
        raise RuntimeError("deepspeed.comm.all_reduce stub: ep_group must be None in the oracle")

    comm.ReduceOp = ReduceOp
    comm.all_reduce = all_reduce
    comm.ProcessGroup = object

    utils = types.ModuleType("deepspeed.utils")
    utils.__path__ = []
    groups = types.ModuleType("deepspeed.utils.groups")
    groups.mpu = None
    groups._get_expert_parallel_group_dict = lambda: {}
    groups._create_expert_and_data_parallel = lambda *a, **k: None
    groups._create_expert_data_and_model_parallel = lambda *a, **k: None
    groups._get_expert_parallel_group = lambda name: None
    utils.groups = groups
    utils.log_dist = lambda *a, **k: None

    timer = types.ModuleType("deepspeed.utils.timer")

    class SynchronizedWallClockTimer:
        def __call__(self, name):
            raise RuntimeError("timers are never used (wall_clock_breakdown=False)")

    timer.SynchronizedWallClockTimer = SynchronizedWallClockTimer

    moe = types.ModuleType("deepspeed.moe")
    moe.__path__ = []
    sharded = types.ModuleType("deepspeed.moe.sharded_moe")
    sharded.FIRST_ALLTOALL_TIMER = "1st_a2a"
    sharded.MOE_TIMER = "moe"
    sharded.SECOND_ALLTOALL_TIMER = "2nd_a2a"
    sharded.einsum = torch.einsum  # deepspeed 0.15.1 default USE_EINSUM=True

    def gumbel_rsample(*_a, **_k):
        raise RuntimeError("gumbel_rsample is training-only")

    sharded.gumbel_rsample = gumbel_rsample

    class _AllToAll(torch.autograd.Function):
        @staticmethod
        def forward(ctx, group, input):  # replaced below by the reference's identity patch
            raise RuntimeError("unpatched _AllToAll")

        @staticmethod
        def backward(ctx, *grad):
            return (None, *grad)

    sharded._AllToAll = _AllToAll

    class MOELayer(torch.nn.Module):
        pass

    sharded.MOELayer = MOELayer
    experts = types.ModuleType("deepspeed.moe.experts")

    class Experts(torch.nn.Module):
        pass

    experts.Experts = Experts
    layer = types.ModuleType("deepspeed.moe.layer")

    class MoE(torch.nn.Module):
        pass

    layer.MoE = MoE

    ds.comm = comm
    ds.utils = utils
    ds.moe = moe
    moe.sharded_moe = sharded
    moe.experts = experts
    moe.layer = layer
    for name, mod in {
        "deepspeed": ds,
        "deepspeed.comm": comm,
        "deepspeed.utils": utils,
        "deepspeed.utils.groups": groups,
        "deepspeed.utils.timer": timer,
        "deepspeed.moe": moe,
        "deepspeed.moe.sharded_moe": sharded,
        "deepspeed.moe.experts": experts,
        "deepspeed.moe.layer": layer,
    }.items():
        sys.modules[name] = mod


def _source_span(path: str, first: int, last: int) -> str:
    with open(path, "r", encoding="utf-8") as fh:
        lines = fh.readlines()
    return "".join(lines[first - 1 : last])


_CORE = None


def load_reference_core():
    """Return the reference ``UniMoE_Audio_core`` module, imported from its original path."""
    global _CORE
    if _CORE is not None:
        return _CORE
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_deepspeed_stub()
    utils_py = os.path.join(REFERENCE_ROOT, "utils", "UniMoE_Audio_utils.py")

    pkg = types.ModuleType("_dcmoe_ref_utils")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "utils")]
    sys.modules["_dcmoe_ref_utils"] = pkg

    shim = types.ModuleType("_dcmoe_ref_utils.UniMoE_Audio_utils")
    shim.__dict__.update({"torch": torch})
    # verbatim compress_matrix / decompress_matrix (utils.py:436-523), executed from the tree
    exec(compile(_source_span(utils_py, 436, 523), utils_py, "exec"), shim.__dict__)
    # verbatim identity all-to-all (utils.py:332-335), installed as utils.py:429 does
    ns = {"Any": object, "Tensor": torch.Tensor, "dist": sys.modules["deepspeed.comm"]}
    exec(compile(_source_span(utils_py, 332, 335), utils_py, "exec"), ns)
    sys.modules["deepspeed.moe.sharded_moe"]._AllToAll.forward = staticmethod(ns["_AllToAll_forward"])
    sys.modules["_dcmoe_ref_utils.UniMoE_Audio_utils"] = shim

    core_py = os.path.join(REFERENCE_ROOT, "utils", "UniMoE_Audio_core.py")
    spec = importlib.util.spec_from_file_location("_dcmoe_ref_utils.UniMoE_Audio_core", core_py)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    _CORE = mod
    return mod


def reference_text_config() -> dict:
    with open(os.path.join(REFERENCE_ROOT, "utils", "config.json"), "r", encoding="utf-8") as fh:
        return json.load(fh)["text_config"]


def build_reference_block(config: dict | None = None, dtype=torch.float32, seed: int = 0, std: float = 0.02):
    """Instantiate ``UniMoEAudioSparseMoeBlock`` (core.py:196) with N(0, std^2) weights, eval mode."""
    import contextlib
    import io

    core = load_reference_core()
    cfg = dict(reference_text_config() if config is None else config)
    with contextlib.redirect_stdout(io.StringIO()):
        block = core.UniMoEAudioSparseMoeBlock(SimpleNamespace(**cfg))
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, p in sorted(block.named_parameters(), key=lambda kv: kv[0]):
            p.copy_(torch.randn(p.shape, generator=gen, dtype=torch.float32) * std)
    return block.to(dtype).eval()
