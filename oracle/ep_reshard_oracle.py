"""Restatement of the reference's expert-parallel checkpoint re-sharding rule (test infrastructure only).

ORACLE -- NOT A PRODUCT PATH (see oracle/dcmoe_oracle.py).  Follows
``UniMoEV2-Preview/inference/deepspeed_ep_param_aggregation.py:16-49``: every per-expert file
``layer_{L}_expert_{E}_mp_rank_00_model_states.pt`` holds keys
``model.layers.{L}.mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{E}<rest>``; with
``ep_group_num = source_ep_num // target_ep_size`` the tensor goes to target ``E // ep_group_num`` under the name
with expert id ``E % ep_group_num``; everything in ``mp_rank_00_model_states.pt["module"]`` is copied to every target.
Parity status: the rule is four lines of integer arithmetic in the reference; it is restated here and the product's
``checkpoint.plan_layer_load`` is checked against it on synthetic checkpoints (no golden files are needed).
"""
import re
from typing import Dict, List

EXPERT_FILE = r"layer_(\d+)_expert_(\d+)_mp_rank_00_model_states.pt"
MLP_NAME = r"model\.layers\.(\d+)\.mlp\.dynamic_real_moe\.deepspeed_moe\.experts\.deepspeed_experts\.(\d+)"
RENAME = "model.layers.{layer_id}.mlp.dynamic_real_moe.deepspeed_moe.experts.deepspeed_experts.{new_expert_id}{rest}"


def aggregation_names(module_keys: List[str], expert_files: Dict[str, List[str]], source_ep_num: int,
                      target_ep_size: int) -> List[Dict[str, str]]:
    """For each target rank: {target key -> source key}.  ``expert_files``: file name -> keys inside."""
    assert source_ep_num % target_ep_size == 0
    ep_group_num = source_ep_num // target_ep_size
    target = [{k: k for k in module_keys} for _ in range(target_ep_size)]
    for fname, names in expert_files.items():
        m = re.match(EXPERT_FILE, fname)
        if not m:
            continue
        layer_id, expert_id = int(m.group(1)), int(m.group(2))
        for name in names:
            nm = re.match(MLP_NAME, name)
            assert nm and int(nm.group(1)) == layer_id and int(nm.group(2)) == expert_id
            rest = name[len(nm.group(0)):]
            new = RENAME.format(layer_id=layer_id, new_expert_id=expert_id % ep_group_num, rest=rest)
            assert new not in target[expert_id // ep_group_num]
            target[expert_id // ep_group_num][new] = name
    return target
