"""unimoe_audio_b200 -- B200-native DCMoE layer forward (drop-in for UniMoE-Audio's
``UniMoEAudioSparseMoeBlock``, reference utils/UniMoE_Audio_core.py:196-358).

Only what the hot path needs lives here: ``csrc/`` (hand-written sm_100a CUDA + the C ABI of
``include/dcmoe_b200.h``), ``ops`` (torch-facing wrappers that pass device pointers and the current
stream through ctypes) and ``dcmoe`` (the host-side mirror of the reference module interface).
"""
from .dcmoe import DCMoE, PostAttentionMoE, UniMoEAudioSparseMoeBlock  # noqa: F401
from .ops import LayerDims, Workspace  # noqa: F401

__all__ = ["DCMoE", "PostAttentionMoE", "UniMoEAudioSparseMoeBlock", "LayerDims", "Workspace"]
