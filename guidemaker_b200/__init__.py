"""guidemaker_b200 -- B200-native engine for GuideMaker's off-target hot path.

Drop-in for the hot-path surface of ``guidemaker.core`` (``PamTarget.find_targets``,
``TargetProcessor.find_unique_near_pam / create_index / get_neighbors / get_control_seqs``);
see DESIGN.md and INTEGRATION.md.  All arithmetic runs in ``lib/libgm_b200.so`` (hand-written
sm_100a CUDA behind the C ABI of ``include/gm_b200.h``); there is no CPU fallback.
"""
from .core import Annotation, PamTarget, TargetProcessor, cfd_score, extend_ambiguous_dna  # noqa: F401

__all__ = ["PamTarget", "TargetProcessor", "Annotation", "cfd_score", "extend_ambiguous_dna", "core"]
__version__ = "0.1.0"
