"""Import the LIVE reference classes from /root/reference (build container only).  TEST INFRASTRUCTURE.

/root/reference does not exist on the GPU box, so nothing that runs there imports this module; it is used by
``oracle/make_golden.py`` and by CPU tests that are skipped when the tree is absent.

Recipe: import transformers first, then install two stub modules for imports that
the hot path never calls (``librosa.filters.mel`` is imported by src/hifigan/data.py:6; ``einx.multiply`` is the
outer product used at src/flow_matching/modules/fourier_embed.py:38), then import the reference package.
The reference's own ``from_pretrained`` fails on transformers 5.x (it never calls ``post_init``), so models are
built with ``cls(config)`` + ``load_state_dict``.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
# the unmodified reference files installed by oracle/install_reference.py (git-ignored, travels to the GPU box):
# used by bench.py's reference arm when /root/reference itself is not mounted
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root():
    """Where the reference's `src` package can be imported from: the mounted tree, else the installed copy, else None."""
    for root in (REFERENCE_ROOT, INSTALLED_ROOT):
        if os.path.isfile(os.path.join(root, "src", "flow_matching", "models.py")):
            return root
    return None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "flow_matching"))


def load_reference():
    """Returns (ConditionalFlowMatchingWithHifiGan, ConditionalFlowMatchingWithHifiGanConfig, ConditionalFlowMatchingConfig)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT + " nor installed under " + INSTALLED_ROOT)
    import transformers  # noqa: F401  (must precede the librosa stub)
    from transformers import FastSpeech2ConformerHifiGan  # noqa: F401

    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        filt = types.ModuleType("librosa.filters")
        filt.mel = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("librosa stub: not on the hot path"))
        lib.filters = filt
        sys.modules["librosa"] = lib
        sys.modules["librosa.filters"] = filt
    if "einx" not in sys.modules:
        einx = types.ModuleType("einx")

        def multiply(pattern, a, b):
            assert pattern.replace(" ", "") == "i,j->ij", pattern
            return a[:, None] * b[None, :]

        einx.multiply = multiply
        sys.modules["einx"] = einx
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    from src.flow_matching.configs import ConditionalFlowMatchingConfig, ConditionalFlowMatchingWithHifiGanConfig
    from src.flow_matching.models import ConditionalFlowMatchingWithHifiGan

    return ConditionalFlowMatchingWithHifiGan, ConditionalFlowMatchingWithHifiGanConfig, ConditionalFlowMatchingConfig


def build_reference_model(state_dict, predict_duration: bool = False):
    """Reference model in eval mode carrying ``state_dict`` (strict)."""
    cls, cfg_cls, cfm_cfg_cls = load_reference()
    cfg = cfg_cls(
        model_config=cfm_cfg_cls(predict_duration=predict_duration).to_dict(),
        vocoder_config=dict(upsample_rates=[5, 4, 4, 2, 2], upsample_kernel_sizes=[10, 9, 8, 4, 4], normalize_before=False),
    )
    model = cls(cfg).eval()
    model.load_state_dict(state_dict, strict=True)
    return model
