"""speech_resynth_b200 -- B200-native unit-to-speech resynthesis (drop-in for speech_resynth's decoder API).

    from speech_resynth_b200 import ConditionalFlowMatchingWithHifiGan
    decoder = ConditionalFlowMatchingWithHifiGan.from_pretrained(path).cuda()
    wavs = decoder(units + 1)            # list of (1, 320 * len + 80) float tensors, 16 kHz
"""
from .configs import (ConditionalFlowMatchingConfig, ConditionalFlowMatchingWithHifiGanConfig, reference_config)
from .models import ConditionalFlowMatchingModel, ConditionalFlowMatchingWithHifiGan, HifiGanVocoder

__all__ = [
    "ConditionalFlowMatchingConfig",
    "ConditionalFlowMatchingWithHifiGanConfig",
    "ConditionalFlowMatchingModel",
    "ConditionalFlowMatchingWithHifiGan",
    "HifiGanVocoder",
    "reference_config",
]
