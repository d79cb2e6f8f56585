"""On-disk configuration of the drop-in model (``config.json``).

Field names, defaults and nesting follow the reference so that checkpoints written by either side load
in the other (reference: src/flow_matching/configs.py:7-24 for the flow-matching fields, :44-61 for the
nested ``model_config`` / ``vocoder_config`` pair).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

from transformers import FastSpeech2ConformerHifiGanConfig, PretrainedConfig

# (field, default) pairs of the flow-matching velocity model -- reference configs.py:9-23
_CFM_FIELDS = (
    ("vocab_size", 2000),
    ("dim_in", 80),
    ("dim_cond_emb", 768),
    ("hidden_size", 256),
    ("depth", 4),
    ("heads", 2),
    ("intermediate_size", 896),
    ("ff_dropout", 0.0),
    ("use_unet_skip_connection", False),
    ("conv_pos_embed_kernel_size", 31),
    ("conv_pos_embed_groups", 256),
    ("attn_dropout", 0.0),
    ("mean", -5.8843),
    ("std", 2.2615),
    ("predict_duration", False),
)

# vocoder hyper-parameters the reference trains with (src/hifigan/train.py:36-42)
REFERENCE_VOCODER_KWARGS = dict(
    upsample_rates=[5, 4, 4, 2, 2],
    upsample_kernel_sizes=[10, 9, 8, 4, 4],
    normalize_before=False,
)


class ConditionalFlowMatchingConfig(PretrainedConfig):
    def __init__(self, *args: Any, **kwargs: Any):
        """Positional arguments follow the reference's parameter order (configs.py:9-23: vocab_size, dim_in, ...)."""
        if len(args) > len(_CFM_FIELDS):
            raise TypeError(f"ConditionalFlowMatchingConfig takes at most {len(_CFM_FIELDS)} positional arguments ({len(args)} given)")
        for (name, _), value in zip(_CFM_FIELDS, args):
            if name in kwargs:
                raise TypeError(f"ConditionalFlowMatchingConfig got multiple values for argument {name!r}")
            kwargs[name] = value
        for name, default in _CFM_FIELDS:
            setattr(self, name, kwargs.pop(name, default))
        super().__init__(**kwargs)


class ConditionalFlowMatchingWithHifiGanConfig(PretrainedConfig):
    sub_configs = {"model_config": ConditionalFlowMatchingConfig, "vocoder_config": FastSpeech2ConformerHifiGanConfig}

    def __init__(self, model_config: Optional[Dict] = None, vocoder_config: Optional[Dict] = None, **kwargs: Any):
        def as_dict(cfg):
            if cfg is None:
                return {}
            return cfg.to_dict() if isinstance(cfg, PretrainedConfig) else dict(cfg)

        self.model_config = ConditionalFlowMatchingConfig(**as_dict(model_config))
        self.vocoder_config = FastSpeech2ConformerHifiGanConfig(**as_dict(vocoder_config))
        super().__init__(**kwargs)


def reference_config() -> ConditionalFlowMatchingWithHifiGanConfig:
    """The mhubert-expresso-2000 configuration (configs/resynth/mhubert-expresso-2000.yaml:48-64,80-81)."""
    return ConditionalFlowMatchingWithHifiGanConfig(
        model_config=ConditionalFlowMatchingConfig().to_dict(), vocoder_config=dict(REFERENCE_VOCODER_KWARGS)
    )
