"""Drop-in model classes: same constructors, config, state-dict keys and call signatures as the reference
(src/flow_matching/models.py:41-256; vocoder = transformers FastSpeech2ConformerHifiGan, HF:1376-1491), with the
arithmetic executed by the sm_100a kernels of libsrb.so.

The ``nn.Module`` tree below exists to *own the parameters under the reference's names* so that
``from_pretrained`` / ``save_pretrained`` / ``state_dict`` / ``.cuda()`` behave identically and the public
checkpoint loads unchanged.  None of these modules has a PyTorch ``forward``: inference goes through
``speech_resynth_b200.engine`` and raises if the CUDA library or a Blackwell GPU is missing (no CPU fallback).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import nn
from transformers import FastSpeech2ConformerHifiGanConfig, PreTrainedModel
from transformers import initialization as hf_init

from . import engine as _engine
from .configs import ConditionalFlowMatchingConfig, ConditionalFlowMatchingWithHifiGanConfig


class _ParamsOnly(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} only stores parameters; inference runs in the CUDA engine "
            "(speech_resynth_b200 has no PyTorch/CPU fallback path)"
        )


# ---- parameter containers named like the reference modules --------------------------------------------------
class RandomFourierEmbed(_ParamsOnly):  # fourier_embed.py:31-35
    def __init__(self, hidden_size: int):
        super().__init__()
        self.register_buffer("weights", torch.randn(hidden_size // 2))


class RotaryEmbedding(_ParamsOnly):  # transformer.py:40-48
    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim)))


class ConvPositionEmbed(_ParamsOnly):  # transformer.py:76-82
    def __init__(self, hidden_size: int, kernel_size: int, groups: int):
        super().__init__()
        self.dw_conv1d = nn.Sequential(
            nn.Conv1d(hidden_size, hidden_size, kernel_size, groups=groups, padding=kernel_size // 2), nn.GELU()
        )


class AdaptiveRMSNorm(_ParamsOnly):  # norm.py:30-35
    def __init__(self, hidden_size: int):
        super().__init__()
        self.to_weight = nn.Linear(hidden_size, hidden_size, bias=False)
        nn.init.zeros_(self.to_weight.weight)


class Attention(_ParamsOnly):  # transformer.py:99-106
    def __init__(self, hidden_size: int):
        super().__init__()
        self.to_qkv = nn.Linear(hidden_size, hidden_size * 3, bias=False)
        self.to_out = nn.Linear(hidden_size, hidden_size, bias=False)


class FeedForward(_ParamsOnly):  # fastspeech/modules.py:39-47
    def __init__(self, hidden_size: int, intermediate_size: int, kernel_size: int = 3):
        super().__init__()
        pad = (kernel_size - 1) // 2
        self.conv1 = nn.Conv1d(hidden_size, intermediate_size * 2, kernel_size, padding=pad)
        self.conv2 = nn.Conv1d(intermediate_size, hidden_size, kernel_size, padding=pad)


class ConditionalFlowMatchingDurationPredictor(_ParamsOnly):  # fastspeech/modules.py:76-86
    def __init__(self, dim_cond_emb: int):
        super().__init__()
        self.conv = nn.Conv1d(dim_cond_emb, 1, kernel_size=3, padding=1)


class Transformer(_ParamsOnly):  # transformer.py:133-170
    def __init__(self, hidden_size: int, depth: int, heads: int, intermediate_size: int):
        super().__init__()
        self.rotary_emb = RotaryEmbedding(hidden_size // heads)
        self.layers = nn.ModuleList(
            nn.ModuleList([None, AdaptiveRMSNorm(hidden_size), Attention(hidden_size), AdaptiveRMSNorm(hidden_size),
                           FeedForward(hidden_size, intermediate_size)])
            for _ in range(depth)
        )
        self.final_norm = nn.RMSNorm(hidden_size)


def _reset_like_reference(module: nn.Module) -> None:
    """The reference never calls post_init(), so its effective initialisation is torch's constructor default plus
    the zero-initialised adaptive-norm projection (norm.py:35).  Reproduce that under transformers' init hook."""
    # hf_init.* and (under transformers' guard) torch.nn.init.* skip tensors already loaded from a checkpoint
    if isinstance(module, AdaptiveRMSNorm):
        hf_init.zeros_(module.to_weight.weight)
        module.to_weight._is_hf_initialized = True  # keep the zero init when the walk reaches the inner Linear
    elif isinstance(module, RandomFourierEmbed):
        hf_init.normal_(module.weights, mean=0.0, std=1.0)
    elif isinstance(module, RotaryEmbedding):
        dim = module.inv_freq.shape[0] * 2
        hf_init.copy_(module.inv_freq, 1.0 / (10000.0 ** (torch.arange(0, dim, 2).float() / dim)))
    elif isinstance(module, (nn.Linear, nn.Conv1d, nn.ConvTranspose1d, nn.Embedding, nn.RMSNorm)):
        module.reset_parameters()


def _check_right_padded(counts, extents) -> None:
    """counts = non-pad ids per row, extents = index of the last non-pad id + 1.  The reference masks per position
    (input_ids.ne(0), models.py:152); the kernels take a prefix length per utterance.  Both agree for right-padded rows,
    which is what every caller of the reference passes (synthesize.py:39-42, data.py:132,200)."""
    if not bool((counts == extents).all()):
        bad = [i for i, (c, e) in enumerate(zip(counts.tolist(), extents.tolist())) if c != e]
        raise ValueError(f"input_ids rows {bad[:8]} contain pad ids (0) before their last unit: rows must be right-padded")
    if int(counts.min()) <= 0:
        raise ValueError("every row of input_ids needs at least one unit (the reference yields NaN for an all-pad row)")


def _param_signature(module: nn.Module):
    """Changes whenever a parameter / buffer is replaced or written in place (load_state_dict, load_adapter, manual
    edits): the packed bf16 weights and captured graphs are rebuilt when it does."""
    try:
        return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))
    except RuntimeError:       # inference tensors carry no version counter
        return tuple(t.data_ptr() for t in list(module.parameters()) + list(module.buffers()))


def _fresh_state(module: nn.Module):
    return {k: v.detach() for k, v in module.state_dict().items()}


def _default_precision() -> str:
    p = os.environ.get("SRB_PRECISION", "bf16")
    if p not in ("bf16", "tight"):
        raise ValueError(f"SRB_PRECISION must be 'bf16' or 'tight', got {p!r}")
    return p


class _PrecisionMixin:
    """precision = "bf16" (default: the product kernels, bf16 tensor-core operands, parity ~1e-3) or "tight" (the same
    kernels over split bf16 operands hi + lo with fp32 accumulation and fp32 CUDA-core attention: parity ~1e-5 against the
    reference's fp32 path, at several times the cost -- BASELINE.json north_star: "tight in fp32/TF32, looser in bf16")."""

    _precision: Optional[str] = None

    @property
    def precision(self) -> str:
        return self._precision or _default_precision()

    def set_precision(self, precision: str) -> "_PrecisionMixin":
        if precision not in ("bf16", "tight"):
            raise ValueError(f"precision must be 'bf16' or 'tight', got {precision!r}")
        if precision != self.precision:
            self._precision = precision
            self.refresh()
        for child in (getattr(self, "model", None), getattr(self, "vocoder", None)):
            if isinstance(child, _PrecisionMixin):
                child.set_precision(precision)
        return self


class ConditionalFlowMatchingModel(_PrecisionMixin, PreTrainedModel):
    """Parameter layout of the reference model (models.py:41-71); ``sample`` runs on the GPU kernels."""

    config_class = ConditionalFlowMatchingConfig
    base_model_prefix = "model"

    def __init__(self, config: ConditionalFlowMatchingConfig, embedding: Optional[nn.Embedding] = None):
        super().__init__(config)
        self.config = config
        if config.use_unet_skip_connection:
            raise NotImplementedError("the U-Net skip variant is not built (no shipped config uses it)")
        if (config.hidden_size, config.heads, config.dim_in, config.intermediate_size, config.conv_pos_embed_kernel_size) != (
                256, 2, 80, 896, 31) or config.conv_pos_embed_groups != config.hidden_size:
            raise NotImplementedError("kernels are specialised to the mhubert-expresso-2000 architecture")
        h = config.hidden_size
        self.time_cond_mlp = nn.Sequential(RandomFourierEmbed(h), nn.Linear(h + 1, h), nn.SiLU())
        self.to_cond_emb = (
            nn.Embedding(config.vocab_size + 1, config.dim_cond_emb, padding_idx=0) if embedding is None else embedding
        )
        self.to_embed = nn.Linear(config.dim_in + config.dim_cond_emb, h)
        self.conv_embed = ConvPositionEmbed(h, config.conv_pos_embed_kernel_size, config.conv_pos_embed_groups)
        self.transformer = Transformer(h, config.depth, config.heads, config.intermediate_size)
        self.to_pred = nn.Linear(h, config.dim_in, bias=False)
        # second shipped config (mhubert-expresso-2000-duration-prediction.yaml): de-duplicated units + predicted durations
        self.duration_predictor = ConditionalFlowMatchingDurationPredictor(config.dim_cond_emb) if config.predict_duration else None
        self._sampler = None
        self._engine = None
        self._sig = None
        self.post_init()

    def _init_weights(self, module):
        with torch.no_grad():
            _reset_like_reference(module)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, *args, **kwargs):
        raise NotImplementedError("the training loss (models.py:77-130) is outside the accelerated inference path")

    # -- engine plumbing --------------------------------------------------------------------------------------
    def refresh(self) -> None:
        """Drop packed weights / graphs (call after changing parameters or moving devices)."""
        self._sampler = None
        self._engine = None

    def sampler(self) -> "_engine.CFMSampler":
        dev = self.device
        sig = _param_signature(self)
        if self._sampler is None or self._sampler.device != dev or self._sig != sig:
            sd = {"model." + k: v for k, v in _fresh_state(self).items()}
            self._sampler = _engine.build_sampler(sd, dev, depth=self.config.depth, mean=self.config.mean, std=self.config.std,
                                                  tight=self.precision == "tight")
            self._sig = sig
            self._engine = None
        return self._sampler

    def _own_engine(self) -> "_engine.ResynthEngine":
        sampler = self.sampler()
        if self._engine is None:
            self._engine = _engine.ResynthEngine(sampler, None)
        return self._engine

    @torch.inference_mode()
    def sample(self, input_ids: torch.LongTensor, dt: float = 0.1, truncation_value: Optional[float] = None) -> torch.FloatTensor:
        """Same contract as the reference (models.py:132-189): (B, N) unit ids (0 = pad) -> (B, N, 80) log-mel; with
        config.predict_duration the units are first expanded by their predicted durations (models.py:157-164) and N
        becomes the longest expanded length."""
        ids = input_ids.to(self.device)
        if self.config.predict_duration:
            ids, _ = self.sampler().regulate(ids)
        return self._own_engine().sample(ids, dt, truncation_value)

    @torch.inference_mode()
    def predict_durations(self, input_ids: torch.LongTensor) -> torch.LongTensor:
        """duration_predictor(to_cond_emb(ids)).masked_fill(~mask, 0) (models.py:158-159): (B, N) frames per unit."""
        _, dur = self.sampler().regulate(input_ids.to(self.device))
        return dur.long()

    @torch.inference_mode()
    def embed_units(self, input_ids: torch.LongTensor) -> torch.FloatTensor:
        """to_cond_emb(input_ids) (models.py:154) through the gather kernel: (B, N) -> (B, N, 768), bit exact."""
        from . import _native as nat

        sampler = self.sampler()
        ids = input_ids.to(self.device).contiguous()
        table = sampler.w.emb_table
        out = torch.empty(*ids.shape, table.shape[1], dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            nat.call("srb_embed_gather", nat.ptr(table), nat.ptr(ids), nat.ptr(out), ids.numel(), table.shape[0], table.shape[1])
        return out


class _HifiGanResidualBlock(_ParamsOnly):  # HF:1308-1339
    def __init__(self, channels: int, kernel_size: int, dilation):
        super().__init__()
        self.convs1 = nn.ModuleList(
            nn.Conv1d(channels, channels, kernel_size, dilation=d, padding=(kernel_size * d - d) // 2) for d in dilation)
        self.convs2 = nn.ModuleList(
            nn.Conv1d(channels, channels, kernel_size, padding=(kernel_size - 1) // 2) for _ in dilation)


class HifiGanVocoder(_PrecisionMixin, PreTrainedModel):
    """Parameter layout of transformers' FastSpeech2ConformerHifiGan (HF:1376-1416); call = mel -> waveform."""

    config_class = FastSpeech2ConformerHifiGanConfig
    main_input_name = "spectrogram"

    def __init__(self, config: FastSpeech2ConformerHifiGanConfig):
        super().__init__(config)
        if (list(config.upsample_rates), list(config.upsample_kernel_sizes), list(config.resblock_kernel_sizes),
                config.upsample_initial_channel, config.model_in_dim, config.normalize_before) != (
                [5, 4, 4, 2, 2], [10, 9, 8, 4, 4], [3, 7, 11], 512, 80, False) or any(
                list(d) != [1, 3, 5] for d in config.resblock_dilation_sizes):
            raise NotImplementedError("kernels are specialised to the reference's HiFi-GAN hyper-parameters "
                                      "(src/hifigan/train.py:36-42)")
        c0 = config.upsample_initial_channel
        self.conv_pre = nn.Conv1d(config.model_in_dim, c0, kernel_size=7, padding=3)
        self.upsampler = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        c = c0
        for rate, k in zip(config.upsample_rates, config.upsample_kernel_sizes):
            self.upsampler.append(nn.ConvTranspose1d(c, c // 2, k, stride=rate, padding=(k - rate) // 2))
            c //= 2
            for rk, dil in zip(config.resblock_kernel_sizes, config.resblock_dilation_sizes):
                self.resblocks.append(_HifiGanResidualBlock(c, rk, dil))
        self.conv_post = nn.Conv1d(c, 1, kernel_size=7, padding=3)
        self.register_buffer("mean", torch.zeros(config.model_in_dim))
        self.register_buffer("scale", torch.ones(config.model_in_dim))
        self._generator = None
        self._engine = None
        self._sig = None
        self.post_init()

    def _init_weights(self, module):
        with torch.no_grad():
            _reset_like_reference(module)
            if isinstance(module, HifiGanVocoder):
                hf_init.zeros_(module.mean)
                hf_init.ones_(module.scale)

    @property
    def device(self):
        return next(self.parameters()).device

    def refresh(self) -> None:
        self._generator = None
        self._engine = None

    def generator(self) -> "_engine.HifiGanGenerator":
        dev = self.device
        sig = _param_signature(self)
        if self._generator is None or self._generator.device != dev or self._sig != sig:
            sd = {"vocoder." + k: v for k, v in _fresh_state(self).items()}
            self._generator = _engine.build_vocoder(sd, dev, slope=self.config.leaky_relu_slope, tight=self.precision == "tight")
            self._sig = sig
            self._engine = None
        return self._generator

    @torch.inference_mode()
    def forward(self, spectrogram: torch.FloatTensor, **kwargs) -> torch.FloatTensor:
        """(B, T, 80) or (T, 80) log-mel -> (B, 320 T + 80) or (320 T + 80,) waveform (HF:1451-1491)."""
        gen = self.generator()
        if self._engine is None:
            self._engine = _engine.ResynthEngine(None, gen)
        return self._engine.vocode(spectrogram.to(self.device))


class ConditionalFlowMatchingWithHifiGan(_PrecisionMixin, PreTrainedModel):
    """units -> list of waveforms; the reference's public entry point (models.py:192-256)."""

    config_class = ConditionalFlowMatchingWithHifiGanConfig

    def __init__(self, config: ConditionalFlowMatchingWithHifiGanConfig):
        super().__init__(config)
        self.model = ConditionalFlowMatchingModel(config.model_config)
        self.vocoder = HifiGanVocoder(config.vocoder_config)
        self._engine = None
        self.post_init()

    def _init_weights(self, module):
        with torch.no_grad():
            _reset_like_reference(module)
            if isinstance(module, HifiGanVocoder):
                hf_init.zeros_(module.mean)
                hf_init.ones_(module.scale)

    @classmethod
    def load_pretrained(cls, model_path, vocoder_path) -> "ConditionalFlowMatchingWithHifiGan":
        """Assemble from two separately saved checkpoints (models.py:200-209)."""
        model_config = ConditionalFlowMatchingConfig.from_pretrained(model_path)
        vocoder_config = FastSpeech2ConformerHifiGanConfig.from_pretrained(vocoder_path)
        config = ConditionalFlowMatchingWithHifiGanConfig(model_config.to_dict(), vocoder_config.to_dict())
        model = cls(config)
        model.model = ConditionalFlowMatchingModel.from_pretrained(model_path)
        model.vocoder = HifiGanVocoder.from_pretrained(vocoder_path)
        return model

    @property
    def device(self):
        return next(self.parameters()).device

    def refresh(self) -> None:
        self.model.refresh()
        self.vocoder.refresh()
        self._engine = None

    def engine(self) -> "_engine.ResynthEngine":
        sampler, gen = self.model.sampler(), self.vocoder.generator()
        if self._engine is None or self._engine.sampler is not sampler or self._engine.vocoder is not gen:
            self._engine = _engine.ResynthEngine(sampler, gen)
        return self._engine

    def _get_waveform_lengths(self, spectrogram_lengths):
        """(L - 1) * s - 2 * ((k - s) // 2) + k over the five up-samplers (models.py:211-221) = 320 L + 80."""
        vc = self.config.vocoder_config
        for k, s in zip(vc.upsample_kernel_sizes, vc.upsample_rates):
            spectrogram_lengths = (spectrogram_lengths - 1) * s - 2 * ((k - s) // 2) + k
        return spectrogram_lengths

    @torch.inference_mode()
    def resynthesize_flat(self, input_ids: torch.LongTensor, dt: float = 0.1, truncation_value: Optional[float] = None,
                          out: Optional[torch.Tensor] = None):
        """The work of `forward`, returned as (flat, wav_lengths): `flat` is a fresh 1-D CUDA tensor holding the cropped
        waveforms back to back (utterance i occupies wav_lengths[i] = 320 len_i + 80 samples).  Batch drivers use this form
        (one read-back per bucket instead of one per utterance); with `out` (1-D float32 CUDA tensor of at least that
        many samples) the waveforms are written there instead.

        `input_ids` may live on the host: its valid-frame counts are then taken there and the call never waits for the
        GPU.  For a CUDA tensor they are read back through pinned memory, enqueued ahead of the kernels (the reference
        syncs once per utterance, models.py:252-256).  Rows must be right-padded (the reference's input convention,
        synthesize.py:39-42): the kernels mask by length where the reference masks by position."""
        dev = self.device
        eng = self.engine()
        if self.config.model_config.predict_duration:
            input_ids, _ = self.model.sampler().regulate(input_ids.to(dev))     # models.py:157-164
        if not input_ids.is_cuda:
            nz = input_ids.ne(0)
            counts = nz.sum(dim=1)
            n = input_ids.shape[1]
            extents = (nz * torch.arange(1, n + 1)).amax(dim=1) if n > 0 else counts
            _check_right_padded(counts, extents)
            ids = input_ids.to(dev, non_blocking=True)
            wav_lengths = self._get_waveform_lengths(counts.to(torch.int64)).tolist()
            flat = eng.resynthesize_ragged(ids, dt, truncation_value, sum(wav_lengths), out=out)
            return flat, wav_lengths
        ids = input_ids.to(dev)
        host, ready = eng.lengths_async(ids)
        ready.synchronize()     # waits for the length kernel only: it is queued ahead of this call's work
        _check_right_padded(host[0], host[1])
        wav_lengths = self._get_waveform_lengths(host[0].to(torch.int64)).tolist()
        flat = eng.resynthesize_ragged(ids, dt, truncation_value, sum(wav_lengths), out=out)
        return flat, wav_lengths

    @torch.inference_mode()
    def forward(self, input_ids: torch.LongTensor, dt: float = 0.1, truncation_value: Optional[float] = None
                ) -> List[torch.FloatTensor]:
        """Same contract as the reference (models.py:223-256): list of B tensors (1, 320 * len_i + 80).

        The valid-frame counts come from ``input_ids`` (the reference re-derives them by scanning the mel for the
        exact pad value, models.py:245-247 -- identical result, see tests); the per-utterance crop (models.py:252-256)
        happens in the last kernel's store, so the returned tensors are views of one fresh buffer.
        """
        flat, wav_lengths = self.resynthesize_flat(input_ids, dt, truncation_value)
        outs, off = [], 0
        for n in wav_lengths:
            outs.append(flat[off: off + n].unsqueeze(0))
            off += n
        return outs
