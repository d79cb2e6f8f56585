"""CPU oracle for the unit-to-speech hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain functional restatement (torch CPU tensor ops, any float
dtype; run it in float64 for a tight anchor) of the arithmetic performed by the
reference path

    ConditionalFlowMatchingWithHifiGan.forward   /root/reference/src/flow_matching/models.py:223-256
      -> ConditionalFlowMatchingModel.sample      models.py:132-189
      -> transformers FastSpeech2ConformerHifiGan (third party, pinned 4.49.0 by
         requirements/requirements.txt:15; installed 5.5.0; "HF:" below =
         transformers/models/fastspeech2_conformer/modeling_fastspeech2_conformer.py)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  The product package
(``speech_resynth_b200``) never does: it fails loudly when its CUDA library is
missing.

Parity pin: the reference ships no tests, golden vectors or checkpoints
(SURVEY.md section 4), so this restatement is pinned against the *live reference
classes* imported from /root/reference in the build container
(``oracle/make_golden.py``; results committed under ``tests/golden/``) and the
CPU test-suite re-checks the oracle against those committed vectors.

Every function takes the reference's own ``state_dict`` (same key names) so the
same weights drive the reference, this oracle and the CUDA path.
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, List, Optional, Tuple, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

warnings.filterwarnings("ignore", message="TF32 acceleration on top of oneDNN")

# vocoder hyper-parameters fixed by the reference (src/hifigan/train.py:36-42 and HF config defaults)
UPSAMPLE_RATES = (5, 4, 4, 2, 2)
UPSAMPLE_KERNELS = (10, 9, 8, 4, 4)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)
LRELU_SLOPE = 0.1


def pad_value() -> float:
    """log(clamp(0, 1e-5)) evaluated in float32 -- hifigan/data.py:9-10, used at models.py:187,245."""
    return float(torch.log(torch.clamp(torch.tensor(0.0, dtype=torch.float32), min=1e-5)))


# --------------------------------------------------------------------------------------
# conditional flow matching model
# --------------------------------------------------------------------------------------
def embed_gather(table: Tensor, ids: Tensor) -> Tensor:
    """nn.Embedding(vocab+1, 768, padding_idx=0) forward -- models.py:50-52,154.  Pure row gather."""
    return table.index_select(0, ids.reshape(-1)).reshape(*ids.shape, table.shape[1])


def ode_times(dt: float) -> Tensor:
    """Step times are the float32 tensor torch.arange(0, 1, dt) -- models.py:172."""
    return torch.arange(0, 1, dt)


def time_embedding(sd: Dict[str, Tensor], t: Tensor) -> Tensor:
    """time_cond_mlp = RandomFourierEmbed -> Linear(257,256) -> SiLU -- models.py:47-49,179; fourier_embed.py:37-40.

    ``t`` is a 0-d/1-element tensor; the reference expands it over the batch, every row is identical,
    so one row (256,) is returned.  freqs = ((t*w)*2)*pi in that association order.
    """
    w = sd["model.time_cond_mlp.0.weights"]
    t = t.reshape(1).to(device=w.device, dtype=w.dtype)
    freqs = (t[:, None] * w[None, :]) * 2 * math.pi
    four = torch.cat([t[:, None], freqs.sin(), freqs.cos()], dim=-1)  # pack((x, sin, cos), "b *")
    h = F.linear(four, sd["model.time_cond_mlp.1.weight"], sd["model.time_cond_mlp.1.bias"])
    return F.silu(h)[0]


def adaptive_rmsnorm(x: Tensor, cond: Tensor, to_weight: Tensor) -> Tensor:
    """AdaptiveRMSNorm.forward -- norm.py:37-43: F.normalize(x) * sqrt(H) * (W_gamma c + 1)."""
    gamma = F.linear(cond, to_weight)
    denom = x.norm(dim=-1, keepdim=True).clamp_min(1e-12)  # F.normalize eps
    return x / denom * (x.shape[-1] ** 0.5) * (gamma + 1.0)


def rotary_table(inv_freq: Tensor, n: int) -> Tensor:
    """RotaryEmbedding.forward -- transformer.py:55-63: pos*inv_freq, cat(freqs, freqs)."""
    t = torch.arange(n, device=inv_freq.device).to(inv_freq.dtype)
    freqs = t[:, None] * inv_freq[None, :]
    return torch.cat([freqs, freqs], dim=-1)


def apply_rotary(pos: Tensor, t: Tensor) -> Tensor:
    """rotate_half / apply_rotary_pos_emb -- transformer.py:66-73."""
    t1, t2 = t.chunk(2, dim=-1)
    return t * pos.cos() + torch.cat([-t2, t1], dim=-1) * pos.sin()


def attention(sd: Dict[str, Tensor], prefix: str, x: Tensor, mask: Tensor, rot: Tensor, heads: int) -> Tensor:
    """Attention.forward -- transformer.py:108-130 (key-padding mask, scale 1/sqrt(d_head), no dropout)."""
    b, n, hdim = x.shape
    d = hdim // heads
    qkv = F.linear(x, sd[prefix + "to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (z.reshape(b, n, heads, d).permute(0, 2, 1, 3) for z in (q, k, v))
    q, k = apply_rotary(rot, q), apply_rotary(rot, k)
    s = torch.einsum("bhid,bhjd->bhij", q, k) / math.sqrt(d)
    s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", p, v)
    o = o.permute(0, 2, 1, 3).reshape(b, n, hdim)
    return F.linear(o, sd[prefix + "to_out.weight"])


def feed_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, mask: Tensor) -> Tensor:
    """FeedForward.forward + SIGLU -- fastspeech/modules.py:27-30,49-73.

    Pads are zeroed before each conv; SIGLU splits the *channel* dim: first half = value, second = gate.
    """
    h = x.transpose(1, 2)
    m = mask[:, None, :]
    h = h.masked_fill(~m, 0.0)
    h = F.conv1d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    val, gate = h.chunk(2, dim=1)
    h = F.silu(gate) * val
    h = h.masked_fill(~m, 0.0)
    h = F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], padding=1)
    return h.transpose(1, 2)


def conv_pos_embed(sd: Dict[str, Tensor], x: Tensor, mask: Tensor) -> Tensor:
    """ConvPositionEmbed.forward -- transformer.py:84-96: mask -> depthwise k=31 -> exact GELU -> mask."""
    w = sd["model.conv_embed.dw_conv1d.0.weight"]
    bias = sd["model.conv_embed.dw_conv1d.0.bias"]
    h = x.masked_fill(~mask[..., None], 0.0).transpose(1, 2)
    h = F.conv1d(h, w, bias, padding=w.shape[-1] // 2, groups=w.shape[0])
    h = F.gelu(h).transpose(1, 2)
    return h.masked_fill(~mask[..., None], 0.0)


def rms_norm(x: Tensor, weight: Tensor) -> Tensor:
    """nn.RMSNorm(H) with eps=None -> torch.finfo(x.dtype).eps -- transformer.py:170,208."""
    eps = torch.finfo(x.dtype).eps
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps) * weight


def transformer(sd: Dict[str, Tensor], x: Tensor, mask: Tensor, cond: Tensor, depth: int, heads: int) -> Tensor:
    """Transformer.forward -- transformer.py:176-208 (U-Net skips disabled in the target config)."""
    n = x.shape[1]
    d_head = x.shape[-1] // heads
    rot = rotary_table(sd["model.transformer.rotary_emb.inv_freq"], n).to(x.dtype)
    assert rot.shape[-1] == d_head
    for i in range(depth):
        p = f"model.transformer.layers.{i}."
        a_in = adaptive_rmsnorm(x, cond, sd[p + "1.to_weight.weight"])
        x = attention(sd, p + "2.", a_in, mask, rot, heads) + x
        f_in = adaptive_rmsnorm(x, cond, sd[p + "3.to_weight.weight"])
        x = feed_forward(sd, p + "4.", f_in, mask) + x
    return rms_norm(x, sd["model.transformer.final_norm.weight"])


def velocity(sd: Dict[str, Tensor], xt: Tensor, cond_emb: Tensor, mask: Tensor, t: Tensor,
             depth: int = 4, heads: int = 2) -> Tensor:
    """One evaluation of the velocity field -- the loop body models.py:175-183."""
    x = torch.cat([xt, cond_emb], dim=-1)
    x = F.linear(x, sd["model.to_embed.weight"], sd["model.to_embed.bias"])
    x = conv_pos_embed(sd, x, mask) + x
    c = time_embedding(sd, t).to(x.dtype)
    x = transformer(sd, x, mask, c, depth, heads)
    return F.linear(x, sd["model.to_pred.weight"])


def sample(sd: Dict[str, Tensor], ids: Tensor, x0: Tensor, dt: float = 0.1,
           truncation_value: Optional[float] = None, mean: float = -5.8843, std: float = 2.2615,
           depth: int = 4, heads: int = 2) -> Tensor:
    """ConditionalFlowMatchingModel.sample -- models.py:132-189 with the prior sample ``x0`` injected.

    (The reference draws x0 = torch.randn(B, N, 80) itself at :168; callers reproduce it by seeding.)
    """
    dtype = sd["model.to_embed.weight"].dtype
    mask = ids.ne(0)
    cond_emb = embed_gather(sd["model.to_cond_emb.weight"], ids)
    xt = x0.to(dtype)
    if truncation_value is not None:
        xt = xt.clamp(-truncation_value, truncation_value)
    for t in ode_times(dt):
        vt = velocity(sd, xt, cond_emb, mask, t, depth, heads)
        xt = xt + vt * dt
    x1 = xt * std + mean
    x1[~mask] = pad_value()
    return x1


# --------------------------------------------------------------------------------------
# HiFi-GAN generator (third-party transformers code restated; HF:1308-1367, 1451-1491)
# --------------------------------------------------------------------------------------
def duration_predict(sd: Dict[str, Tensor], ids: Tensor) -> Tensor:
    """ConditionalFlowMatchingDurationPredictor.forward in eval mode + the pad masking of sample()
    -- fastspeech/modules.py:87-107, models.py:158-159.  (B, N) ids -> (B, N) int64 frames per unit."""
    hs = embed_gather(sd["model.to_cond_emb.weight"], ids)
    x = F.conv1d(hs.transpose(1, 2), sd["model.duration_predictor.conv.weight"], sd["model.duration_predictor.conv.bias"],
                 padding=1).squeeze(1)
    d = torch.clamp(torch.round(x.exp() - 1.0), min=0).long()
    return d.masked_fill(~ids.ne(0), 0)


def length_regulate_ids(ids: Tensor, durations: Tensor) -> Tuple[Tensor, Tensor]:
    """transformers length_regulator (HF:88-134) applied to the unit ids instead of their embeddings (the embedding of
    the pad id is the zero row, so gathering the expanded ids gives exactly the regulator's zero-padded output), and
    the mask update of models.py:162-164.  Returns (expanded ids (B, max_len), lengths (B,))."""
    durations = durations.clone()
    if int(durations.sum()) == 0:
        durations[durations.sum(dim=1).eq(0)] = 1      # HF:113-114 (in place in the reference: lengths see it too)
    lengths = durations.sum(dim=1)
    out = torch.zeros(ids.shape[0], int(lengths.max()), dtype=ids.dtype)
    for b in range(ids.shape[0]):
        rep = torch.repeat_interleave(ids[b], durations[b])
        out[b, : rep.numel()] = rep
    return out, lengths


def conv_transpose1d_polyphase(x: Tensor, w: Tensor, bias: Tensor, stride: int, padding: int) -> Tensor:
    """ConvTranspose1d written in the polyphase (gather) form the CUDA kernels use.

    out[o] = sum over taps j with (o + p - j) % s == 0 of  W[:, :, j]^T x[(o + p - j) / s]
    (skipping out-of-range inputs).  Equivalent to F.conv_transpose1d (HF:1392-1402); kept explicit
    so the tests can pin the phase/tap bookkeeping on small cases.
    x: (B, C_in, L), w: (C_in, C_out, k).
    """
    b, cin, lin = x.shape
    _, cout, k = w.shape
    lout = (lin - 1) * stride - 2 * padding + k
    out = bias.reshape(1, cout, 1).expand(b, cout, lout).clone()
    for r in range(stride):
        j0 = (r + padding) % stride
        c_r = (r + padding) // stride
        nq = (lout - r + stride - 1) // stride
        if nq <= 0:
            continue
        q = torch.arange(nq)
        for m_, j in enumerate(range(j0, k, stride)):
            src = q + c_r - m_
            ok = (src >= 0) & (src < lin)
            contrib = torch.einsum("bcl,cd->bdl", x[:, :, src.clamp(0, lin - 1)], w[:, :, j]) * ok.to(x.dtype)
            out[:, :, r::stride] += contrib
    return out


def conv_transpose1d(x: Tensor, w: Tensor, bias: Tensor, stride: int, padding: int) -> Tensor:
    """nn.ConvTranspose1d forward (HF:1392-1402, 1473).

    torch 2.11's oneDNN float32 deconvolution was observed to return wrong values (5-10 % relative error,
    thread-count dependent, >= 4 threads) on the build container's CPU, so the oracle always takes ATen's
    native path for this one op.  See DESIGN.md "reference CPU path: oneDNN deconvolution defect".
    """
    with torch.backends.mkldnn.flags(enabled=False):
        return F.conv_transpose1d(x, w, bias, stride=stride, padding=padding)


def hifigan_resblock(sd: Dict[str, Tensor], prefix: str, x: Tensor, k: int) -> Tensor:
    """HifiGanResidualBlock.forward -- HF:1359-1367."""
    for q, dil in enumerate(RESBLOCK_DILATIONS):
        r = x
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv1d(x, sd[f"{prefix}convs1.{q}.weight"], sd[f"{prefix}convs1.{q}.bias"],
                     dilation=dil, padding=(k * dil - dil) // 2)
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv1d(x, sd[f"{prefix}convs2.{q}.weight"], sd[f"{prefix}convs2.{q}.bias"], padding=(k - 1) // 2)
        x = x + r
    return x


def hifigan(sd: Dict[str, Tensor], mel: Tensor, polyphase: bool = False,
            stages: Optional[List[Tensor]] = None) -> Tensor:
    """FastSpeech2ConformerHifiGan.forward, batched, normalize_before=False -- HF:1451-1491.

    mel: (B, T, 80) -> waveform (B, 320*T + 80).  ``stages`` (optional list) receives each stage output.
    """
    h = mel.transpose(1, 2)
    h = F.conv1d(h, sd["vocoder.conv_pre.weight"], sd["vocoder.conv_pre.bias"], padding=3)
    if stages is not None:
        stages.append(h)
    for i, (s, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        h = F.leaky_relu(h, LRELU_SLOPE)
        w, b = sd[f"vocoder.upsampler.{i}.weight"], sd[f"vocoder.upsampler.{i}.bias"]
        if polyphase:
            h = conv_transpose1d_polyphase(h, w, b, s, (k - s) // 2)
        else:
            h = conv_transpose1d(h, w, b, s, (k - s) // 2)
        acc = None
        for j, rk in enumerate(RESBLOCK_KERNELS):
            y = hifigan_resblock(sd, f"vocoder.resblocks.{i * len(RESBLOCK_KERNELS) + j}.", h, rk)
            acc = y if acc is None else acc + y
        h = acc / len(RESBLOCK_KERNELS)
        if stages is not None:
            stages.append(h)
    h = F.leaky_relu(h)  # default slope 0.01 -- HF:1480
    h = F.conv1d(h, sd["vocoder.conv_post.weight"], sd["vocoder.conv_post.bias"], padding=3)
    return torch.tanh(h).squeeze(1)


def waveform_lengths(spec_lengths: Tensor) -> Tensor:
    """_get_waveform_lengths -- models.py:211-221: five (L-1)*s - 2*((k-s)//2) + k steps = 320*T + 80."""
    for k, s in zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES):
        spec_lengths = (spec_lengths - 1) * s - 2 * ((k - s) // 2) + k
    return spec_lengths


def resynthesize(sd: Dict[str, Tensor], ids: Tensor, x0: Tensor, dt: float = 0.1,
                 truncation_value: Optional[float] = None) -> List[Tensor]:
    """ConditionalFlowMatchingWithHifiGan.forward -- models.py:223-256 (x0 injected)."""
    mel = sample(sd, ids, x0, dt, truncation_value)
    pv = torch.tensor(pad_value(), dtype=mel.dtype, device=mel.device)
    lengths = mel.ne(pv).all(dim=2).sum(dim=1)
    wav_len = waveform_lengths(lengths)
    wav = hifigan(sd, mel)
    return [w[:n].unsqueeze(0) for w, n in zip(wav, wav_len)]


# --------------------------------------------------------------------------------------
# Log-mel front end (src/hifigan/data.py:17-53); the mel filter bank is third-party (librosa, not installed here):
# restated from librosa.filters.mel's published definition (Slaney scale, area normalisation) and pinned against the two
# librosa-compatible banks that are installed (transformers.audio_utils.mel_filter_bank slaney/slaney: 9e-10,
# torchaudio.functional.melscale_fbanks: 6e-8; tests/test_oracle_cpu.py) -- not against librosa's own output, the one
# remaining caveat; the rest of the function is pinned by running the live reference with this bank injected.
# --------------------------------------------------------------------------------------
def librosa_mel_filter_bank(sr: int = 16000, n_fft: int = 400, n_mels: int = 80, fmin: float = 0.0,
                            fmax: float = 8000.0) -> Tensor:
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0

    def hz_to_mel(f: float) -> float:
        return f / f_sp if f < min_log_hz else min_log_hz / f_sp + math.log(f / min_log_hz) / logstep

    def mel_to_hz(m: float) -> float:
        return f_sp * m if m < min_log_hz / f_sp else min_log_hz * math.exp(logstep * (m - min_log_hz / f_sp))

    lo, hi = hz_to_mel(fmin), hz_to_mel(fmax)
    mel_f = [mel_to_hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    nb = 1 + n_fft // 2
    freqs = [sr / 2.0 * k / (nb - 1) for k in range(nb)]
    w = torch.zeros(n_mels, nb, dtype=torch.float64)
    for i in range(n_mels):
        for k, f in enumerate(freqs):
            lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i])
            upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1])
            w[i, k] = max(0.0, min(lower, upper)) * 2.0 / (mel_f[i + 2] - mel_f[i])
    return w.float()


def mel_spectrogram(y: Tensor, mel_basis: Optional[Tensor] = None) -> Tensor:
    """mel_spectrogram -- hifigan/data.py:17-53 with its defaults: (B, T) -> (B, 80, 1 + (T - 400) // 320)."""
    basis = librosa_mel_filter_bank() if mel_basis is None else mel_basis
    spec = torch.stft(y, 400, hop_length=320, window=torch.hann_window(400).to(y), center=False, onesided=True,
                      return_complex=True).abs()
    return torch.log(torch.clamp(torch.matmul(basis.to(spec), spec), min=1e-5))


def to_dtype(sd: Dict[str, Tensor], dtype: torch.dtype) -> Dict[str, Tensor]:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# --------------------------------------------------------------------------------------
# FLOP accounting; shared by bench.py's roofline line and the tests
# --------------------------------------------------------------------------------------
def transformer_flops(n_frames: int, nfe: int, hoisted: bool = True) -> int:
    """Per utterance of padded length N: NFE*N*(19 103 232 + 4096 N); minus the hoisted cond projection."""
    per = 19_103_232 + 4096 * n_frames
    if hoisted:
        per -= 393_216
    return nfe * n_frames * per


def vocoder_flops(t_frames: int) -> int:
    fl = 2 * 7 * 80 * 512 * t_frames
    c, length = 512, t_frames
    for s, k in zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS):
        lout = (length - 1) * s - 2 * ((k - s) // 2) + k
        fl += 2 * c * (c // 2) * k * length
        fl += 252 * (c // 2) ** 2 * lout
        c, length = c // 2, lout
    fl += 2 * 16 * 7 * length
    return fl
