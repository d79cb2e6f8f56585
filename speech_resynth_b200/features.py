"""Log-mel front end on the GPU: the reference's ``mel_spectrogram`` (src/hifigan/data.py:17-53), used there for feature
extraction (src/flow_matching/preprocess.py) and for the mel-L1 validation metric of the vocoder (src/hifigan/train.py:233-235).
Here it closes the loop on the device: resynthesised waveform -> log-mel -> compare with the mel the sampler produced.

The 80 x 201 mel filter bank is librosa's (``librosa.filters.mel``: Slaney mel scale, area normalisation), restated from its
published definition because librosa is not a dependency of this package.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from . import _native as nat

N_FFT, HOP, N_MELS, SAMPLE_RATE, FMIN, FMAX = 400, 320, 80, 16000, 0.0, 8000.0


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    """Slaney scale (librosa.hz_to_mel, htk=False): linear below 1 kHz (200/3 Hz per mel), logarithmic above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filter_bank(sr: int = SAMPLE_RATE, n_fft: int = N_FFT, n_mels: int = N_MELS, fmin: float = FMIN,
                    fmax: float = FMAX) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney', float32)."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


_TABLES: Dict[torch.device, Tuple[torch.Tensor, ...]] = {}


def _tables(device: torch.device):
    t = _TABLES.get(device)
    if t is None:
        ang = 2.0 * math.pi * torch.arange(N_FFT, dtype=torch.float64) / N_FFT
        t = (torch.hann_window(N_FFT).to(device), ang.cos().float().to(device), ang.sin().float().to(device),
             torch.from_numpy(mel_filter_bank()).to(device).contiguous())
        _TABLES[device] = t
    return t


@torch.inference_mode()
def mel_spectrogram(y: torch.Tensor, n_fft: int = N_FFT, num_mels: int = N_MELS, sampling_rate: int = SAMPLE_RATE,
                    hop_size: int = HOP, fmin=0, fmax=8000) -> torch.Tensor:
    """Same signature and result layout as the reference (hifigan/data.py:17-53): y (B, T) or (T,) float waveform on a
    CUDA device -> (B, 80, 1 + (T - 400) // 320) log-mel.  Only the reference's own hyper-parameters are built."""
    if (n_fft, num_mels, sampling_rate, hop_size, float(fmin), float(fmax)) != (N_FFT, N_MELS, SAMPLE_RATE, HOP, FMIN, FMAX):
        raise NotImplementedError("the kernel is specialised to n_fft 400, hop 320, 80 mels, 0-8000 Hz at 16 kHz")
    nat.require_blackwell()
    squeeze = y.dim() == 1
    y = (y.unsqueeze(0) if squeeze else y).contiguous().float()
    b, t = y.shape
    if t < N_FFT:
        raise ValueError("waveform shorter than one analysis window (torch.stft would fail too)")
    frames = 1 + (t - N_FFT) // HOP
    win, cs, sn, basis = _tables(y.device)
    out = torch.empty(b, N_MELS, frames, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        nat.call("srb_log_mel", nat.ptr(y), t, b, t, nat.ptr(win), nat.ptr(cs), nat.ptr(sn), nat.ptr(basis), nat.ptr(out), frames,
                 flops=2.0 * b * frames * (2 * 201 * 400 + 80 * 201), nbytes=4.0 * b * (t + 80 * frames))
    return out[0] if squeeze else out
