"""Batch resynthesis driver: unit sequences -> waveform files (the decoder half of src/flow_matching/synthesize.py:36-54).

The reference loops ``decoder(...)`` -> per-utterance ``hyp_wav.cpu()`` -> ``torchaudio.save(path, wav, 16000)`` on one
stream; once the GPU path takes ~30 ms per 64 x 10 s batch those synchronous read-backs and file writes are the
wall clock.  Here the same work is pipelined:

* utterances are length-bucketed (``sharding.bucket_by_length``: a bucket is padded to its longest member, exactly what
  ``pad_sequence`` gives the reference, ``synthesize.py:42``) so no bucket pays for another's padding;
* a bucket's waveforms leave the last kernel already cropped and back to back (``resynthesize_flat``), so they go
  device -> pinned host buffer as ONE copy on a second CUDA stream (two buffers in rotation) while the next bucket
  computes; the unit ids stay on the host until the decoder copies them, so it never waits for the GPU;
* a writer thread turns finished host buffers into RIFF/WAVE files (32-bit float PCM, what ``torchaudio.save`` writes for
  a float32 tensor; 16-bit on request).

The unit encoder (textlesslib mHuBERT + k-means, ``synthesize.py:26-31,38``) is a different model and stays outside:
callers hand over the unit ids (``units + 1``, 0 = pad, ``synthesize.py:39``).
"""
from __future__ import annotations

import os
import queue
import struct
import threading
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import sharding

SAMPLE_RATE = 16000


def write_wav(path: str, samples: np.ndarray, sample_rate: int = SAMPLE_RATE, bits: int = 32) -> None:
    """Mono RIFF/WAVE: bits = 32 -> IEEE float (format tag 3, as torchaudio.save does for float32 tensors),
    bits = 16 -> signed PCM (format tag 1, round to nearest, clipped to [-1, 1))."""
    x = np.asarray(samples, dtype=np.float32).reshape(-1)
    if bits == 32:
        tag, data = 3, x.astype("<f4").tobytes()
    elif bits == 16:
        tag, data = 1, np.clip(np.rint(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
    else:
        raise ValueError("bits must be 16 or 32")
    block = bits // 8
    fmt = struct.pack("<HHIIHH", tag, 1, sample_rate, sample_rate * block, block, bits)
    parent = os.path.dirname(path)
    if parent:
        os.makedirs(parent, exist_ok=True)
    with open(path, "wb") as f:
        if tag == 3:
            # non-PCM formats carry a fact chunk with the sample count
            fact = struct.pack("<4sII", b"fact", 4, x.size)
            riff_size = 4 + (8 + len(fmt)) + len(fact) + (8 + len(data))
            f.write(struct.pack("<4sI4s4sI", b"RIFF", riff_size, b"WAVE", b"fmt ", len(fmt)) + fmt + fact)
        else:
            riff_size = 4 + (8 + len(fmt)) + (8 + len(data))
            f.write(struct.pack("<4sI4s4sI", b"RIFF", riff_size, b"WAVE", b"fmt ", len(fmt)) + fmt)
        f.write(struct.pack("<4sI", b"data", len(data)))
        f.write(data)


class _Writer(threading.Thread):
    """Consumes (event, host buffer, [(path, offset, length)], slot) jobs: waits for the copy, writes the files, frees the slot."""

    def __init__(self, bits: int, free_slots: "queue.Queue[int]"):
        super().__init__(daemon=True)
        self.jobs: "queue.Queue" = queue.Queue()
        self.bits = bits
        self.free_slots = free_slots
        self.error: Optional[BaseException] = None

    def run(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            event, host, files, slot = job
            try:
                event.synchronize()
                flat = host.numpy()
                for path, off, n in files:
                    write_wav(path, flat[off: off + n], SAMPLE_RATE, self.bits)
            except BaseException as e:  # noqa: BLE001 - surfaced by synthesize_units
                self.error = e
            finally:
                self.free_slots.put(slot)


@torch.inference_mode()
def synthesize_units(decoder, units: Sequence[torch.Tensor], out_paths: Sequence[str], dt: float = 0.0625,
                     truncation_value: Optional[float] = 1.0, batch_size: int = 32, bits: int = 32,
                     granularity: int = 64) -> List[int]:
    """Resynthesise every unit sequence and write ``out_paths[i]``; returns the number of samples of each file.

    ``decoder`` is a ``ConditionalFlowMatchingWithHifiGan`` on a CUDA device; ``units[i]`` a 1-D tensor of ids
    (unit + 1); ``dt`` / ``truncation_value`` / ``batch_size`` default to configs/resynth/mhubert-expresso-2000.yaml:42-43,98.
    """
    if len(units) != len(out_paths):
        raise ValueError("one output path per unit sequence")
    device = decoder.device
    lengths = [int(u.numel()) for u in units]
    buckets = sharding.bucket_by_length(lengths, granularity=granularity, max_batch=batch_size)
    cap = max(sum(320 * lengths[i] + 80 for i in b.indices) for b in buckets) if buckets else 0
    copy_stream = torch.cuda.Stream(device=device)
    hosts = [torch.empty(cap, dtype=torch.float32).pin_memory() for _ in range(2)]
    free_slots: "queue.Queue[int]" = queue.Queue()
    for s in range(2):
        free_slots.put(s)
    writer = _Writer(bits, free_slots)
    writer.start()
    n_samples = [0] * len(units)
    try:
        for b in buckets:
            ids = sharding.pad_bucket(units, b, pin=True)
            flat, wav_lengths = decoder.resynthesize_flat(ids, dt, truncation_value)   # cropped, back to back, fresh storage
            slot = free_slots.get()                                  # blocks only if both host buffers are still in use
            if writer.error is not None:
                raise writer.error
            copy_stream.wait_stream(torch.cuda.current_stream(device))
            files, off = [], 0
            for i, n in zip(b.indices, wav_lengths):
                files.append((out_paths[i], off, n))
                n_samples[i] = n
                off += n
            with torch.cuda.stream(copy_stream):
                flat.record_stream(copy_stream)
                hosts[slot][:off].copy_(flat[:off], non_blocking=True)
                done = torch.cuda.Event()
                done.record(copy_stream)
            writer.jobs.put((done, hosts[slot], files, slot))
    finally:
        writer.jobs.put(None)
        writer.join()
    if writer.error is not None:
        raise writer.error
    return n_samples
