"""Host post-processing helpers of the reference's tools: ``tools/audio/np.py`` and the WAV
container writer of ``tools/audio/pcm.py`` - numpy only (the reference JIT-compiles the same arithmetic with numba)."""
from __future__ import annotations

import io
import math
import wave

import numpy as np


def float_to_int16(audio: np.ndarray) -> np.ndarray:
    """tools/audio/np.py:6-11: scale by ``32767 * 32768 // (ceil(max|x|) * 32768)`` (peak <= 1 -> 32767) and truncate.
    Like the reference, an all-zero input divides by zero."""
    am = int(math.ceil(float(np.abs(audio).max())) * 32768)
    am = 32767 * 32768 // am
    return np.multiply(audio, am).astype(np.int16)


def pcm_to_wav_bytes(pcm: np.ndarray, sample_rate: int = 24000) -> bytes:
    """16-bit mono PCM -> RIFF/WAVE bytes (what ``tools/audio/pcm.py`` writes for the API / web examples)."""
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(np.ascontiguousarray(pcm, dtype=np.int16).tobytes())
    return buf.getvalue()


def strip_silence(wav: np.ndarray, threshold: float = 1e-5) -> np.ndarray:
    """core.py:261-265 (quirk Q20): drop every sample with |x| <= threshold, also inside speech."""
    return wav[np.abs(wav) > np.float32(threshold)]
