"""base16384 text codec used by every string-typed asset of the reference (speaker embeddings, speaker prompts,
``spk_stat``, DVAE ``coef``): ``pybase16384.encode_to_string`` / ``decode_from_string`` as called at
``ChatTTS/model/speaker.py:14,100,113,145,157`` and ``ChatTTS/model/dvae.py:226,252``.

``pybase16384`` is a third-party dependency that is absent here (requirements.txt, unpinned), so this restates the
published base16384 format: the byte string is read as a big-endian bit stream and cut into 14-bit groups, each
emitted as the code point ``0x4E00 + group``; 7 bytes make 4 characters.  A tail of ``n = len % 7`` bytes is
zero-padded to ``ceil(8 n / 14)`` characters and followed by the marker character ``0x3D00 + n``.
Pinned by decoding the reference's own strings (``tests/test_speaker.py``): the payloads are raw LZMA2 streams /
fp16 tables whose decompression and sizes only come out right if every bit is in place.
"""
from __future__ import annotations

import numpy as np

_BASE = 0x4E00
_MARK = 0x3D00
_TAIL_CHARS = (0, 1, 2, 2, 3, 3, 4)          # characters that carry a tail of n bytes


def encode_to_string(data: bytes) -> str:
    raw = np.frombuffer(bytes(data), dtype=np.uint8)
    n_tail = raw.size % 7
    pad = (7 - n_tail) % 7
    bits = np.unpackbits(np.concatenate([raw, np.zeros(pad, np.uint8)]))            # MSB first
    groups = bits.reshape(-1, 14).astype(np.uint32) @ (1 << np.arange(13, -1, -1, dtype=np.uint32))
    if n_tail:
        groups = groups[:groups.size - 4 + _TAIL_CHARS[n_tail]]
    out = "".join(map(chr, (groups + _BASE).tolist()))
    return out + chr(_MARK + n_tail) if n_tail else out


def decode_from_string(text: str) -> bytes:
    if not text:
        return b""
    n_tail = 0
    last = ord(text[-1])
    if _MARK < last <= _MARK + 6:
        n_tail = last - _MARK
        text = text[:-1]
    codes = np.fromiter(map(ord, text), dtype=np.int64, count=len(text)) - _BASE
    if codes.size and (codes.min() < 0 or codes.max() >= 1 << 14):
        raise ValueError("not a base16384 string")
    full = codes.size - _TAIL_CHARS[n_tail]
    if full < 0 or full % 4:
        raise ValueError("base16384 string has a truncated group")
    want = full // 4 * 7 + n_tail
    pad = (-codes.size) % 4
    codes = np.concatenate([codes, np.zeros(pad, np.int64)])
    bits = ((codes[:, None] >> np.arange(13, -1, -1)) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(bits).tobytes()[:want]
