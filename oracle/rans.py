"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the range coder behind
`compressai.ans.BufferedRansEncoder / RansDecoder` as the reference calls it
(/root/reference/models/dcae.py:722, 755-756 encode, :875-876, 893 decode).

compressai is a third-party dependency that is absent from /root/reference (README.md:30, unpinned) and cannot be
installed here, so this follows its PUBLISHED algorithm: ryg_rans `rans64.h` (state in [2^31, 2^63), 32-bit
renormalisation words, `scale_bits` = 16) wrapped by `rans_interface.cpp` (symbols pushed in order, coded in reverse
at flush; values outside the table leave through the sentinel symbol `cdf_size - 2` and a 4-bit bypass code).
**Parity unpinned** against the real extension: no golden bitstream exists in the reference; the checks are
round trips, the coder contract, and byte equality between this restatement and the native coder
(dcae_b200/csrc/rans_coder.cpp).  Python ints, small cases only.
"""
from __future__ import annotations

import struct
from typing import List, Sequence

PRECISION = 16
BYPASS_BITS = 4
BYPASS_MAX = (1 << BYPASS_BITS) - 1
RANS_L = 1 << 31


def encode(symbols: Sequence[int], indexes: Sequence[int], cdfs: Sequence[Sequence[int]], cdf_sizes: Sequence[int],
           offsets: Sequence[int]) -> bytes:
    """encode_with_indexes + flush."""
    steps = []                                   # (start, range, is_bypass)
    for sym, k in zip(symbols, indexes):
        cdf, max_value = cdfs[k], cdf_sizes[k] - 2
        value, raw = sym - offsets[k], 0
        if value < 0:
            raw, value = -2 * value - 1, max_value
        elif value >= max_value:
            raw, value = 2 * (value - max_value), max_value
        steps.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_BITS)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= BYPASS_MAX:
                steps.append((BYPASS_MAX, BYPASS_MAX + 1, True))
                val -= BYPASS_MAX
            steps.append((val, val + 1, True))
            for j in range(n_bypass):
                v = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX
                steps.append((v, v + 1, True))
    x, words = RANS_L, []                         # words are prepended: collect reversed
    for start, rng, bypass in reversed(steps):
        if not bypass:                            # Rans64EncPut(start, freq, scale_bits = 16)
            x_max = ((RANS_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:                                     # Rans64EncPutBits(val, 4)
            freq = 1 << (16 - BYPASS_BITS)
            x_max = ((RANS_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << BYPASS_BITS) | start
    words.append(x >> 32)                         # Rans64EncFlush: low word first in memory
    words.append(x & 0xFFFFFFFF)
    return struct.pack("<%dI" % len(words), *reversed(words))


class Decoder:
    def __init__(self, stream: bytes):
        self.words = list(struct.unpack("<%dI" % (len(stream) // 4), stream))
        self.pos = 2
        self.x = self.words[0] | (self.words[1] << 32)       # Rans64DecInit

    def _word(self) -> int:
        w = self.words[self.pos] if self.pos < len(self.words) else 0
        self.pos += 1
        return w

    def _bits(self) -> int:                                    # Rans64DecGetBits(4)
        val = self.x & BYPASS_MAX
        self.x >>= BYPASS_BITS
        if self.x < RANS_L:
            self.x = (self.x << 32) | self._word()
        return val

    def decode(self, indexes: Sequence[int], cdfs, cdf_sizes, offsets) -> List[int]:
        out = []
        for k in indexes:
            cdf, max_value = cdfs[k], cdf_sizes[k] - 2
            cum = self.x & ((1 << PRECISION) - 1)              # Rans64DecGet
            s = next(i for i in range(cdf_sizes[k]) if cdf[i] > cum) - 1
            start, freq = cdf[s], cdf[s + 1] - cdf[s]
            self.x = freq * (self.x >> PRECISION) + cum - start    # Rans64DecAdvance
            if self.x < RANS_L:
                self.x = (self.x << 32) | self._word()
            value = s
            if value == max_value:
                val = self._bits()
                n_bypass = val
                while val == BYPASS_MAX:
                    val = self._bits()
                    n_bypass += val
                raw = 0
                for j in range(n_bypass):
                    raw |= self._bits() << (j * BYPASS_BITS)
                value = raw >> 1
                value = -value - 1 if raw & 1 else value + max_value
            out.append(value + offsets[k])
        return out
