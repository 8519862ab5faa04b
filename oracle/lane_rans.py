"""CPU restatement of the device coder's "lane container" (csrc/rans_device.cu) -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import this module; the product never does.

The probability model is the reference's: 16-bit quantised CDF rows selected by ``indexes``
(compressai/entropy_models/entropy_models.py:206-235) and the escape scheme of
``compressai/cpp_exts/rans/rans_interface.cpp:117-171`` (a symbol outside its row's range codes the row's last entry followed
by 4-bit bypass nibbles: the nibble count in base 15, then the raw value's nibbles, least significant first).  What differs from
the reference's coder (rans_interface.cpp:108-200 over third_party/ryg_rans/rans64.h: ONE 64-bit-state chain with 32-bit words
per image) is the container: the symbols are dealt round-robin onto S independent lanes, each a word-renormalised rANS with a
32-bit state in [2^16, 2^32) and 16-bit words (the `rans_word` configuration of ryg_rans), so that a GPU codes the lanes in
parallel.  This container has no counterpart in the reference -- **parity unpinned** for the byte layout itself; what IS pinned:
the token sequence (checked against the host coder's decode of the reference-compatible stream in the tests) and the round trip.

Layout (little endian): u32 "MMCL" | u32 n | u32 S | u32 0 | u32 state[S] | u32 words[S] | u16 payload[...] | pad to 4 bytes.
Pure-Python loops: small cases only.
"""
from __future__ import annotations

import struct
from typing import List, Sequence

MAGIC = b"MMCL"
PRECISION = 16
BYPASS_BITS = 4
MAX_BYPASS = (1 << BYPASS_BITS) - 1
LANE_L = 1 << 16


def lanes_default(n: int) -> int:
    """The smallest power of two in [4, 1024] that keeps a lane at <= 8192 symbols; small tensors are split further, up to 64
    lanes, while a lane keeps >= 2048 symbols."""
    s = 4
    while s < 1024 and s * 8192 < n:
        s *= 2
    while s < 64 and n // (2 * s) >= 2048:
        s *= 2
    return s


def tokens_of(symbol: int, index: int, cdfs, sizes, offsets) -> List[tuple]:
    """(start, freq, bits) tokens of one symbol in stream (decode) order -- rans_interface.cpp:117-171."""
    cdf = cdfs[index]
    max_value = int(sizes[index]) - 2
    value = int(symbol) - int(offsets[index])
    raw = 0
    if value < 0:
        raw = -2 * value - 1
        value = max_value
    elif value >= max_value:
        raw = 2 * (value - max_value)
        value = max_value
    out = [(int(cdf[value]), int(cdf[value + 1]) - int(cdf[value]), PRECISION)]
    if value == max_value:
        n_bypass = 0
        while (raw >> (n_bypass * BYPASS_BITS)) != 0:
            n_bypass += 1
        v = n_bypass
        while v >= MAX_BYPASS:
            out.append((MAX_BYPASS, 1, BYPASS_BITS))
            v -= MAX_BYPASS
        out.append((v, 1, BYPASS_BITS))
        for j in range(n_bypass):
            out.append(((raw >> (j * BYPASS_BITS)) & MAX_BYPASS, 1, BYPASS_BITS))
    return out


def encode(symbols: Sequence[int], indexes: Sequence[int], cdfs, sizes, offsets, lanes: int | None = None) -> bytes:
    n = len(symbols)
    S = lanes if lanes is not None else lanes_default(n)
    states, payloads = [], []
    for lane in range(S):
        toks = []
        for i in range(lane, n, S):
            toks.extend(tokens_of(symbols[i], indexes[i], cdfs, sizes, offsets))
        x, emitted = LANE_L, []
        for start, freq, bits in reversed(toks):
            x_max = ((LANE_L >> bits) << 16) * freq
            if x >= x_max:
                emitted.append(x & 0xFFFF)
                x >>= 16
            x = ((x // freq) << bits) + (x % freq) + start
        states.append(x)
        payloads.append(list(reversed(emitted)))          # decoder read order
    out = bytearray(MAGIC + struct.pack("<III", n, S, 0))
    out += struct.pack(f"<{S}I", *states)
    out += struct.pack(f"<{S}I", *[len(p) for p in payloads])
    for p in payloads:
        out += struct.pack(f"<{len(p)}H", *p)
    while len(out) % 4:
        out += b"\0"
    return bytes(out)


def decode(stream: bytes, indexes: Sequence[int], cdfs, sizes, offsets) -> List[int]:
    if stream[:4] != MAGIC:
        raise ValueError("not a lane container")
    n, S, _ = struct.unpack_from("<III", stream, 4)
    if n != len(indexes):
        raise ValueError("symbol count mismatch")
    states = struct.unpack_from(f"<{S}I", stream, 16)
    words = struct.unpack_from(f"<{S}I", stream, 16 + 4 * S)
    pos = 16 + 8 * S
    out = [0] * n
    for lane in range(S):
        w = struct.unpack_from(f"<{words[lane]}H", stream, pos)
        pos += 2 * words[lane]
        x, rp = states[lane], 0

        def pull(bits):
            nonlocal x, rp
            val = x & ((1 << bits) - 1)
            x >>= bits
            if x < LANE_L:
                x = (x << 16) | w[rp]
                rp += 1
            return val

        for i in range(lane, n, S):
            idx = indexes[i]
            cdf, max_value = cdfs[idx], int(sizes[idx]) - 2
            cum = x & 0xFFFF
            s = 0
            while s + 1 <= max_value + 1 and int(cdf[s + 1]) <= cum:
                s += 1
            start, freq = int(cdf[s]), int(cdf[s + 1]) - int(cdf[s])
            x = freq * (x >> PRECISION) + cum - start
            if x < LANE_L:
                x = (x << 16) | w[rp]
                rp += 1
            value = s
            if value == max_value:
                val = pull(BYPASS_BITS)
                n_bypass = val
                while val == MAX_BYPASS:
                    val = pull(BYPASS_BITS)
                    n_bypass += val
                raw = 0
                for j in range(n_bypass):
                    raw |= pull(BYPASS_BITS) << (j * BYPASS_BITS)
                value = raw >> 1
                value = -value - 1 if (raw & 1) else value + max_value
            out[i] = value + int(offsets[idx])
        if rp != len(w) or x != LANE_L:
            raise ValueError("lane %d did not end on its initial state" % lane)
    return out
