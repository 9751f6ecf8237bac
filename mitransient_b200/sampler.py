"""Host-side PCG32 / TEA (integer arithmetic only).

Used for the one place the reference's *host* logic touches the sampler: the per-pass seeds of a render
that does not fit in 2^32 samples (mitransient/integrators/common.py:66-72:
``seeds = mi.UInt32(sampler.next_1d() * 2**32)`` with the sampler seeded by ``(seed, num_passes)``).
"""
from __future__ import annotations

import struct

_M64 = (1 << 64) - 1
_M32 = (1 << 32) - 1
PCG32_MULT = 0x5851F42D4C957F2D


def sample_tea_32(v0: int, v1: int, rounds: int = 4):
    s = 0
    for _ in range(rounds):
        s = (s + 0x9E3779B9) & _M32
        v0 = (v0 + ((((v1 << 4) & _M32) + 0xA341316C) & _M32 ^ ((v1 + s) & _M32) ^ (((v1 >> 5) + 0xC8013EA4) & _M32))) & _M32
        v1 = (v1 + ((((v0 << 4) & _M32) + 0xAD90777D) & _M32 ^ ((v0 + s) & _M32) ^ (((v0 >> 5) + 0x7E95761E) & _M32))) & _M32
    return v0, v1


class PCG32:
    def __init__(self, initstate: int = 0x853C49E6748FEA9B, initseq: int = 0xDA3E39CB94B95BDB):
        self.seed(initstate, initseq)

    def seed(self, initstate: int, initseq: int):
        self.state = 0
        self.inc = ((initseq << 1) | 1) & _M64
        self.next_uint32()
        self.state = (self.state + initstate) & _M64
        self.next_uint32()

    def next_uint32(self) -> int:
        old = self.state
        self.state = (old * PCG32_MULT + self.inc) & _M64
        xorshifted = (((old >> 18) ^ old) >> 27) & _M32
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & _M32

    def next_float32(self) -> float:
        bits = (self.next_uint32() >> 9) | 0x3F800000
        return struct.unpack("<f", struct.pack("<I", bits))[0] - 1.0


def independent_sampler_first_float(seed: int, lane: int) -> float:
    """First ``next_1d()`` of lane ``lane`` of an ``independent`` sampler seeded with ``seed``."""
    v0, v1 = sample_tea_32(seed & _M32, lane & _M32)
    return PCG32(v0, v1).next_float32()
