"""dm_fp64 -- the 64-bit value fingerprint shared by the oracle and the CUDA path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference keeps Python ``set``s of
value strings (detectmatelibrary NewValueDetector; call site src/service/core.py:201-203),
which have no collisions.  The device path keeps 64-bit fingerprints; the oracle keeps
the exact byte strings *and* audits that no two distinct trained strings share a
fingerprint, so a fingerprint collision shows up as a test failure, not as a silent
false negative.

Definition (all arithmetic mod 2**32, words are little-endian, the value is
zero-padded to a multiple of 4 bytes, n = len(value)):

    stream A (murmur3-style):  h = 0x9747B28C
        for each word w:  k = rotl(w*0xCC9E2D51, 15)*0x1B873593
                          h = rotl(h ^ k, 13)*5 + 0xE6546B64
        h ^= n ; h = fmix32(h)               (murmur3 finaliser)
    stream B (xxh32-style):    g = 0x165667B1
        for each word w:  g = rotl(g + w*0x85EBCA77, 13)*0x9E3779B1
        g ^= n ; g = xxh32_avalanche(g)
    fp = (h << 32) | g ;  fp == 0 -> 1       (0 is the empty-slot marker)

The known-set key of (field f, value) is  fp ^ SALT[f]  with
SALT[f] = splitmix64(f + 1), again with 0 remapped to 1.
"""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & M32


def _fmix32(h: int) -> int:
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M32
    h ^= h >> 16
    return h


def _xxh_avalanche(g: int) -> int:
    g ^= g >> 15
    g = (g * 0x85EBCA77) & M32
    g ^= g >> 13
    g = (g * 0xC2B2AE3D) & M32
    g ^= g >> 16
    return g


def fp64(value: bytes) -> int:
    n = len(value)
    h = 0x9747B28C
    g = 0x165667B1
    pad = (-n) % 4
    data = value + b"\0" * pad
    for i in range(0, len(data), 4):
        w = int.from_bytes(data[i:i + 4], "little")
        k = (w * 0xCC9E2D51) & M32
        k = _rotl(k, 15)
        k = (k * 0x1B873593) & M32
        h ^= k
        h = _rotl(h, 13)
        h = (h * 5 + 0xE6546B64) & M32
        g = (g + w * 0x85EBCA77) & M32
        g = _rotl(g, 13)
        g = (g * 0x9E3779B1) & M32
    h = _fmix32(h ^ (n & M32))
    g = _xxh_avalanche(g ^ (n & M32))
    fp = (h << 32) | g
    return fp if fp else 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def field_salt(field_index: int) -> int:
    return splitmix64(field_index + 1)


def table_key(field_index: int, value: bytes) -> int:
    k = fp64(value) ^ field_salt(field_index)
    return k if k else 1
