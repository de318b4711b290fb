"""Counter-based synthetic sector vectors: v(i) is a pure function of the 0-based GLOBAL index i, so every rank
generates exactly its own shard, results at different GPU counts are comparable element by element, and single
rows of H x v can be checked without materialising a vector that does not fit in host memory (Ns = 18: 38 GB).

    x = mix64((i + 1) * 0x9E3779B97F4A7C15 + seed * 0xBF58476D1CE4E5B9)        (arithmetic mod 2^64)
    y = mix64(x * 0x94D049BB133111EB + 0x2545F4914F6CDD1D)
    mix64(x): x ^= x >> 31; x *= 0xD6E8FEB86659FD93; x ^= x >> 32
    v(i) = scale * ( (x >> 11) * 2^-53 * 2 - 1  +  1j * ((y >> 11) * 2^-53 * 2 - 1) )

The test infrastructure carries its own C copy of this function (edo_counter_vec); tests/test_synth_cpu.py checks
that the three agree bit for bit."""
from __future__ import annotations

import numpy as np

_M1, _M2, _M3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0xD6E8FEB86659FD93
_M4, _M5 = 0x94D049BB133111EB, 0x2545F4914F6CDD1D


def default_scale(dim: int) -> float:
    """|re|,|im| uniform in [-1,1): E|v|^2 = 2/3 per element -> norm ~ 1."""
    return float(np.sqrt(1.5 / max(dim, 1)))


def counter_vec_numpy(i0: int, n: int, seed: int, scale: float) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(i0 + 1, i0 + n + 1, dtype=np.uint64)

        def mix(x):
            x = x ^ (x >> np.uint64(31))
            x = x * np.uint64(_M3)
            return x ^ (x >> np.uint64(32))

        x = mix(idx * np.uint64(_M1) + np.uint64((seed * _M2) & 0xFFFFFFFFFFFFFFFF))
        y = mix(x * np.uint64(_M4) + np.uint64(_M5))
        re = (x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 * 2.0 - 1.0
        im = (y >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 * 2.0 - 1.0
    return (scale * re + 1j * (scale * im)).astype(np.complex128)


def _s64(u: int) -> int:  # two's-complement int64 view of a 64-bit constant
    u &= 0xFFFFFFFFFFFFFFFF
    return u - (1 << 64) if u >= (1 << 63) else u


def counter_vec_torch(i0: int, n: int, seed: int, scale: float, device="cuda", chunk: int = 1 << 26):
    """complex128 tensor [n] on `device`; int64 arithmetic wraps mod 2^64, logical shifts are arithmetic shift + mask;
    generated in chunks so the int64 temporaries stay small next to a multi-GB shard."""
    import torch
    out = torch.empty(n, dtype=torch.complex128, device=device)
    ov = torch.view_as_real(out)

    def lsr(x, k):
        return (x >> k) & ((1 << (64 - k)) - 1)

    def mix(x):
        x = x ^ lsr(x, 31)
        x = x * _s64(_M3)
        return x ^ lsr(x, 32)

    add = _s64(seed * _M2)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        idx = torch.arange(i0 + a + 1, i0 + b + 1, dtype=torch.int64, device=device)
        x = mix(idx * _s64(_M1) + add)
        y = mix(x * _s64(_M4) + _s64(_M5))
        ov[a:b, 0] = (lsr(x, 11).to(torch.float64) * 2.0 ** -53 * 2.0 - 1.0) * scale
        ov[a:b, 1] = (lsr(y, 11).to(torch.float64) * 2.0 ** -53 * 2.0 - 1.0) * scale
    return out
