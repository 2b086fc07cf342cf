"""Counter-based random numbers of the Dropout2d mask kernel, restated in numpy.  TEST INFRASTRUCTURE ONLY.

The reference draws its Dropout2d channel masks (cm/models/pspnet.py:49,55,64-73) from torch's CUDA generator; that
stream cannot be reproduced by another kernel, so network parity runs inject masks.  The product's own mask generator
(`hn_dropout2d_scale`) is Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) keyed by a seed with the counter (call index, element index); this file restates that published algorithm so
the kernel can be checked bit for bit.  Known-answer vectors of the Random123 distribution pin the restatement
(tests/test_oracle.py::test_philox_known_answers).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: (k0, k1) -> uint32 array [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def dropout2d_scale(seed: int, call: int, n: int, p: float) -> np.ndarray:
    """The [n] FP32 scale vector of call number `call`: 1/(1-p) where the 24-bit uniform is >= p, else 0."""
    i = np.arange(n, dtype=np.uint64)
    ctr = np.stack([i & MASK32, i >> np.uint64(32), np.full(n, call & 0xFFFFFFFF, dtype=np.uint64),
                    np.full(n, (call >> 32) & 0xFFFFFFFF, dtype=np.uint64)], axis=-1)
    w0 = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[..., 0]
    u = (w0 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    keep = np.float32(1.0) / (np.float32(1.0) - np.float32(p)) if p < 1.0 else np.float32(0.0)
    return np.where(u >= np.float32(p), keep, np.float32(0.0)).astype(np.float32)
