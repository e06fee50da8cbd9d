"""numpy restatement of the counter-based RNG used for per-step noise (test infrastructure).

Philox4x32-10 (Salmon et al., SC'11) + Box-Muller, keyed so that the noise of sample ``s`` at step
``i`` does not depend on batch composition or GPU count:

    key     = (seed_lo, seed_hi)
    counter = (pixel * ceil(C/4) + c // 4, step, sample_global_idx, stream)      # 4 channels of one pixel per block

The reference draws torch.randn_like on whichever device it runs (gaussian_diffusion.py:431,591);
CPU mt19937 and CUDA Philox streams differ, so parity tests inject noise and this file only pins
the device generator itself (uint32 stream bit-exact, normals to fp32 tolerance).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [N,4] uint32, key [2] uint32 -> [N,4] uint32."""
    c = ctr.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = c[:, 0].astype(np.uint64) * M0
        p1 = c[:, 2].astype(np.uint64) * M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        with np.errstate(over="ignore"):
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c


def normals(seed: int, sample_idx: int, step: int, C: int, hw: int, stream: int = 0) -> np.ndarray:
    """[C, hw] standard normals (fp32) of one sample at one step.

    Element (c, pixel) is component c & 3 of the Philox block whose counter is
    (pixel * ceil(C/4) + c // 4, step, sample_idx, stream): one block per (pixel, channel quad)."""
    nq = (C + 3) // 4
    nctr = hw * nq
    ctr = np.zeros((nctr, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(nctr, dtype=np.uint32)
    ctr[:, 1] = np.uint32(step)
    ctr[:, 2] = np.uint32(sample_idx)
    ctr[:, 3] = np.uint32(stream)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    # Box-Muller on pairs (r0,r1) and (r2,r3); u in (0,1]: (x + 1) * 2^-32 evaluated in fp32 like the kernel
    u = (r.astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -32)   # may round to 1.0 -> log = 0
    u = np.minimum(u, np.float32(1.0))
    rad0 = np.sqrt(np.float32(-2.0) * np.log(u[:, 0]), dtype=np.float32)
    rad1 = np.sqrt(np.float32(-2.0) * np.log(u[:, 2]), dtype=np.float32)
    th0 = np.float32(2.0 * np.pi) * u[:, 1]
    th1 = np.float32(2.0 * np.pi) * u[:, 3]
    z = np.stack([rad0 * np.cos(th0), rad0 * np.sin(th0), rad1 * np.cos(th1), rad1 * np.sin(th1)], axis=1)   # [hw*nq, 4]
    z = z.reshape(hw, nq * 4).T[:C]                                        # [C, hw]
    return np.ascontiguousarray(z).astype(np.float32)
