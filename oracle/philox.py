"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based RNG used by the
CUDA renderer (moleculardiffusion_mivit_b200/csrc/philox.cuh), so that noisy renders
of the GPU path can be checked draw-for-draw and not only statistically.

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
same constants as curand_philox4x32_x.h / Random123).  The reference itself draws
from the global np.random state (helpers/helpersGeneration.py:300,312,317), which is
not reproducible on a GPU; parity with the reference is therefore statistical
(moments + KS, BASELINE.json north_star) and parity GPU<->oracle is exact-by-stream.

Stream layout (identical in philox.cuh):
    key     = (seed_lo, seed_hi)
    counter = (item, block, seq_id, stream | variant << 8)
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

# stream ids (low byte of counter word 3)
STREAM_TRAJ = 0       # Brownian steps      item = t,              2 normals per call pair
STREAM_D = 1          # diffusion coeff     item = 0
STREAM_INTENSITY = 2  # spot intensity      item = frame*n + p  (V1)  /  frame (PSFNoise)
STREAM_PIXEL = 3      # background + Poisson, item = frame*P*P + pixel
STREAM_LOCERR = 4     # localisation error  item = frame


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    shape = np.broadcast(c0, c1, c2, c3).shape
    c0, c1, c2, c3 = [np.broadcast_to(c, shape).copy() for c in (c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + PHILOX_W0)
            k1 = np.uint32(k1 + PHILOX_W1)
    return c0, c1, c2, c3


def u01(x):
    """uint32 -> float32 uniform in (0,1]: (x + 1) * 2^-32 computed like the kernel
    (x * 2^-32 + 2^-33 in fp32, curand's _curand_uniform)."""
    x = np.asarray(x, dtype=np.uint32)
    return (x.astype(np.float32) * np.float32(2.3283064365386963e-10)
            + np.float32(2.3283064365386963e-10 / 2.0)).astype(np.float32)


def box_muller(xa, xb):
    """Two uint32 -> two float32 standard normals (same formula as philox.cuh)."""
    u = u01(xa)
    v = u01(xb).astype(np.float32) * np.float32(6.2831853071795860)
    r = np.sqrt(np.float32(-2.0) * np.log(u)).astype(np.float32)
    return (r * np.sin(v)).astype(np.float32), (r * np.cos(v)).astype(np.float32)


def box_muller_shifted(xa, xb):
    """box_muller_fast of philox.cuh: r = sqrt(-2 ln 2 * log2(u)), angle (v - 1/2) * 2 pi in (-pi, pi]."""
    u = u01(xa)
    th = ((u01(xb) - np.float32(0.5)) * np.float32(6.2831853071795860)).astype(np.float32)
    r = np.sqrt((np.float32(-1.3862943611198906) * np.log2(u).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return (r * np.sin(th)).astype(np.float32), (r * np.cos(th)).astype(np.float32)


def seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)


def stream_word(stream, variant=0):
    return np.uint32((int(stream) & 0xFF) | ((int(variant) & 0xFFFFFF) << 8))
