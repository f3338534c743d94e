"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the device trajectory source
(csrc/render.cu brownian_kernel), i.e. of helpers/helpersGeneration.py:9-45
brownian_motion with the per-sequence D of the andi_datasets call sites
(Experiments/PSFNoise/trainModelsPSFNoise.py:128-132): D ~ N(mean, sqrt(var)) redrawn until
positive, steps ~ N(0, 2D) per axis, cumulative sum from the origin, / div.

Parity status: the third-party generator (andi_datasets, no version pinned anywhere in the
reference) is absent, so bit-level parity with it is UNPINNED; what is pinned is the
statistical contract the reference's own notebooks check (tests/Simulator_tests/AnDi-Tests.ipynb
cells 4,6: MSD-recovered D, step std = sqrt(2D)) -- see tests/test_oracle_trajectory.py."""
import numpy as np

from . import philox as px

F32 = np.float32


def brownian_oracle(N, T, group_mean, group_var, div, seed, seq_offset=0):
    k0, k1 = px.seed_key(seed)
    gm = np.asarray(group_mean, dtype=F32)
    gv = np.asarray(group_var, dtype=F32)
    traj = np.zeros((N, T, 2), dtype=np.float64)
    D_out = np.zeros(N, dtype=F32)
    t = np.arange(T, dtype=np.uint32)
    for s in range(N):
        gid = seq_offset + s
        seq = np.uint32(gid & 0xFFFFFFFF)
        grp = gid % len(gm)
        mu, sd = gm[grp], np.sqrt(gv[grp]).astype(F32)
        D = F32(0)
        for blk in range(64):
            w = px.philox4x32_10(np.uint32(0), np.uint32(blk), seq, px.stream_word(px.STREAM_D), k0, k1)
            z0, z1 = px.box_muller(w[0], w[1])
            z2, z3 = px.box_muller(w[2], w[3])
            for z in (z0, z1, z2, z3):
                if not D > 0:
                    D = F32(mu + F32(sd * F32(z)))
            if D > 0:
                break
        D_out[s] = D
        sig = np.sqrt(F32(2.0) * D).astype(F32)
        w = px.philox4x32_10(t, np.uint32(0), seq, px.stream_word(px.STREAM_TRAJ), k0, k1)
        zx, zy = px.box_muller(w[0], w[1])
        steps = np.stack([(sig * zx).astype(F32), (sig * zy).astype(F32)], axis=1).astype(np.float64)
        steps[0] = 0.0
        traj[s] = np.cumsum(steps, axis=0) / div
    return traj, D_out
