"""Host-side counter-based random numbers for the state-initialisation helpers (numpy, vectorised):
Philox4x32-10 and HOOMD's Seed / Counter / distributions (hoomd/RandomNumbers.h:63-87,153-187,264-270,
361-374; Salmon et al. SC'11). The step loop's draws happen on the device (csrc/rng.cuh); this module
serves the one-off helpers that the reference also runs on the host."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)
RNG_ID_THERMALIZE = 35  # RNGIdentifier::ParticleGroupThermalize, hoomd/RNGIdentifiers.h:60


def philox4x32_10(ctr, key):
    """ctr: (N, 4) uint32 counters, key: (2,) uint32. Returns (N, 4) uint32."""
    c = [np.asarray(ctr)[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(int(key[0])), np.uint64(int(key[1]))
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    for _ in range(10):
        p0 = m0 * c[0]
        p1 = m1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & M32,
             p0 & M32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return np.stack(c, axis=1).astype(np.uint32)


def seed_key(rng_id, timestep, seed):
    """hoomd/RandomNumbers.h:153-158: {id << 24 | seed16 << 8 | timestep bits 39..32, timestep low 32}"""
    return np.array([(int(rng_id) << 24) | ((int(seed) & 0xFFFF) << 8) | ((int(timestep) >> 32) & 0xFF),
                     int(timestep) & 0xFFFFFFFF], dtype=np.uint64).astype(np.uint32)


def normal_draws(rng_id, timestep, seed, tags, n_draws):
    """n_draws successive NormalDistribution(1) values of RandomGenerator(Seed(id, timestep, seed),
    Counter(tag)) for every tag: (N, n_draws) float64. One Philox call per draw (Box-Muller, the
    second value is discarded as in hoomd/RandomNumbers.h:361-374)."""
    tags = np.asarray(tags, dtype=np.uint32)
    key = seed_key(rng_id, timestep, seed)
    out = np.empty((len(tags), n_draws))
    ctr = np.zeros((len(tags), 4), dtype=np.uint32)
    ctr[:, 3] = tags  # Counter(a) -> {0, 0, 0, a}; the low word counts the draws
    for d in range(n_draws):
        ctr[:, 0] = d
        u = philox4x32_10(ctr, key).astype(np.uint64)
        u0 = (u[:, 0] << np.uint64(32)) | u[:, 1]
        u1 = (u[:, 2] << np.uint64(32)) | u[:, 3]
        # u64 -> (-1, 1): int64(in) * 2^-63 + 2^-64; u64 -> (0, 1): in * 2^-64 + 2^-65
        a = u0.astype(np.int64).astype(np.float64) * 2.0 ** -63 + 2.0 ** -64
        b = u1.astype(np.float64) * 2.0 ** -64 + 2.0 ** -65
        out[:, d] = np.sin(np.pi * a) * np.sqrt(-2.0 * np.log(b))
    return out


def thermal_velocities(tags, mass, kT, seed, timestep):
    """ParticleGroup::thermalizeParticleMomenta (hoomd/ParticleGroup.cc:690-806) for point particles in
    3D, all particles in the group: v = N(0, sqrt(kT / m)) per component from the particle's own
    stream, then the centre-of-mass momentum is removed. Returns (N, 3) float64 in the order of `tags`."""
    mass = np.asarray(mass, dtype=np.float64)
    v = normal_draws(RNG_ID_THERMALIZE, timestep, seed, tags, 3) * np.sqrt(kT / mass)[:, None]
    com = (mass[:, None] * v).sum(axis=0) / len(mass)
    return v - com / mass[:, None]
