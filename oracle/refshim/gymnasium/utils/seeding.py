"""gymnasium.utils.seeding stand-in: PCG64 Generator, as gymnasium 0.29.1 does."""
import numpy as np


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and 0 <= seed):
        raise ValueError(f"Seed must be a non-negative integer or omitted, not {seed!r}")
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed
