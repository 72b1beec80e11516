"""gym 0.26.2 `np_random`: Generator(PCG64(SeedSequence(seed)))."""
import numpy as np

RandomNumberGenerator = np.random.Generator


def np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = RandomNumberGenerator(np.random.PCG64(seed_seq))
    return rng, np_seed
