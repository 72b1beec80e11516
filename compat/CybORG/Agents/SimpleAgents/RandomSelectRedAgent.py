from cage_challenge_4_b200.wrappers import RandomSelectRedAgent  # noqa: F401
