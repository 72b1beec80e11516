from cage_challenge_4_b200.wrappers import BlueFlatWrapper  # noqa: F401
