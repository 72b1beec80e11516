from cage_challenge_4_b200.wrappers import BlueFixedActionWrapper  # noqa: F401
