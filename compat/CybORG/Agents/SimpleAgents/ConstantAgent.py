from cage_challenge_4_b200.wrappers import SleepAgent  # noqa: F401
