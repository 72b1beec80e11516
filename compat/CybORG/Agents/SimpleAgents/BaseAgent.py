from cage_challenge_4_b200.wrappers import BaseAgent  # noqa: F401
