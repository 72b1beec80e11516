from cage_challenge_4_b200.wrappers import FiniteStateRedAgent  # noqa: F401
