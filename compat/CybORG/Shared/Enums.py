from cage_challenge_4_b200.wrappers import TernaryEnum  # noqa: F401  (Shared/Enums.py:5)
