from cage_challenge_4_b200.wrappers import CybORG  # noqa: F401  (CybORG/env.py:30)
