from cage_challenge_4_b200.wrappers import EnterpriseMAE  # noqa: F401
