from cage_challenge_4_b200.wrappers import EnterpriseGreenAgent  # noqa: F401
