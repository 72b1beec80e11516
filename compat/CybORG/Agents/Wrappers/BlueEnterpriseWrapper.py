from cage_challenge_4_b200.wrappers import BlueEnterpriseWrapper  # noqa: F401
