from cage_challenge_4_b200.wrappers import BlueEnterpriseWrapper, BlueFixedActionWrapper, BlueFlatWrapper, EnterpriseMAE  # noqa: F401
