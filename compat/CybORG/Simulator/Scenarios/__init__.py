from cage_challenge_4_b200.wrappers import EnterpriseScenarioGenerator  # noqa: F401
