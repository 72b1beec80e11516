from cage_challenge_4_b200.wrappers import DiscoveryFSRed  # noqa: F401
