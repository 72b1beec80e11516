"""CybORG/Agents/SimpleAgents/RandomAgent.py:14-69 for the wrappers' Discrete action spaces: `get_action` samples the space
(RandomAgent.py:30-31).  The device-side batched form of cc4BlueRandomAgent (class first, then parameters) is
cage_challenge_4_b200.evaluation.RandomBluePolicy."""
from cage_challenge_4_b200.wrappers import BaseAgent


class RandomAgent(BaseAgent):
    def get_action(self, observation, action_space):
        return action_space.sample()


class cc4BlueRandomAgent(RandomAgent):
    pass
