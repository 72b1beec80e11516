from gym.spaces import *  # noqa
from gym.spaces import Space, Discrete, MultiBinary, MultiDiscrete, Box
