"""Stand-in for gymnasium (absent in this image); shares the gym stand-in's spaces."""
from gym import spaces
from gym.spaces import Space
