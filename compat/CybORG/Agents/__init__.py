"""CybORG/Agents/__init__.py: the agent classes the CC4 scenario generator and the evaluation use."""
from .SimpleAgents.BaseAgent import BaseAgent  # noqa: F401
from .SimpleAgents.ConstantAgent import SleepAgent  # noqa: F401
from .SimpleAgents.EnterpriseGreenAgent import EnterpriseGreenAgent  # noqa: F401
from .SimpleAgents.FiniteStateRedAgent import FiniteStateRedAgent  # noqa: F401
from .SimpleAgents.FSMRedVariants import DiscoveryFSRed  # noqa: F401
from .SimpleAgents.RandomAgent import RandomAgent, cc4BlueRandomAgent  # noqa: F401
from .SimpleAgents.RandomSelectRedAgent import RandomSelectRedAgent  # noqa: F401
