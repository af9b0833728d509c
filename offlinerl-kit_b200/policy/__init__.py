from .base_policy import BasePolicy
from .sac import SACPolicy
from .cql import CQLPolicy
from .td3bc import TD3BCPolicy
from .iql import IQLPolicy
from .mopo import MOPOPolicy
from .edac import EDACPolicy
from .combo import COMBOPolicy

__all__ = ["BasePolicy", "SACPolicy", "CQLPolicy", "TD3BCPolicy", "IQLPolicy", "MOPOPolicy", "EDACPolicy", "COMBOPolicy"]
