from .base_policy import BasePolicy
from .sac import SACPolicy
from .cql import CQLPolicy

__all__ = ["BasePolicy", "SACPolicy", "CQLPolicy"]
