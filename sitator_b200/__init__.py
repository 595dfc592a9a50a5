"""sitator_b200 -- B200-native landmark analysis behind sitator's own API (LandmarkAnalysis.run path only)."""
from .SiteNetwork import SiteNetwork
from .SiteTrajectory import SiteTrajectory
from .structure import Atoms

__all__ = ["SiteNetwork", "SiteTrajectory", "Atoms"]
