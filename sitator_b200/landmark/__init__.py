from .errors import StaticLatticeError, ZeroLandmarkError, LandmarkAnalysisError
from .LandmarkAnalysis import LandmarkAnalysis
