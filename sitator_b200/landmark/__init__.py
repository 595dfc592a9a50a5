from .errors import StaticLatticeError, ZeroLandmarkError, LandmarkAnalysisError
from .LandmarkAnalysis import LandmarkAnalysis
from .frames import ChunkedFrames
