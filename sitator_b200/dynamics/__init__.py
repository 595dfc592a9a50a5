from .JumpAnalysis import JumpAnalysis
