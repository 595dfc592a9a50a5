"""Landmark-analysis exceptions (mirrors reference ``sitator/landmark/errors.py:2-40``)."""


class LandmarkAnalysisError(Exception):
    pass


class StaticLatticeError(LandmarkAnalysisError):
    """A static-lattice atom broke the limits on its movement (``landmark/errors.py:5-24``).

    Attributes:
        lattice_atoms: indexes (into the static lattice) of the offending atoms.
        frame: frame at which it happened.
    """
    TRY_RECENTERING_MSG = "Try recentering the input trajectory (sitator.util.RecenterTrajectory)"

    def __init__(self, message, lattice_atoms=None, frame=None, try_recentering=False):
        if try_recentering:
            message = message + "\n" + StaticLatticeError.TRY_RECENTERING_MSG
        super().__init__(message)
        self.lattice_atoms = lattice_atoms
        self.frame = frame


class ZeroLandmarkError(LandmarkAnalysisError):
    """A landmark vector came out all zeros (``landmark/errors.py:26-40``)."""

    def __init__(self, mobile_index, frame):
        super().__init__(
            "Encountered a zero landmark vector for mobile ion %i at frame %i. Try increasing "
            "`cutoff_midpoint` and/or decreasing `cutoff_steepness`." % (mobile_index, frame))
        self.mobile_index = mobile_index
        self.frame = frame
