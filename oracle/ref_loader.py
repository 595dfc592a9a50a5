"""Import the compiled, unmodified reference from ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY (used by tests/, bench.py's cpu_baseline / --impl reference
legs and the golden-vector generator).  Nothing under ``sitator_b200/`` imports this.

The reference imports ``ase`` and ``matplotlib`` at module scope
(``sitator/SiteNetwork.py:8-12``, ``sitator/visualization/common.py:3-5``); neither
is installed here and neither is on the landmark-analysis path.  This loader
installs minimal stand-ins for them (and for ``sitator.visualization`` and the
``sitator.dynamics`` package ``__init__``, whose imports drag in ASE
calculators), restores the NumPy aliases the reference still uses
(``np.int/np.float/np.bool``; e.g. ``LandmarkAnalysis.py:195-196``), and then
imports the compiled modules.  No reference logic is altered.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_loaded = None


class StubAtoms(object):
    """Just enough of ``ase.Atoms`` for ``SiteNetwork`` (SiteNetwork.py:58-67)."""

    def __init__(self, positions=None, numbers=None, cell=None, pbc=True, symbols=None):
        import numpy as np
        self.positions = np.array(positions, dtype=float).reshape(-1, 3)
        n = len(self.positions)
        self.numbers = np.zeros(n, dtype=int) if numbers is None else np.array(numbers, dtype=int)
        self.cell = np.array(cell, dtype=float) if cell is not None else np.zeros((3, 3))
        self.pbc = pbc

    def __len__(self):
        return len(self.positions)

    def copy(self):
        return StubAtoms(self.positions.copy(), self.numbers.copy(), self.cell.copy(), self.pbc)

    def __delitem__(self, mask):
        import numpy as np
        mask = np.asarray(mask)
        keep = ~mask if mask.dtype == bool else np.setdiff1d(np.arange(len(self)), mask)
        self.positions = self.positions[keep]
        self.numbers = self.numbers[keep]

    def get_positions(self):
        return self.positions.copy()

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_cell(self):
        return self.cell.copy()

    def get_masses(self):
        import numpy as np
        return np.ones(len(self))

    def get_chemical_symbols(self):
        return ["X"] * len(self)


def _install_stubs():
    import numpy as np
    from unittest import mock

    for alias, typ in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)

    if "ase" not in sys.modules:
        ase = types.ModuleType("ase")
        ase.Atoms = StubAtoms
        ase_io = types.ModuleType("ase.io")
        ase_data = types.ModuleType("ase.data")
        ase_data.atomic_masses = np.ones(120)
        ase_data.chemical_symbols = ["X"] * 120
        ase.io = ase_io
        ase.data = ase_data
        sys.modules["ase"] = ase
        sys.modules["ase.io"] = ase_io
        sys.modules["ase.data"] = ase_data
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)

    # sitator.visualization: plotting only; provide inert names (decorators must pass through).
    if "sitator.visualization" not in sys.modules:
        viz = types.ModuleType("sitator.visualization")

        def plotter(is3D=True, **outer):
            def wrap(func):
                return func
            return wrap

        viz.plotter = plotter
        viz.plot_atoms = viz.plot_points = viz.layers = viz.grid = viz.set_axes_equal = (lambda *a, **k: None)
        viz.DEFAULT_COLORS = []
        viz.SiteNetworkPlotter = mock.MagicMock(name="SiteNetworkPlotter")
        viz.SiteTrajectoryPlotter = mock.MagicMock(name="SiteTrajectoryPlotter")
        sys.modules["sitator.visualization"] = viz


def available():
    try:
        from . import build_ref
    except ImportError:
        import build_ref
    return build_ref.is_built()


def load():
    """Return a namespace with the reference's classes (compiled from /root/reference)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("oracle/_ref is not built; run `python oracle/build_ref.py` "
                          "in a container that has /root/reference")
    os.environ.setdefault("SITATOR_PROGRESSBAR", "false")
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import sitator  # compiled __init__
    sitator.visualization = sys.modules["sitator.visualization"]
    # namespace stub for the sitator.dynamics package (see build_ref.MODULES)
    if "sitator.dynamics" not in sys.modules:
        dyn = types.ModuleType("sitator.dynamics")
        dyn.__path__ = [os.path.join(REF_DIR, "sitator", "dynamics")]
        sys.modules["sitator.dynamics"] = dyn
        sitator.dynamics = dyn
    from sitator.landmark import LandmarkAnalysis
    from sitator.landmark import helpers
    from sitator.landmark.cluster import mcl as cluster_mcl
    from sitator.util import PBCCalculator, DotProdClassifier
    from sitator.util.mcl import markov_clustering
    from sitator.dynamics.JumpAnalysis import JumpAnalysis
    from sitator.dynamics.RemoveUnoccupiedSites import RemoveUnoccupiedSites
    sitator.dynamics.RemoveUnoccupiedSites = RemoveUnoccupiedSites      # what the package __init__ would export
    from sitator.dynamics.SmoothSiteTrajectory import SmoothSiteTrajectory
    # site merging: namespace stub for sitator.network, the class names the package
    # __init__ files would export, then the two modules
    sitator.dynamics.JumpAnalysis = JumpAnalysis
    if "sitator.network" not in sys.modules:
        net = types.ModuleType("sitator.network")
        net.__path__ = [os.path.join(REF_DIR, "sitator", "network")]
        sys.modules["sitator.network"] = net
        sitator.network = net
    try:
        from sitator.network import merging
        from sitator.dynamics.MergeSitesByDynamics import MergeSitesByDynamics
    except ImportError:                       # an oracle/_ref built before these modules were added
        merging = MergeSitesByDynamics = None
    from sitator import SiteNetwork, SiteTrajectory
    import sitator.errors as errors
    import sitator.landmark.errors as lerrors

    ns = types.SimpleNamespace(
        sitator=sitator, LandmarkAnalysis=LandmarkAnalysis, helpers=helpers,
        cluster_mcl=cluster_mcl, PBCCalculator=PBCCalculator,
        DotProdClassifier=DotProdClassifier, markov_clustering=markov_clustering,
        JumpAnalysis=JumpAnalysis, RemoveUnoccupiedSites=RemoveUnoccupiedSites,
        SmoothSiteTrajectory=SmoothSiteTrajectory, merging=merging, MergeSitesByDynamics=MergeSitesByDynamics,
        SiteNetwork=SiteNetwork, SiteTrajectory=SiteTrajectory,
        errors=errors, landmark_errors=lerrors, Atoms=StubAtoms,
    )
    _loaded = ns
    return ns
