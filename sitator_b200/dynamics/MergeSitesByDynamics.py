"""``MergeSitesByDynamics``: merge sites by Markov clustering of the jump statistics (mirrors reference
``sitator/dynamics/MergeSitesByDynamics.py:12-153``).

The connectivity matrix (default: ``n_ij`` of :class:`JumpAnalysis`) is cut at ``distance_threshold`` and clustered by
the same device Markov clustering the landmark path uses (``csrc/sitb_mcl.cu`` through ``sitator_b200.util.mcl``).
The reference's constructor raises ``NameError`` (``self.iterlimit = iterlimit`` names a parameter it does not have,
``MergeSitesByDynamics.py:54``); here ``iterlimit`` is a real parameter (default: Markov clustering's own 100).
"""
import logging

import numpy as np

from ..network.merging import MergeSites, pbc_distances
from ..util.mcl import markov_clustering
from .JumpAnalysis import JumpAnalysis

logger = logging.getLogger(__name__)


class MergeSitesByDynamics(MergeSites):
    def __init__(self, connectivity_matrix_generator=None, distance_threshold=1.0, post_check_thresh_factor=1.5,
                 check_types=True, markov_parameters={}, iterlimit=None, **kwargs):
        super().__init__(maximum_merge_distance=post_check_thresh_factor * distance_threshold, check_types=check_types,
                         **kwargs)
        if connectivity_matrix_generator is None:
            connectivity_matrix_generator = MergeSitesByDynamics.connectivity_n_ij
        assert callable(connectivity_matrix_generator)
        self.connectivity_matrix_generator = connectivity_matrix_generator
        self.distance_threshold = distance_threshold
        self.post_check_thresh_factor = post_check_thresh_factor
        self.check_types = check_types
        self.iterlimit = iterlimit
        self.markov_parameters = markov_parameters

    # -- connectivity matrix generation schemes (MergeSitesByDynamics.py:60-108)
    @staticmethod
    def connectivity_n_ij(sn):
        """Uses ``n_ij`` directly as connectivity matrix."""
        return sn.n_ij

    @staticmethod
    def connectivity_jump_lag_biased(jump_lag_coeff=1.0, jump_lag_sigma=20.0, jump_lag_cutoff=np.inf,
                                     distance_coeff=0.5, distance_sigma=1.0):
        """``p_ij`` biased by Gaussians of the jump lag and of the site distance (``:70-108``)."""
        def cfunc(sn):
            jl = np.array(sn.jump_lag, dtype=np.float64)
            jl -= 1.0
            jl /= jump_lag_sigma
            np.square(jl, out=jl)
            jl *= -0.5
            np.exp(jl, out=jl)
            jl[sn.jump_lag > jump_lag_cutoff] = 0.
            dmat = pbc_distances(np.asarray(sn.structure.cell), sn.centers, sn.centers)
            dmat /= distance_sigma
            np.square(dmat, out=dmat)
            dmat *= -0.5
            np.exp(dmat, out=dmat)
            return (sn.p_ij + jump_lag_coeff * jl) * (distance_coeff * dmat + (1 - distance_coeff))
        return cfunc

    def _get_sites_to_merge(self, st):
        sn = st.site_network
        if not sn.has_attribute('n_ij'):
            JumpAnalysis().run(st)
        connectivity_matrix = np.array(self.connectivity_matrix_generator(sn), dtype=np.float64, copy=True)
        n_sites_before = sn.n_sites
        assert n_sites_before == connectivity_matrix.shape[0]
        centers_before = np.asarray(sn.centers)

        # diagnostics threshold (:125-130)
        no_diag_graph = connectivity_matrix.copy()
        np.fill_diagonal(no_diag_graph, np.nan)
        edge_threshold = np.nanmean(no_diag_graph) + 3 * np.nanstd(no_diag_graph)

        # distance threshold (:133-143): all pairwise periodic distances in one device call
        dists = pbc_distances(np.asarray(sn.structure.cell), centers_before, centers_before)
        too_far = np.triu(dists > self.distance_threshold, 1)
        too_far = too_far | too_far.T
        alarming = too_far & (connectivity_matrix > edge_threshold)
        n_alarming_ignored_edges = int(np.count_nonzero(np.any(np.triu(alarming | alarming.T, 1), axis=1)))
        connectivity_matrix[too_far] = 0
        if n_alarming_ignored_edges > 0:
            logger.warning("  At least %i site pairs with high (z-score > 3) fluxes were over the given distance cutoff.\n"
                           "  This may or may not be a problem; but if `distance_threshold` is low, consider raising it."
                           % n_alarming_ignored_edges)

        mp = dict(self.markov_parameters)
        if self.iterlimit is not None:
            mp.setdefault('iterlimit', self.iterlimit)
        return markov_clustering(connectivity_matrix, **mp)
