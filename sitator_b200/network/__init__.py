from .merging import MergeSites, MergeSitesError, MergedSitesTooDistantError  # noqa: F401
