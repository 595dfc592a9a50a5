"""Exception types of the site-analysis boundary (mirrors reference ``sitator/errors.py:2-21``:
same class names, constructor arguments, attributes and messages)."""


class SiteAnaysisError(Exception):
    """Base class (the reference spells it this way, ``errors.py:2``)."""


SiteAnalysisError = SiteAnaysisError


class MultipleOccupancyError(SiteAnaysisError):
    """More mobile atoms than allowed share one site in one frame (``errors.py:6-14``)."""

    def __init__(self, mobile, site, frame):
        super().__init__("Multiple mobile particles %s were assigned to site %i at frame %i."
                         % (mobile, site, frame))
        self.mobile_particles = mobile
        self.site = site
        self.frame = frame


class InsufficientSitesError(SiteAnaysisError):
    """Fewer sites than mobile particles (``errors.py:16-21``)."""

    def __init__(self, verb, n_sites, n_mobile):
        super().__init__("%s resulted in only %i sites for %i mobile particles."
                         % (verb, n_sites, n_mobile))
        self.n_sites = n_sites
        self.n_mobile = n_mobile
