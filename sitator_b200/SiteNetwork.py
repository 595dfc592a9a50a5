"""``SiteNetwork``: sites (centres, defining static atoms, types, attributes) in a host lattice.

Host-side data contract of the landmark-analysis path -- the input (landmark basis) and
output (sites) container of ``LandmarkAnalysis.run``.  Mirrors the public surface of the
reference's ``sitator/SiteNetwork.py:14-391`` (same names, argument meaning and errors);
plotting and ASE export are out of scope.  No kernels here.
"""
import re

import numpy as np

_ATTR_NAME = re.compile(r"^[a-zA-Z][a-zA-Z0-9_]*$")


def _as_object_rows(rows):
    """Ragged-safe 1-D object array of the given rows (ref ``SiteNetwork.py:110`` breaks on ragged)."""
    out = np.empty(len(rows), dtype=object)
    for i, r in enumerate(rows):
        out[i] = r
    return out


class SiteNetwork(object):
    """A network of mobile-particle sites in a static lattice (ref ``SiteNetwork.py:14-76``).

    Args:
        structure: ``ase.Atoms`` or :class:`sitator_b200.structure.Atoms` with every atom of the MD cell.
        static_mask (bool ndarray): atoms of the host lattice.
        mobile_mask (bool ndarray): atoms whose motion is analysed.
    """

    ATTR_NAME_REGEX = _ATTR_NAME

    def __init__(self, structure, static_mask, mobile_mask):
        static_mask = np.asarray(static_mask)
        mobile_mask = np.asarray(mobile_mask)
        assert static_mask.ndim == mobile_mask.ndim == 1, "The masks must be one-dimensional"
        assert len(structure) == len(static_mask) == len(mobile_mask), \
            "The masks must have the same length as the # of atoms in the strucutre."
        assert not np.any(static_mask & mobile_mask), "static_mask and mobile_mask cannot overlap."

        self.structure = structure
        self.static_mask = static_mask
        self.mobile_mask = mobile_mask
        self.n_static = int(np.sum(static_mask))
        self.n_mobile = int(np.sum(mobile_mask))

        self.static_structure = structure.copy()
        del self.static_structure[(~static_mask) | mobile_mask]
        assert len(self.static_structure) == self.n_static

        self._centers = None
        self._vertices = None
        self._types = None
        self._site_attrs = {}
        self._edge_attrs = {}
        self._attr_computed = {}

    # -- copying / slicing (ref :78-125) ---------------------------------------------------
    def copy(self, with_computed=True):
        sn = self[np.ones(self.n_sites, dtype=bool)]
        if not with_computed:
            sn.clear_computed_attributes()
        return sn

    def __len__(self):
        return self.n_sites

    def __getitem__(self, key):
        sn = self.__new__(type(self))
        SiteNetwork.__init__(sn, self.structure, self.static_mask, self.mobile_mask)
        if self._centers is not None:
            sn.centers = self._centers[key]
        if self._vertices is not None:
            sn.vertices = _as_object_rows(list(self._vertices))[key].tolist()
        if self._types is not None:
            sn.site_types = self._types[key]
        for name, val in self._site_attrs.items():
            sn.add_site_attribute(name, val[key], computed=self._attr_computed.get(name, True))
        for name, mat in self._edge_attrs.items():
            sn.add_edge_attribute(name, mat[key][:, key], computed=self._attr_computed.get(name, True))
        return sn

    def of_type(self, stype):
        if self._types is None:
            raise ValueError("This SiteNetwork has no type information.")
        if stype not in self._types:
            raise ValueError("This SiteNetwork has no sites of type %i" % stype)
        return self[self._types == stype]

    # -- basic properties (ref :167-256) ---------------------------------------------------
    @property
    def n_sites(self):
        return 0 if self._centers is None else len(self._centers)

    @property
    def n_total(self):
        return len(self.static_mask)

    @property
    def centers(self):
        view = self._centers.view()
        view.flags.writeable = False
        return view

    @centers.setter
    def centers(self, value):
        value = np.asarray(value)
        if value.ndim != 2 or value.shape[1] != 3:
            raise ValueError("`centers` must be a list of points")
        # new centres invalidate everything derived from the old ones (ref :190-197)
        self._vertices = None
        self._types = None
        self._site_attrs = {}
        self._edge_attrs = {}
        self._attr_computed = {}
        self._centers = value

    def update_centers(self, newcenters):
        if newcenters.shape != self._centers.shape:
            raise ValueError("New `centers` must have same shape as old; try using the setter `.centers = ...`")
        self._centers = newcenters

    @property
    def vertices(self):
        return self._vertices

    @vertices.setter
    def vertices(self, value):
        if len(value) != len(self._centers):
            raise ValueError("Wrong # of vertices %i; expected %i" % (len(value), len(self._centers)))
        self._vertices = value

    @property
    def site_ids(self):
        return np.arange(self.n_sites)

    @property
    def number_of_vertices(self):
        return None if self._vertices is None else [len(v) for v in self._vertices]

    @property
    def site_types(self):
        if self._types is None:
            return None
        view = self._types.view()
        view.flags.writeable = False
        return view

    @site_types.setter
    def site_types(self, value):
        value = np.asarray(value)
        if value.shape != (len(self._centers),):
            raise ValueError("Wrong # of types %s; expected %i" % (value.shape, len(self._centers)))
        self._types = value

    @property
    def n_types(self):
        return len(np.unique(self.site_types))

    @property
    def types(self):
        return np.unique(self.site_types)

    # -- attributes (ref :258-386) ---------------------------------------------------------
    @property
    def site_attributes(self):
        return list(self._site_attrs.keys())

    @property
    def edge_attributes(self):
        return list(self._edge_attrs.keys())

    def has_attribute(self, attr):
        return attr in self._site_attrs or attr in self._edge_attrs

    def remove_attribute(self, attr):
        if attr in self._site_attrs:
            del self._site_attrs[attr]
        elif attr in self._edge_attrs:
            del self._edge_attrs[attr]
        else:
            raise AttributeError("This SiteNetwork has no site or edge attribute `%s`" % attr)
        self._attr_computed.pop(attr, None)

    def clear_attributes(self):
        self._site_attrs = {}
        self._edge_attrs = {}
        self._attr_computed = {}

    def clear_computed_attributes(self):
        for name in [k for k, c in self._attr_computed.items() if c]:
            self.remove_attribute(name)

    def __getattr__(self, name):
        d = self.__dict__
        if "_site_attrs" in d and name in d["_site_attrs"]:
            return d["_site_attrs"][name]
        if "_edge_attrs" in d and name in d["_edge_attrs"]:
            return d["_edge_attrs"][name]
        raise AttributeError("This SiteNetwork has no site or edge attribute `%s`" % name)

    def get_site(self, site):
        out = {"center": self.centers[site]}
        if self._vertices is not None:
            out["vertices"] = self._vertices[site]
        if self._types is not None:
            out["type"] = self._types[site]
        for name, val in self._site_attrs.items():
            out[name] = val[site]
        return out

    def get_edge(self, edge):
        if not self._edge_attrs:
            raise ValueError("This SiteNetwork has no edge attributes")
        return {name: mat[edge] for name, mat in self._edge_attrs.items()}

    def add_site_attribute(self, name, attr, computed=True):
        self._check_name(name)
        attr = np.asarray(attr)
        if attr.shape[0] != self.n_sites:
            raise ValueError("Attribute array has only %i entries; need one for all %i sites."
                             % (len(attr), self.n_sites))
        self._site_attrs[name] = attr
        self._attr_computed[name] = computed

    def add_edge_attribute(self, name, attr, computed=True):
        self._check_name(name)
        attr = np.asarray(attr)
        if not (attr.ndim >= 2 and attr.shape[0] == attr.shape[1] == self.n_sites):
            raise ValueError("Attribute matrix has shape %s; need first two dimensions to be %i"
                             % (attr.shape, self.n_sites))
        self._edge_attrs[name] = attr
        self._attr_computed[name] = computed

    def _check_name(self, name):
        if not _ATTR_NAME.match(name):
            raise ValueError("Attribute name `%s` invalid; must begin with a letter and contain only "
                             "letters, numbers, and underscores." % name)
        if name in self._edge_attrs or name in self._site_attrs:
            raise KeyError("Attribute with name `%s` already exists" % name)
        if name in self.__dict__:
            raise ValueError("Attribute name `%s` reserved." % name)

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the scope of sitator_b200 (SURVEY.md section 2, #21)")
