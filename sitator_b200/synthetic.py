"""Seeded synthetic systems shaped like BASELINE.json's configs.

The reference ships no data, no tests and needs Zeo++ (absent) to make a landmark
basis, so every workload here is generated: a static host lattice, a landmark
basis (centre + the static atoms that define it, what ``VoronoiSiteGenerator``
would return: reference ``voronoi.py:23-41``), a set of true sites, and a
trajectory in which mobile atoms vibrate about sites and hop between them.

Only numpy/scipy; shared by tests, bench.py and the golden-vector generator.
"""
import numpy as np

__all__ = ["SynthSystem", "toy_bcc", "blue_noise_cell", "CONFIGS", "make_config"]


def _min_image_diag(d, lengths):
    return d - lengths * np.round(d / lengths)


class SynthSystem(object):
    """A static lattice + landmark basis + true sites in an orthorhombic cell.

    Atom order in a frame: static atoms first, then mobile (``interleave=True``
    shuffles them, seeded, to exercise the masks).
    """

    def __init__(self, cell, static_pos, n_mobile, lm_centers, lm_vertices, site_pos, seed,
                 interleave=False, name="synthetic"):
        self.name = name
        self.cell = np.asarray(cell, dtype=np.float64)
        self.lengths = np.diag(self.cell).copy()
        self.static_pos = np.asarray(static_pos, dtype=np.float64)
        self.n_static = len(self.static_pos)
        self.n_mobile = int(n_mobile)
        self.n_total = self.n_static + self.n_mobile
        self.lm_centers = np.asarray(lm_centers, dtype=np.float64)
        self.lm_vertices = [list(map(int, v)) for v in lm_vertices]
        self.site_pos = np.asarray(site_pos, dtype=np.float64)
        self.seed = seed
        rng = np.random.default_rng(seed + 7919)
        order = np.arange(self.n_total)
        if interleave:
            rng.shuffle(order)
        # frame index of static atom s is static_idx[s]; of mobile j is mobile_idx[j];
        # both ascending so that frame[mask] enumerates them in lattice order.
        self.static_idx = np.sort(order[:self.n_static])
        self.mobile_idx = np.sort(order[self.n_static:])
        self.static_mask = np.zeros(self.n_total, dtype=bool)
        self.static_mask[self.static_idx] = True
        self.mobile_mask = ~self.static_mask
        # site neighbour lists (for hops)
        d = self.site_pos[:, None, :] - self.site_pos[None, :, :]
        d = _min_image_diag(d, self.lengths)
        self.site_dist = np.sqrt((d ** 2).sum(-1))

    @property
    def n_landmarks(self):
        return len(self.lm_centers)

    def initial_structure_positions(self):
        """Positions of the 'structure' (ideal statics + mobiles parked on the first sites)."""
        pos = np.zeros((self.n_total, 3))
        pos[self.static_idx] = self.static_pos
        pos[self.mobile_idx] = self.site_pos[:self.n_mobile]
        return pos

    def site_sequence(self, n_frames, hop_prob=0.01, hop_radius=4.5, seed=None):
        """(n_frames, n_mobile) int array of true site per atom; hops only to unoccupied sites."""
        rng = np.random.default_rng((self.seed if seed is None else seed) + 104729)
        n_sites = len(self.site_pos)
        M = self.n_mobile
        cur = rng.permutation(n_sites)[:M].copy()
        occupied = np.zeros(n_sites, dtype=bool)
        occupied[cur] = True
        hop_mask = rng.random((n_frames, M)) < hop_prob
        hop_mask[0] = False
        ev_f, ev_a = np.nonzero(hop_mask)
        seq = np.empty((n_frames, M), dtype=np.int64)
        change_frames = [[0] for _ in range(M)]
        change_sites = [[int(cur[a])] for a in range(M)]
        picks = rng.random(len(ev_f))
        for e in range(len(ev_f)):
            a = ev_a[e]
            s = cur[a]
            cand = np.nonzero((self.site_dist[s] < hop_radius) & (~occupied))[0]
            if len(cand) == 0:
                continue
            t = cand[int(picks[e] * len(cand))]
            occupied[s] = False
            occupied[t] = True
            cur[a] = t
            change_frames[a].append(int(ev_f[e]))
            change_sites[a].append(int(t))
        for a in range(M):
            cf = np.asarray(change_frames[a] + [n_frames])
            seq[:, a] = np.repeat(np.asarray(change_sites[a]), np.diff(cf))
        return seq

    def trajectory(self, n_frames, sigma_static=0.05, sigma_mobile=0.10, hop_prob=0.01,
                   transit_frames=1, unwrapped=True, swap_statics_at=None, seed=None,
                   return_sites=False):
        """Frames (n_frames, n_total, 3) float64.

        ``transit_frames``: on a hop the atom spends this many frames on the straight
        line between the two sites (produces unassigned frames, as real MD does).
        ``swap_statics_at``: frame from which the two mutually closest static atoms
        trade places (exercises ``dynamic_lattice_mapping``, helpers.pyx:60-64).
        """
        seed = self.seed if seed is None else seed
        rng = np.random.default_rng(seed + 15485863)
        seq = self.site_sequence(n_frames, hop_prob=hop_prob, seed=seed)
        M = self.n_mobile
        frames = np.empty((n_frames, self.n_total, 3), dtype=np.float64)
        # statics
        st = self.static_pos[None, :, :] + rng.normal(0.0, sigma_static, size=(n_frames, self.n_static, 3))
        if swap_statics_at is not None:
            d = self.static_pos[:, None, :] - self.static_pos[None, :, :]
            d = _min_image_diag(d, self.lengths)
            dist = np.sqrt((d ** 2).sum(-1)) + np.eye(self.n_static) * 1e9
            i, j = np.unravel_index(np.argmin(dist), dist.shape)
            tmp = st[swap_statics_at:, i].copy()
            st[swap_statics_at:, i] = st[swap_statics_at:, j]
            st[swap_statics_at:, j] = tmp
        frames[:, self.static_idx] = st
        del st
        # mobiles: site position (unwrapped along the hop path) + noise
        base = self.site_pos[seq]  # (F, M, 3)
        if unwrapped:
            step = np.zeros_like(base)
            step[1:] = _min_image_diag(base[1:] - base[:-1], self.lengths)
            step[0] = base[0]
            base = np.cumsum(step, axis=0)
        if transit_frames > 0:
            hopped = np.zeros((n_frames, M), dtype=bool)
            hopped[1:] = seq[1:] != seq[:-1]
            hf, ha = np.nonzero(hopped)
            for t in range(transit_frames):
                # frame hf + t sits at fraction (t+1)/(transit+1) of the way
                f = hf + t
                ok = f < n_frames
                # do not overwrite a later hop's frames
                frac = (t + 1.0) / (transit_frames + 1.0)
                prev = base[hf[ok] - 1, ha[ok]]
                nxt = base[np.minimum(hf[ok] + transit_frames, n_frames - 1), ha[ok]]
                same = seq[np.minimum(hf[ok] + transit_frames, n_frames - 1), ha[ok]] == seq[hf[ok], ha[ok]]
                val = prev + frac * (nxt - prev)
                fo, ao = f[ok][same], ha[ok][same]
                base[fo, ao] = val[same]
        frames[:, self.mobile_idx] = base + rng.normal(0.0, sigma_mobile, size=(n_frames, M, 3))
        if return_sites:
            return frames, seq
        return frames


# ----------------------------------------------------------------------------------------
# builders
# ----------------------------------------------------------------------------------------

def _greedy_sites(cands, clearance, lengths, n_sites, min_sep):
    order = np.argsort(-clearance)
    chosen = []
    for i in order:
        p = cands[i]
        ok = True
        for c in chosen:
            d = _min_image_diag(p - cands[c], lengths)
            if d @ d < min_sep * min_sep:
                ok = False
                break
        if ok:
            chosen.append(i)
            if len(chosen) == n_sites:
                break
    return np.asarray(chosen, dtype=int)


def toy_bcc(seed=0, n_cells=4, a=3.0, n_mobile=16, n_landmarks=200):
    """Config 1: 4x4x4 BCC host (128 atoms), tetrahedral interstitials as landmarks."""
    rng = np.random.default_rng(seed)
    L = n_cells * a
    cell = np.eye(3) * L
    lengths = np.array([L, L, L])
    grid = np.stack(np.meshgrid(*[np.arange(n_cells)] * 3, indexing="ij"), -1).reshape(-1, 3)
    static = np.concatenate([grid * a, (grid + 0.5) * a]).astype(np.float64)
    # the 12 tetrahedral interstitials per conventional cell: (1/2, 1/4, 0) and permutations
    tet = []
    for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
        for q in (0.25, 0.75):
            for z in (0.0,):
                v = np.zeros(3)
                v[perm[0]] = 0.5
                v[perm[1]] = q
                v[perm[2]] = z
                tet.append(v)
                v2 = np.zeros(3)
                v2[perm[0]] = q
                v2[perm[1]] = 0.5
                v2[perm[2]] = z
                tet.append(v2)
    tet = np.unique(np.round(np.asarray(tet), 6), axis=0)
    allt = (grid[:, None, :] + tet[None, :, :]).reshape(-1, 3) * a
    allt = np.mod(allt, L)
    pick = np.sort(rng.permutation(len(allt))[:n_landmarks])
    centers = allt[pick]
    verts = []
    clearance = np.empty(len(centers))
    for i, c in enumerate(centers):
        d = _min_image_diag(static - c, lengths)
        r = np.sqrt((d ** 2).sum(-1))
        nn = np.argsort(r, kind="stable")[:4]
        verts.append(sorted(int(x) for x in nn))
        clearance[i] = r[nn[0]]
    n_sites = int(round(1.75 * n_mobile))
    chosen = _greedy_sites(centers, clearance + rng.random(len(centers)) * 1e-3, lengths, n_sites, 2.2)
    sites = centers[chosen]
    return SynthSystem(cell, static, n_mobile, centers, verts, sites, seed, name="toy_bcc")


def _dart_throw(rng, lengths, n, min_sep, max_tries=2000000):
    pts = np.empty((n, 3))
    k = 0
    tries = 0
    while k < n:
        tries += 1
        if tries > max_tries:
            raise RuntimeError("dart throwing did not converge")
        p = rng.random(3) * lengths
        if k:
            d = _min_image_diag(pts[:k] - p, lengths)
            if np.min((d ** 2).sum(-1)) < min_sep * min_sep:
                continue
        pts[k] = p
        k += 1
    return pts


def blue_noise_cell(seed, lengths, n_static, n_mobile, n_landmarks=None, min_sep=2.0,
                    site_factor=1.7, site_sep=2.2, name="blue_noise", interleave=False):
    """Configs 2-5: blue-noise static atoms, periodic-Voronoi landmark basis.

    Landmarks are the 4-generator Voronoi nodes of the static lattice (what Zeo++'s
    ``-nt2`` output gives the reference, ``util/zeo.py:144-182``).  If ``n_landmarks``
    exceeds their number, circumcentres of Delaunay faces (3 generating atoms) are
    added, seeded, until the target is met -- this also produces ragged (-1 padded)
    vertex tables (LandmarkAnalysis.py:194-195).
    """
    from scipy.spatial import Voronoi
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.float64)
    cell = np.diag(lengths)
    static = _dart_throw(rng, lengths, n_static, min_sep)
    shifts = np.stack(np.meshgrid(*[np.array([0, -1, 1])] * 3, indexing="ij"), -1).reshape(-1, 3)
    rep = (static[None, :, :] + shifts[:, None, :] * lengths).reshape(-1, 3)  # image 0 first
    vor = Voronoi(rep)
    nv = len(vor.vertices)
    gens = [set() for _ in range(nv)]
    faces = {}
    for (p, q), rv in zip(vor.ridge_points, vor.ridge_vertices):
        for v in rv:
            if v >= 0:
                gens[v].add(int(p))
                gens[v].add(int(q))
    inside = np.all((vor.vertices >= 0.0) & (vor.vertices < lengths), axis=1)
    centers, verts, clear = [], [], []
    for v in np.nonzero(inside)[0]:
        g = sorted(gens[v])
        home = sorted(set(x % n_static for x in g))
        if len(g) != 4 or len(home) != 4:
            continue
        centers.append(vor.vertices[v])
        verts.append(home)
        d = _min_image_diag(static[home[0]] - vor.vertices[v], lengths)
        clear.append(np.sqrt(d @ d))
    centers = np.asarray(centers)
    clear = np.asarray(clear)
    n4 = len(centers)
    if n_landmarks is not None and n_landmarks < n4:
        keep = np.sort(rng.permutation(n4)[:n_landmarks])
        centers, clear = centers[keep], clear[keep]
        verts = [verts[i] for i in keep]
    elif n_landmarks is not None and n_landmarks > n4:
        # triangular Delaunay faces = triples of generators shared by two adjacent nodes
        tri = {}
        for v in np.nonzero(inside)[0]:
            g = sorted(gens[v])
            if len(g) != 4:
                continue
            for drop in range(4):
                t = tuple(g[:drop] + g[drop + 1:])
                tri.setdefault(t, None)
        extra_c, extra_v = [], []
        for t in tri:
            home = sorted(set(x % n_static for x in t))
            if len(home) != 3:
                continue
            a, b, c = rep[list(t)]
            # circumcentre of triangle abc
            ab, ac = b - a, c - a
            n = np.cross(ab, ac)
            nn = n @ n
            if nn < 1e-12:
                continue
            cc = a + (np.cross(n, ab) * (ac @ ac) + np.cross(ac, n) * (ab @ ab)) / (2.0 * nn)
            if not np.all((cc >= 0.0) & (cc < lengths)):
                continue
            extra_c.append(cc)
            extra_v.append(home)
        need = n_landmarks - n4
        if need > len(extra_c):
            raise RuntimeError("cannot reach %d landmarks (have %d + %d)" % (n_landmarks, n4, len(extra_c)))
        pick = np.sort(rng.permutation(len(extra_c))[:need])
        centers4 = centers
        centers = np.concatenate([centers, np.asarray(extra_c)[pick]])
        verts = verts + [extra_v[i] for i in pick]
        # keep landmark order mixed so ragged rows are not all at the end
        perm = rng.permutation(len(centers))
        centers = centers[perm]
        verts = [verts[i] for i in perm]
        clear_all = np.concatenate([clear, np.zeros(need)])[perm]
        clear = clear_all
        del centers4
    n_sites = int(round(site_factor * n_mobile))
    is4 = np.asarray([len(v) == 4 for v in verts])
    cand = np.nonzero(is4)[0]
    chosen = _greedy_sites(centers[cand], clear[cand], lengths, n_sites, site_sep)
    if len(chosen) < n_mobile + 2:
        raise RuntimeError("only %d sites found" % len(chosen))
    sites = centers[cand][chosen]
    return SynthSystem(cell, static, n_mobile, centers, verts, sites, seed, interleave=interleave, name=name)


# BASELINE.json configs (seeds 0-4 = configs 1-5).  At these mobile densities
# (56 Li in 2182 A^3) true sites sit ~2.2 A apart and MCL merges some neighbours into one site, so the
# dense configs run with max_mobile_per_site=2, as the reference's docstring advises for such
# systems (LandmarkAnalysis.py:69-78).
CONFIGS = {
    # name: (builder kwargs, default n_frames, analysis kwargs)
    # the toy keeps ~200 of 768 interstitials as landmarks, so an atom in mid-hop can see none:
    # zero landmark vectors are counted instead of raised (LandmarkAnalysis.py:224-225)
    "toy_bcc": dict(kind="bcc", seed=0, n_frames=2000, dynamic=False, check_for_zero_landmarks=False),
    "llzo": dict(kind="blue", seed=1, lengths=(12.97, 12.97, 12.97), n_static=136, n_mobile=56,
                 n_landmarks=1500, n_frames=100000, dynamic=False, max_mobile_per_site=2),
    "llzo_v4": dict(kind="blue", seed=1, lengths=(12.97, 12.97, 12.97), n_static=136, n_mobile=56,
                    n_landmarks=None, n_frames=100000, dynamic=False, max_mobile_per_site=2),
    "lgps_dynamic": dict(kind="blue", seed=2, lengths=(17.4, 17.4, 12.6), n_static=120, n_mobile=80,
                         n_landmarks=None, n_frames=100000, dynamic=True, min_sep=2.6, max_mobile_per_site=2),
    "laso": dict(kind="blue", seed=3, lengths=(27.0, 27.0, 27.0), n_static=1400, n_mobile=200,
                 n_landmarks=3000, n_frames=500000, dynamic=False, min_sep=2.0, max_mobile_per_site=2),
    "llzo_sweep": dict(kind="blue", seed=1, lengths=(12.97, 12.97, 12.97), n_static=136, n_mobile=56,
                       n_landmarks=1500, n_frames=1000000, dynamic=False, max_mobile_per_site=2),
}


def make_config(name, **over):
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    if cfg["kind"] == "bcc":
        return toy_bcc(seed=cfg["seed"]), cfg
    sysm = blue_noise_cell(cfg["seed"], cfg["lengths"], cfg["n_static"], cfg["n_mobile"],
                           n_landmarks=cfg.get("n_landmarks"), min_sep=cfg.get("min_sep", 2.0),
                           name=name, interleave=cfg.get("interleave", False))
    return sysm, cfg


def site_network_for(system, site_network_cls=None, atoms_cls=None):
    """Landmark-basis ``SiteNetwork`` for a :class:`SynthSystem`.

    ``site_network_cls``/``atoms_cls`` default to this package's own; tests pass the compiled
    reference's classes to feed it the identical input.
    """
    if site_network_cls is None:
        from .SiteNetwork import SiteNetwork as site_network_cls
    if atoms_cls is None:
        from .structure import Atoms as atoms_cls
    numbers = np.where(system.static_mask, 8, 3)
    try:
        atoms = atoms_cls(positions=system.initial_structure_positions(), cell=system.cell, numbers=numbers)
    except TypeError:
        atoms = atoms_cls(system.initial_structure_positions(), system.cell, numbers)
    sn = site_network_cls(atoms, system.static_mask.copy(), system.mobile_mask.copy())
    sn.centers = system.lm_centers.copy()
    verts = np.empty(system.n_landmarks, dtype=object)   # object rows: ragged-safe in the reference too
    for i, v in enumerate(system.lm_vertices):
        verts[i] = list(v)
    sn.vertices = verts
    return sn
