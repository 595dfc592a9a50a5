"""The clustering plugins' view of the landmark vectors.

The reference hands its clustering plugins a dense (N, L) float64 matrix, memory-mapped from a
temp file (``LandmarkAnalysis.py:209-218``, 67 GB at BASELINE config 2).  Here the landmark
vectors are never stored dense: a plugin receives a :class:`LandmarkVectorSource`, which can
stream reductions over them (each one a pass of the fused kernel over the resident frames) and
materialise selected rows.  In a frame-sharded run every rank holds one source over its shard and
the reductions are all-reduced.
"""
import numpy as np


class LandmarkVectorSource(object):
    def __init__(self, engine, comm=None):
        self.engine = engine
        self.comm = comm                      # landmark.parallel.Comm or None
        self.n_local = engine.n_frames * engine.M
        self.n_total = self.n_local if comm is None else int(comm.allreduce_sum_scalar(self.n_local))
        self.row0 = engine.frame0 * engine.M  # global index of the first local row
        self.seen = None
        self.gram_upper = None

    @property
    def shape(self):
        return (self.n_total, self.engine.L)

    def __len__(self):
        return self.n_total

    def rows(self, global_rows):
        """float64 landmark vectors of the given global rows (each must be resident on some rank)."""
        import torch
        eng = self.engine
        global_rows = np.asarray(global_rows, dtype=np.int64)
        out = np.zeros((len(global_rows), eng.L), dtype=np.float64)
        local = (global_rows >= self.row0) & (global_rows < self.row0 + self.n_local)
        if np.any(local) and self.sparse is not None:
            # straight from the compressed cache: row_ptr -> (landmark, value) entries
            lr = torch.as_tensor(global_rows[local] - self.row0, device=eng.device)
            ptr = self.sparse.ptr.index_select(0, lr)
            cnt = (ptr & 0xFF).cpu().numpy()
            off = (ptr >> 8).cpu().numpy()
            idx = np.concatenate([np.arange(o, o + c, dtype=np.int64) for o, c in zip(off, cnt)]) if len(off) else np.zeros(0, np.int64)
            if len(idx):
                idx_d = torch.as_tensor(idx, device=eng.device)
                kk = self.sparse.k.index_select(0, idx_d).cpu().numpy().view(np.uint16).astype(np.int64)
                vv = self.sparse.v.index_select(0, idx_d).cpu().numpy()
                sub = np.zeros((int(local.sum()), eng.L), dtype=np.float64)
                sub[np.repeat(np.arange(len(cnt)), cnt), kk] = vv
                out[local] = sub
        elif np.any(local):
            lr = global_rows[local] - self.row0
            frames = np.unique(lr // eng.M)
            dense = eng.fill_frames(frames, dtype=torch.float64)            # (len(frames) * M, L) on the device
            pos = np.searchsorted(frames, lr // eng.M) * eng.M + lr % eng.M
            sel = dense.index_select(0, torch.as_tensor(pos, device=dense.device))
            out[local] = sel.cpu().numpy()
        if self.comm is not None:
            out = self.comm.allreduce_sum_numpy(out)
        return out

    def row_norms2(self, global_rows):
        """|landmark vector|^2 of the given global rows (float64 numpy), each resident on some rank."""
        import torch
        eng = self.engine
        global_rows = np.asarray(global_rows, dtype=np.int64)
        if self.sparse is None:
            lv = self.rows(global_rows)
            return np.einsum('ij,ij->i', lv, lv)
        local = torch.as_tensor(global_rows - self.row0, device=eng.device)
        out = torch.empty((len(global_rows),), dtype=torch.float64, device=eng.device)
        from .. import _native
        _native.check(eng._lib.sitb_sparse_row_norm2(eng._ctx, eng._ptr(self.sparse.ptr), eng._ptr(self.sparse.v),
                                                     self.sparse.n_rows, eng._ptr(local), len(global_rows), eng._ptr(out)))
        if self.comm is not None:
            self.comm.allreduce_sum_(out)
        return out.cpu().numpy()

    # set by the plugin's first pass when the rows were cached (engine.SparseRows)
    sparse = None
    cache_rows = True

    def assign(self, threshold, **outputs):
        """One assign pass over all local landmark vectors: from the cached compressed rows if present,
        otherwise by rerunning the fused fill + assign kernel over the resident frames."""
        if self.sparse is not None:
            self.engine.assign_sparse(self.sparse, threshold, **outputs)
        else:
            self.engine.pass_assign(threshold, **outputs)
