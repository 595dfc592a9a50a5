"""Trajectories delivered in pieces.

The reference's ``LandmarkAnalysis.run`` takes one (n_frames, n_atoms, 3) float64 ndarray that must sit in host
memory as a whole (``LandmarkAnalysis.py:148,182-183``).  At BASELINE config 5 that is a 4.6 GB host array that
exists only to be copied to the device once.  :class:`ChunkedFrames` lets ``run`` take the trajectory as a stream
of frame blocks instead -- slices of a memory-mapped ``.npy`` file, or whatever an MD reader yields -- and
assembles the resident float64 frame array directly in HBM through two page-locked staging buffers, so the host
never holds more than two blocks.  Everything after the upload is the normal path (the engine borrows the device
array), so results are those of ``run`` on the concatenated ndarray.

There is no host-side copy of the trajectory afterwards: ``SiteTrajectory.real_trajectory`` is the
:class:`ChunkedFrames` object itself, which can be iterated again or sliced with ``[...]`` if its source allows.
"""
import numpy as np


class ChunkedFrames(object):
    """A trajectory as an iterable of (f_i, n_atoms, 3) float64 / float32 blocks.

    Args:
        chunks: a callable returning a fresh iterator over the blocks (preferred: the trajectory can then be read
            again), or a one-shot iterable.
        n_frames (int): total number of frames the blocks add up to (checked while reading).
        n_atoms (int): atoms per frame.
    """

    def __init__(self, chunks, n_frames, n_atoms):
        self._chunks = chunks
        self.n_frames = int(n_frames)
        self.n_atoms = int(n_atoms)
        if self.n_frames <= 0 or self.n_atoms <= 0:
            raise ValueError("ChunkedFrames needs n_frames > 0 and n_atoms > 0")
        self._array = None        # set by from_array / from_npy: supports slicing

    # -- the little of the ndarray surface that run() and SiteTrajectory use
    @property
    def shape(self):
        return (self.n_frames, self.n_atoms, 3)

    def __len__(self):
        return self.n_frames

    def __iter__(self):
        return iter(self._chunks() if callable(self._chunks) else self._chunks)

    def __getitem__(self, key):
        if self._array is None:
            raise TypeError("this ChunkedFrames has no random access (built from an iterator)")
        return self._array[key]

    # -- constructors
    @classmethod
    def from_array(cls, array, chunk_frames=4096):
        """Blocks of ``chunk_frames`` frames of anything sliceable along its first axis (ndarray, ``np.memmap``,
        an HDF5 dataset ...)."""
        if len(array.shape) != 3 or array.shape[2] != 3:
            raise ValueError("Wrong shape %s for frames." % (tuple(array.shape),))
        chunk_frames = max(1, int(chunk_frames))
        n = int(array.shape[0])

        def blocks():
            for f0 in range(0, n, chunk_frames):
                yield array[f0:min(n, f0 + chunk_frames)]
        out = cls(blocks, n, int(array.shape[1]))
        out._array = array
        return out

    @classmethod
    def from_npy(cls, path, chunk_frames=4096):
        """A ``.npy`` file read through ``np.load(mmap_mode='r')``: only the block in flight is paged in."""
        return cls.from_array(np.load(path, mmap_mode='r'), chunk_frames)

    # -- the upload
    def to_device(self, device):
        """The whole trajectory as one contiguous float64 CUDA tensor (n_frames, n_atoms, 3).

        Block i is copied into page-locked buffer i % 2 by the host while block i - 1 crosses PCIe; float32 blocks
        cross at half the size and are widened on the device.  Raises if there is no CUDA device (no CPU path)."""
        import torch
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ChunkedFrames.to_device needs a CUDA device")
        A = self.n_atoms
        dev = torch.empty((self.n_frames, A, 3), dtype=torch.float64, device=device)
        stage = [None, None]                  # pinned host tensors, grown to the largest block seen
        done = [None, None]                   # event: the copy out of stage[i] has finished
        filled = 0
        with torch.cuda.device(device):
            for i, block in enumerate(self):
                block = np.asarray(block)
                if block.ndim != 3 or block.shape[1:] != (A, 3):
                    raise ValueError("Wrong shape %s for a block of frames." % (block.shape,))
                if block.dtype not in (np.float64, np.float32):
                    raise ValueError("frames must be float64 (as in the reference) or float32")
                f = int(block.shape[0])
                if f == 0:
                    continue
                if filled + f > self.n_frames:
                    raise ValueError("the blocks hold more than the announced %d frames" % self.n_frames)
                b = i & 1
                tdtype = torch.float64 if block.dtype == np.float64 else torch.float32
                if done[b] is not None:
                    done[b].synchronize()     # the buffer's previous block is on the device
                if stage[b] is None or stage[b].dtype != tdtype or stage[b].shape[0] < f:
                    stage[b] = torch.empty((f, A, 3), dtype=tdtype, pin_memory=True)
                host = stage[b][:f]
                np.copyto(host.numpy(), block)
                if tdtype == torch.float64:
                    dev[filled:filled + f].copy_(host, non_blocking=True)
                else:                         # float32 crosses PCIe as it is and is widened on the device
                    dev[filled:filled + f].copy_(host.to(device, non_blocking=True))
                if done[b] is None:
                    done[b] = torch.cuda.Event()
                done[b].record()
                filled += f
            torch.cuda.synchronize(device)
        if filled != self.n_frames:
            raise ValueError("the blocks hold %d frames, %d were announced" % (filled, self.n_frames))
        return dev
