"""GPU: a trajectory handed to run() in blocks (ChunkedFrames) gives the results of the same
trajectory handed over as one ndarray."""
import numpy as np
import pytest

from sitator_b200 import synthetic as syn
from tests import _util as U

pytestmark = pytest.mark.gpu


def _run(system, cfg, frames):
    from sitator_b200.landmark import LandmarkAnalysis
    la = LandmarkAnalysis(clustering_algorithm='mcl', verbose=False, **U.analysis_kwargs(cfg))
    return la, la.run(syn.site_network_for(system), frames)


def _same(st, want):
    assert np.array_equal(st.traj, want.traj)
    # two runs accumulate the Gram and the site sums in different atomic orders: values agree to rounding
    assert np.max(np.abs(st.confidences - want.confidences)) < U.CONF_ATOL
    assert np.max(np.abs(np.asarray(st.site_network.centers) - np.asarray(want.site_network.centers))) < U.CENTER_ATOL
    assert np.array_equal(st.jump_array(), want.jump_array())


def test_chunked_frames_equal_whole_array(tmp_path):
    from sitator_b200.landmark import ChunkedFrames
    g, system, cfg, frames = U.load_golden("toy_bcc_300")
    _, want = _run(system, cfg, frames)
    assert np.array_equal(want.traj, g["labels"])

    # even blocks of a sliceable array
    src = ChunkedFrames.from_array(frames, chunk_frames=37)
    la, st = _run(system, cfg, src)
    _same(st, want)
    assert st.real_trajectory is src
    assert la.n_all_zero_lvecs == int(g["n_all_zero_lvecs"])

    # a one-shot generator of ragged blocks (an empty one included)
    cuts = [0, 1, 1, 130, 131, 300]
    gen = (frames[a:b] for a, b in zip(cuts[:-1], cuts[1:]))
    _, st = _run(system, cfg, ChunkedFrames(gen, len(frames), frames.shape[1]))
    _same(st, want)

    # float32 blocks out of a memory-mapped .npy file == the float32 ndarray
    f32 = frames.astype(np.float32)
    np.save(tmp_path / "traj.npy", f32)
    _, want32 = _run(system, cfg, f32)
    _, st = _run(system, cfg, ChunkedFrames.from_npy(tmp_path / "traj.npy", chunk_frames=64))
    _same(st, want32)


def test_chunked_frames_wrong_total_raises():
    from sitator_b200.landmark import ChunkedFrames
    g, system, cfg, frames = U.load_golden("toy_bcc_300")
    with pytest.raises(ValueError):
        _run(system, cfg, ChunkedFrames(lambda: iter([frames[:100]]), len(frames), frames.shape[1]))
