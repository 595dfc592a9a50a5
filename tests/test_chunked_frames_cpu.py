"""Host logic of ChunkedFrames (no device): block iteration, the ndarray surface run() uses, no CPU upload path."""
import numpy as np
import pytest

from sitator_b200.landmark import ChunkedFrames


def test_blocks_cover_the_array_and_can_be_read_again(tmp_path):
    a = np.random.default_rng(0).normal(size=(23, 5, 3))
    c = ChunkedFrames.from_array(a, chunk_frames=4)
    assert c.shape == a.shape and len(c) == 23
    for _ in range(2):
        blocks = list(c)
        assert [len(b) for b in blocks] == [4, 4, 4, 4, 4, 3]
        assert np.array_equal(np.concatenate(blocks), a)
    assert np.array_equal(c[3:7], a[3:7])
    np.save(tmp_path / "t.npy", a.astype(np.float32))
    m = ChunkedFrames.from_npy(tmp_path / "t.npy", chunk_frames=10)
    assert m.shape == a.shape
    assert np.array_equal(np.concatenate(list(m)), a.astype(np.float32))


def test_iterator_source_has_no_random_access():
    a = np.zeros((6, 2, 3))
    c = ChunkedFrames(iter([a[:2], a[2:]]), 6, 2)
    with pytest.raises(TypeError):
        c[0]
    assert sum(len(b) for b in c) == 6


def test_bad_arguments():
    with pytest.raises(ValueError):
        ChunkedFrames(iter([]), 0, 3)
    with pytest.raises(ValueError):
        ChunkedFrames.from_array(np.zeros((4, 3)))


def test_no_cpu_upload_path():
    c = ChunkedFrames.from_array(np.zeros((4, 2, 3)))
    with pytest.raises(RuntimeError):
        c.to_device("cpu")
