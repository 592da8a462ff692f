"""CPU tests for the batched GTA driver's host logic (tacotron2_subword_b200/gta.py): batch planning and the .npy writer."""
import numpy as np
import pytest

from tacotron2_subword_b200.gta import AsyncNpyWriter, plan_batches


def test_plan_batches_groups_similar_lengths_and_covers_everything():
    rng = np.random.default_rng(0)
    n_frames = rng.integers(50, 900, size=137).tolist()
    batches = plan_batches(n_frames, max_batch=16)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(137))                        # every utterance exactly once
    assert all(1 <= len(b) <= 16 for b in batches)
    firsts = [n_frames[b[0]] for b in batches]
    assert firsts == sorted(firsts, reverse=True)                  # longest first (data_utils.py:146-160 sorts the same way)
    for b in batches:
        assert n_frames[b[0]] == max(n_frames[i] for i in b)
    waste = sum(len(b) * n_frames[b[0]] - sum(n_frames[i] for i in b) for b in batches) / sum(n_frames)
    assert waste < 0.15                                            # padding overhead of length bucketing


def test_plan_batches_frame_budget_and_edges():
    assert plan_batches([], 8) == []
    assert plan_batches([7], 8) == [[0]]
    batches = plan_batches([800, 790, 780, 100, 90, 80, 70, 60], max_batch=8, max_frames_per_batch=2400)
    assert batches[0] == [0, 1, 2] and all(len(b) * 800 <= 2400 or b is not batches[0] for b in batches)
    assert sorted(i for b in batches for i in b) == list(range(8))
    with pytest.raises(ValueError):
        plan_batches([1, 2], 0)


def test_async_writer_writes_reference_format(tmp_path):
    w = AsyncNpyWriter(str(tmp_path))
    arrays = {f"utt{i}": np.random.default_rng(i).standard_normal((1, 80, 10 + i)).astype(np.float32) for i in range(20)}
    for name, a in arrays.items():
        w.put(name, a)
    files = w.close()
    assert len(files) == 20
    for name, a in arrays.items():
        got = np.load(tmp_path / (name + ".npy"))                   # GTA.py:61: np.save(folder + "/" + file_name, mel[1,80,T])
        assert got.dtype == np.float32 and got.shape == a.shape and np.array_equal(got, a)


def test_async_writer_reports_io_errors(tmp_path):
    w = AsyncNpyWriter(str(tmp_path))
    w.put("missing_dir/utt", np.zeros((1, 80, 3), np.float32))       # parent directory does not exist
    with pytest.raises(OSError):
        w.close()
