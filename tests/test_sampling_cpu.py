"""Random patch sampling (midasmednet/dataset.py:18-88, 285-346): the oracle restatement and the product's host-side
position sampler against positions drawn by the reference's own function bodies (tests/golden/sampling.npz, generated
by oracle/make_golden.py with np.random.seed fixed), plus the properties every sample must have."""
import numpy as np
import pytest
import torch

from oracle import sampling as osamp
from mednet_b200.sampler import PatchPositionSampler, get_random_patch_indices


def test_oracle_reproduces_reference_positions(golden):
    g = golden("sampling")
    maps = osamp.label_any_maps(g["label"], len(g["probs"]))
    np.random.seed(int(g["seed"]))
    for want_ini, want_cls in zip(g["index_ini"], g["selected_class"]):
        ini, cls = osamp.sample_patch_position(g["label"], g["patch"], g["probs"], maps)
        assert cls == want_cls and np.array_equal(ini, want_ini)


@pytest.mark.parametrize("use_state", [False, True])
def test_product_sampler_reproduces_reference_positions(golden, use_state):
    """Same draws from the global NumPy state (what the reference uses) or from an explicit RandomState."""
    g = golden("sampling")
    rng = np.random.RandomState(int(g["seed"])) if use_state else None
    ps = PatchPositionSampler([torch.from_numpy(g["label"])], g["patch"], g["probs"], rng=rng)
    if not use_state:
        np.random.seed(int(g["seed"]))
    for i, (want_ini, want_cls) in enumerate(zip(g["index_ini"], g["selected_class"])):
        subject, ini, cls = ps(i)
        assert subject == 0 and cls == want_cls and np.array_equal(ini, want_ini)


def test_sampled_patches_lie_inside_and_contain_the_selected_class(golden):
    g = golden("sampling")
    label, patch = g["label"], g["patch"]
    ps = PatchPositionSampler([torch.from_numpy(label)], patch, g["probs"], rng=np.random.RandomState(5))
    seen = set()
    for i in range(200):
        _, ini, cls = ps(i)
        assert (ini >= 0).all() and (ini + patch <= np.asarray(label.shape)).all()
        if cls > 0:
            crop = label[ini[0]:ini[0] + patch[0], ini[1]:ini[1] + patch[1], ini[2]:ini[2] + patch[2]]
            assert (crop == cls).any()
        seen.add(cls)
    assert seen == {0, 1, 2}


def test_uniform_sampling_without_class_probabilities_and_subject_rotation():
    """class_probabilities=None -> one randint over the valid origins (dataset.py:80-86); idx % subjects (dataset.py:287)."""
    shapes = [(20, 18, 16), (12, 30, 14)]
    maps = [torch.zeros(s, dtype=torch.uint8) for s in shapes]
    ps = PatchPositionSampler(maps, [8, 8, 8], None, rng=np.random.RandomState(3))
    ref = np.random.RandomState(3)
    for i in range(10):
        subject, ini, cls = ps(i)
        assert subject == i % 2 and cls == 0
        want = ref.randint(low=np.zeros(3, dtype=int), high=np.asarray(shapes[subject]) - 8 + 1)
        assert np.array_equal(ini, want)


def test_absent_class_falls_back_to_uniform_position():
    """get_labeled_position returns None for a class that does not occur (dataset.py:50-51) -> unconstrained patch."""
    lab = torch.zeros((16, 16, 16), dtype=torch.uint8)
    ps = PatchPositionSampler([lab], [8, 8, 8], [0.0, 1.0], rng=np.random.RandomState(0))
    _, ini, cls = ps(0)
    assert cls == 1 and (ini >= 0).all() and (ini <= 8).all()
    o_ini, o_cls = None, None
    np.random.seed(0)
    o_ini, o_cls = osamp.sample_patch_position(lab.numpy(), [8, 8, 8], [0.0, 1.0])
    assert o_cls == 1 and np.array_equal(o_ini, PatchPositionSampler([lab], [8, 8, 8], [0.0, 1.0],
                                                                     rng=np.random.RandomState(0))(0)[1])


def test_patch_as_large_as_the_volume_and_oversize():
    ini, fin = get_random_patch_indices([8, 8, 8], [8, 8, 8], rng=np.random.RandomState(0))
    assert np.array_equal(ini, [0, 0, 0]) and np.array_equal(fin, [8, 8, 8])
    with pytest.raises(ValueError):
        PatchPositionSampler([torch.zeros((8, 8, 4), dtype=torch.uint8)], [8, 8, 8])


def _apply_coefficients(x, row):
    """NumPy evaluation of the coefficient contract in include/mednet_b200.h (mednet_intensity_aug_params)."""
    C = x.shape[0]
    v = x.astype(np.float32) + row[2:2 + C, None, None, None]
    if row[0] > 0:
        minm, rnge = v.min(), v.max() - v.min()
        v = np.power((v - minm) / np.float32(rnge + np.float32(1e-7)), row[0]) * rnge + minm
    if row[1]:
        for c in range(C):
            mn, lo, hi = v[c].mean(), v[c].min(), v[c].max()
            v[c] = np.clip((v[c] - mn) * row[2 + C + c] + mn, lo, hi)
    return v.astype(np.float32)


@pytest.mark.parametrize("channels", [1, 3])
def test_augmentation_draws_follow_the_oracle_chain(channels):
    """Host-drawn coefficients + the documented formula == the restated library chain run with the same seed, patch after
    patch (the draw count per patch is data independent, so the two generators stay in step)."""
    from oracle import augment as oaug
    from mednet_b200.sampler import IntensityAugmentation
    rs = np.random.RandomState(1)
    patches = [rs.randn(channels, 6, 7, 5).astype(np.float32) * 2 + 1 for _ in range(5)]
    np.random.seed(42)
    want = [oaug.augment_patch(p) for p in patches]
    aug = IntensityAugmentation(rng=np.random.RandomState(42))
    for p, w in zip(patches, want):
        row = aug.draw(channels)
        assert row.dtype == np.float32 and 0.7 <= row[0] <= 1.3 and row[1] == 1.0
        assert ((row[2 + channels:] >= 0.3) & (row[2 + channels:] <= 1.7)).all()
        np.testing.assert_allclose(_apply_coefficients(p, row), w, rtol=1e-5, atol=1e-5)


def test_augmentation_can_be_switched_off_per_step():
    from mednet_b200.sampler import IntensityAugmentation
    row = IntensityAugmentation(p_per_sample=0.0, rng=np.random.RandomState(0)).draw(2)
    assert row[0] == 0 and row[1] == 0 and (row[2:4] == 0).all()
    x = np.random.RandomState(0).randn(2, 4, 4, 4).astype(np.float32)
    assert np.array_equal(_apply_coefficients(x, row), x)


def test_reader_adapters_follow_the_reference_interface(tmp_path):
    """DataReader surface of dataset.py:109-207: storage dtypes on preload, lazy handles otherwise, a clear error for the
    container formats whose libraries are not installed."""
    from mednet_b200.dataset import DataReaderArrays, DataReaderHDF5, one_hot_to_label
    arrays = {"images/a": np.ones((1, 4, 5, 6), np.float32) * 1.0009765625, "labels/a": np.full((1, 4, 5, 6), 3, np.int64)}
    path = tmp_path / "cohort.npz"
    np.savez(path, **arrays)
    r = DataReaderArrays(str(path))
    img = r.read_data_to_memory(["a"], "images", dtype=np.float16)[0]
    lab = r.read_data_to_memory(["a"], "labels", dtype=np.uint8)[0]
    assert img.dtype == np.float16 and lab.dtype == np.uint8 and float(img.flat[0]) == 1.0009765625 and lab.max() == 3
    assert r.get_data_shape(["a"], "labels") == {"a": (1, 4, 5, 6)}
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            DataReaderHDF5(str(path))
    one_hot = np.zeros((2, 1, 1, 3))
    one_hot[0, 0, 0, 1] = one_hot[1, 0, 0, 2] = 1
    assert one_hot_to_label(one_hot).tolist() == [[[[0, 1, 2]]]]
