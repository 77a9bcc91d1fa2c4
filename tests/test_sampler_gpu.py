"""GPU-resident patch sampler: device crops bit-exact against the oracle's NumPy crop (midasmednet/dataset.py:315-336)
at the positions the host sampler drew, and the batch feeds the training step unchanged."""
import numpy as np
import pytest
import torch

from oracle import sampling as osamp

pytestmark = pytest.mark.gpu


def _cohort(seed=0, heatmaps=True):
    rs = np.random.RandomState(seed)
    shapes = [(40, 37, 29), (33, 48, 36), (64, 32, 40)]
    images = [rs.randn(2, *s).astype(np.float32) for s in shapes]
    labels = [(rs.rand(1, *s) > 0.8).astype(np.uint8) * rs.randint(1, 3, size=(1,) + s).astype(np.uint8) for s in shapes]
    hms = [rs.randint(0, 256, size=(3,) + s).astype(np.uint8) for s in shapes] if heatmaps else None
    return images, labels, hms


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("heatmaps", [False, True])
def test_batch_matches_oracle_crops(dtype, heatmaps):
    from mednet_b200.sampler import GpuMedDataset
    images, labels, hms = _cohort(1, heatmaps)
    P = [16, 24, 8]
    ds = GpuMedDataset(images, labels, samples_per_subject=4, patch_size=P, heatmaps=hms,
                       class_probabilities=[0.2, 0.5, 0.3], data_dtype=dtype, rng=np.random.RandomState(11))
    assert len(ds) == 12
    batch = ds.batch(range(7))
    data, label = batch["data"], batch["label"]
    assert data.shape == (7, 2, *P) and data.dtype == dtype and data.permute(0, 2, 3, 4, 1).is_contiguous()
    assert label.shape == (7, (3 if heatmaps else 0) + 1, *P) and label.dtype == torch.uint8
    for b in range(7):
        s = b % 3
        assert batch["subject_key"][b] == str(s)
        full = np.concatenate([hms[s], labels[s]], axis=0) if heatmaps else labels[s]
        want_d, want_l = osamp.crop_patch(images[s], full, batch["patch_position"][b], P)
        want_d = torch.from_numpy(want_d).to(dtype)                      # bf16: one round-to-nearest-even, as the kernel
        assert torch.equal(data[b].cpu(), want_d)
        assert np.array_equal(label[b].cpu().numpy(), want_l)
        cls = batch["selected_class"][b]
        if cls > 0:
            assert (want_l[-1] == cls).any()


def test_positions_equal_the_host_oracle_sequence():
    """The product draws from np.random in the reference's order: replaying the oracle with the same seed gives the
    same positions, for several subjects of different shape."""
    from mednet_b200.sampler import GpuMedDataset
    images, labels, _ = _cohort(2, False)
    P, probs = [12, 12, 12], [0.3, 0.3, 0.4]
    ds = GpuMedDataset(images, labels, 5, P, class_probabilities=probs, rng=np.random.RandomState(99))
    got = [ds.sample_position(i) for i in range(15)]
    np.random.seed(99)
    for i, (subject, ini, cls) in enumerate(got):
        lab = labels[i % 3][0]
        o_ini, o_cls = osamp.sample_patch_position(lab, P, probs, osamp.label_any_maps(lab, 3))
        assert subject == i % 3 and cls == o_cls and np.array_equal(ini, o_ini)


def test_getitem_contract_and_loader_epoch():
    from mednet_b200.sampler import GpuMedDataset
    images, labels, hms = _cohort(3, True)
    ds = GpuMedDataset(images, labels, 2, [8, 8, 8], heatmaps=hms, rng=np.random.RandomState(0))
    item = ds[4]
    assert item["data"].shape == (2, 8, 8, 8) and item["label"].shape == (4, 8, 8, 8)
    assert item["subject_key"] == "1" and item["selected_class"] == 0 and item["patch_position"].shape == (3,)
    sizes = [b["data"].shape[0] for b in ds.loader(4, shuffle=True)]
    assert sizes == [4, 2]
    assert [b["data"].shape[0] for b in ds.loader(4, shuffle=False, drop_last=True)] == [4]


def test_sampler_feeds_training_step():
    from mednet_b200.sampler import GpuMedDataset
    rs = np.random.RandomState(0)
    images = [rs.randn(1, 48, 40, 36).astype(np.float32) for _ in range(2)]
    labels = [(rs.rand(1, 48, 40, 36) > 0.7).astype(np.uint8) for _ in range(2)]
    ds = GpuMedDataset(images, labels, 4, [32, 32, 32], class_probabilities=[0.5, 0.5], rng=np.random.RandomState(1))
    import argparse
    from mednet_b200.segmentation import SegmentationUNet3D
    hp = argparse.Namespace(in_channels=1, out_channels=2, fmaps=[16, 32], learning_rate=1e-3, num_workers=0,
                            batch_size=2, loss="DICE", loss_weight=[0.5, 0.5])
    net = SegmentationUNet3D(hp).cuda()
    opt = net.configure_optimizers()
    losses = []
    for batch in ds.loader(2, shuffle=True):
        opt.zero_grad()
        out = net.training_step(batch, 0)
        out["loss"].backward()
        opt.step()
        losses.append(float(out["loss"].detach()))
    assert len(losses) == 4 and all(np.isfinite(losses))


@pytest.mark.parametrize("dtype,channels,size", [(torch.float32, 1, (16, 24, 8)), (torch.float32, 3, (20, 12, 28)),
                                                  (torch.bfloat16, 2, (16, 16, 16)), (torch.float32, 1, (64, 64, 64))])
def test_intensity_augmentation_matches_oracle_chain(dtype, channels, size):
    """Device kernels vs the restated brightness -> gamma -> contrast chain (oracle/augment.py, parity unpinned: the
    library itself is not available), same seed.  fp32 tolerance: 1e-5 of the value range (powf and the summation order
    of the channel mean differ in the last bits); bf16 output: one rounding, 2^-8 relative."""
    from oracle import augment as oaug
    from mednet_b200 import ops
    from mednet_b200.sampler import IntensityAugmentation
    rs = np.random.RandomState(3)
    B = 3
    x = (rs.randn(B, channels, *size) * 1.5 + 0.5).astype(np.float32)
    np.random.seed(7)
    want = np.stack([oaug.augment_patch(x[b]) for b in range(B)])
    aug = IntensityAugmentation(rng=np.random.RandomState(7))
    coef = torch.as_tensor(np.stack([aug.draw(channels) for _ in range(B)])).cuda()
    xd = torch.as_tensor(x).cuda().permute(0, 2, 3, 4, 1).contiguous()
    y = ops.k_intensity_augment(xd, coef, dtype).permute(0, 4, 1, 2, 3).float().cpu().numpy()
    span = float(want.max() - want.min())
    if dtype == torch.float32:
        assert np.abs(y - want).max() <= 1e-5 * span
    else:
        assert (np.abs(y - want) <= 2.0 ** -8 * np.abs(want) + 1e-5 * span).all()
    # range preservation (contrast step): every channel stays inside the range it had after the gamma step
    assert y.min() >= want.min() - 1e-5 * span and y.max() <= want.max() + 1e-5 * span
    y2 = ops.k_intensity_augment(xd, coef, dtype).permute(0, 4, 1, 2, 3).float().cpu().numpy()
    assert np.array_equal(y, y2)                                        # fixed-order partials: run-to-run identical


def test_identity_coefficients_and_brightness_only_are_exact():
    from mednet_b200 import ops
    x = torch.randn(2, 8, 8, 8, 2, device="cuda")
    coef = torch.tensor([[0, 0, 0, 0, 1, 1], [0, 0, 0.25, -0.5, 1, 1]], dtype=torch.float32, device="cuda")
    y = ops.k_intensity_augment(x, coef, torch.float32)
    assert torch.equal(y[0], x[0])
    assert torch.equal(y[1], x[1] + torch.tensor([0.25, -0.5], device="cuda"))
    with pytest.raises(Exception):
        ops.k_intensity_augment(torch.randn(1, 4, 4, 4, 9, device="cuda"), torch.zeros(1, 20, device="cuda"), torch.float32)


def test_sampler_with_augmentation_interleaves_draws_like_getitem():
    """Position draws and augmentation draws share one generator, patch after patch (dataset.py:297-341): replaying the
    oracle's position sampling + augmentation chain with the same seed reproduces the batch."""
    from oracle import augment as oaug
    from mednet_b200.sampler import GpuMedDataset, IntensityAugmentation
    images, labels, _ = _cohort(5, False)
    P, probs = [12, 16, 8], [0.3, 0.4, 0.3]
    ds = GpuMedDataset(images, labels, 2, P, class_probabilities=probs, data_dtype=torch.float32,
                       rng=np.random.RandomState(21), augmentation=IntensityAugmentation())
    batch = ds.batch(range(4))
    np.random.seed(21)
    for b in range(4):
        s = b % 3
        ini, cls = osamp.sample_patch_position(labels[s][0], P, probs, osamp.label_any_maps(labels[s][0], 3))
        assert np.array_equal(ini, batch["patch_position"][b]) and cls == batch["selected_class"][b]
        crop, _ = osamp.crop_patch(images[s], labels[s], ini, P)
        want = oaug.augment_patch(crop)
        got = batch["data"][b].cpu().numpy()
        assert np.abs(got - want).max() <= 1e-5 * float(want.max() - want.min())


def test_tile_gather_channel_major_uint8_with_padding():
    """mednet_tile_gather's channel-major / uint8 mode (one volume, padded coordinates) against a NumPy crop."""
    from mednet_b200 import _abi, ops
    from mednet_b200._abi import check, lib, make
    rs = np.random.RandomState(4)
    vol = rs.randint(0, 256, size=(3, 20, 17, 23)).astype(np.uint8)
    P, O = (8, 6, 10), (2, 1, 3)
    origins = np.array([[0, 0, 0], [12, 11, 13], [14, 12, 16]], dtype=np.int32)      # the last ones reach past the volume
    vd = torch.as_tensor(vol).cuda()
    od = torch.as_tensor(origins).cuda()
    out = torch.empty((3, 3, *P), dtype=torch.uint8, device="cuda")
    gp = make("mednet_tile_gather_params", volume=vd.data_ptr(), tiles=out.data_ptr(), origins=od.data_ptr(), B=3, C=3,
              X=20, Y=17, Z=23, P0=P[0], P1=P[1], P2=P[2], O0=O[0], O1=O[1], O2=O[2], src_dtype=2, dst_dtype=2, ncdhw_out=1)
    check(lib().mednet_tile_gather(_abi.C.byref(gp), ops._stream()), "tile_gather")
    padded = np.pad(vol, [(0, 0)] + [(O[k], P[k]) for k in range(3)])
    for b, o in enumerate(origins):
        want = padded[:, o[0]:o[0] + P[0], o[1]:o[1] + P[1], o[2]:o[2] + P[2]]
        assert np.array_equal(out[b].cpu().numpy(), want)


def test_reference_constructor_reads_groups_and_keeps_all_label_channels():
    """MedDataset(data_path, subject_keys, ...) with the reference's argument order (dataset.py:211-222): groups come from a
    reader, images are stored with the reference's float16 rounding, every label channel is cropped and the LAST one
    drives the class-balanced positions (dataset.py:307, 319-321)."""
    from mednet_b200.dataset import DataReaderArrays, MedDataset
    rs = np.random.RandomState(8)
    shapes = {"s1": (24, 20, 18), "s2": (20, 26, 22)}
    store = {"images": {k: rs.randn(1, *s).astype(np.float32) for k, s in shapes.items()},
             "labels": {k: np.stack([rs.randint(0, 2, s), (rs.rand(*s) > 0.9) * 1]).astype(np.uint8) for k, s in shapes.items()},
             "heatmaps": {k: rs.randint(0, 256, (2,) + s).astype(np.uint8) for k, s in shapes.items()}}
    P, probs = [8, 8, 8], [0.2, 0.8]
    ds = MedDataset(store, ["s1", "s2"], 3, P, "images", "labels", "heatmaps", DataReaderArrays, probs,
                    data_dtype=torch.float32, rng=np.random.RandomState(2))
    assert len(ds) == 6
    batch = ds.batch(range(4))
    assert batch["label"].shape == (4, 2 + 2, *P) and batch["subject_key"] == ["s1", "s2", "s1", "s2"]
    np.random.seed(2)
    for b, key in enumerate(batch["subject_key"]):
        cmap = store["labels"][key][-1]
        ini, cls = osamp.sample_patch_position(cmap, P, probs, osamp.label_any_maps(cmap, 2))
        assert np.array_equal(ini, batch["patch_position"][b]) and cls == batch["selected_class"][b]
        full = np.concatenate([store["heatmaps"][key], store["labels"][key]], axis=0)
        want_d, want_l = osamp.crop_patch(store["images"][key].astype(np.float16), full, ini, P)
        assert np.array_equal(batch["data"][b].cpu().numpy(), want_d)
        assert np.array_equal(batch["label"][b].cpu().numpy(), want_l)
