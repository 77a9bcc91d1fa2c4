"""CPU-only tests of the host-side mirror of the reference interface: module tree / state_dict keys,
constructor errors, tiling geometry against the reference generator's golden output, and that the product
refuses CPU tensors instead of falling back."""
import numpy as np
import pytest
import torch

from conftest import state_dict_from
from mednet_b200.dataset import SyntheticSegmentationDataset, grid_geometry, grid_patch_generator
from mednet_b200.landmarks import LandmarkNet
from mednet_b200.segmentation import SegmentationNet
from mednet_b200.unet.components import SingleConv, create_conv
from mednet_b200.unet.loss import DiceLoss
from mednet_b200.unet.model import ResidualUNet3D, UNet3D


def test_state_dict_keys_and_shapes_match_reference(golden):
    g = golden("unet3d_small")
    ref = {k: tuple(v.shape) for k, v in state_dict_from(g).items()}
    net = UNet3D(1, 2, False, f_maps=[8, 16, 32])
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == ref
    net.load_state_dict(state_dict_from(g))
    g = golden("residual_small")
    ref = {k: tuple(v.shape) for k, v in state_dict_from(g).items()}
    net = ResidualUNet3D(1, 4, False, f_maps=[8, 16, 32])
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == ref
    assert len(UNet3D(1, 2, False).state_dict()) == 44
    assert len(ResidualUNet3D(1, 2, False, f_maps=32).state_dict()) == 91
    for order in ("crg", "cl", "gce"):
        g = golden("unet3d_orders")
        ref = {k: tuple(v.shape) for k, v in state_dict_from(g, f"{order}.sd.").items()}
        net = UNet3D(2, 3, False, f_maps=[8, 16], layer_order=order)
        assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == ref


def test_reference_attributes_and_defaults():
    net = UNet3D(1, 2, False)
    assert len(net.encoders) == 4 and len(net.decoders) == 3 and net.testing is False
    assert net.final_conv.weight.shape == (2, 64, 1, 1, 1)
    assert net.encoders[0].pooling is None and net.encoders[1].pooling is not None
    assert net.encoders[0].basic_module.SingleConv1.groupnorm.num_groups == 1      # quirk Q13: C=1 < 8 groups
    assert net.decoders[0].basic_module.SingleConv1.conv.weight.shape == (256, 768, 3, 3, 3)
    res = ResidualUNet3D(1, 2, False)
    assert len(res.encoders) == 5 and res.decoders[0].upsample.weight.shape == (512, 256, 3, 3, 3)
    assert ResidualUNet3D(1, 2, False, skip_final_activation=True).final_activation is None
    assert UNet3D(1, 2, False, testing=True).testing is True


def test_constructor_errors_match_reference():
    with pytest.raises(AssertionError, match="Conv layer MUST be present"):
        create_conv(4, 4, 3, "gr", 8)
    with pytest.raises(AssertionError, match="Non-linearity cannot be the first"):
        create_conv(4, 4, 3, "rc", 8)
    with pytest.raises(ValueError, match="Unsupported layer type"):
        create_conv(4, 4, 3, "cx", 8)
    with pytest.raises(AssertionError, match="divisible by num_groups"):
        create_conv(12, 12, 3, "gc", 8)
    assert SingleConv(4, 8, order="gcr")._plan == [("g", 0), ("c", 1)]
    assert SingleConv(4, 8, order="cge")._plan == [("c", 0), ("g", 3)]
    assert SingleConv(4, 8, order="crg")._plan == [("c", 1), ("g", 0)]
    assert SingleConv(4, 8, order="cl").conv.bias is not None and SingleConv(4, 8, order="gcr").conv.bias is None


def test_default_init_matches_pytorch_conv3d_statistics():
    torch.manual_seed(0)
    mine = SingleConv(16, 32, order="cr").conv
    ref = torch.nn.Conv3d(16, 32, 3, padding=1)
    bound = 1 / (16 * 27) ** 0.5
    assert mine.weight.abs().max() <= bound and mine.bias.abs().max() <= bound
    assert abs(mine.weight.std().item() - ref.weight.std().item()) < 2e-3


def test_no_cpu_fallback():
    net = UNet3D(1, 2, False, f_maps=[8, 16])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DiceLoss()(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 4, 4, 4, dtype=torch.long))


def test_grid_geometry_matches_reference_generator(golden):
    g = golden("tiling")
    for tag in "abc":
        img, (p, o) = g[f"{tag}.img"], g[f"{tag}.patch"]
        _, _, _, origins = grid_geometry(img.shape[1:], [p] * 3, [o] * 3)
        np.testing.assert_array_equal(origins, g[f"{tag}.pos"])
        sums = [patch.astype(np.float64).sum() for patch, _, _ in grid_patch_generator(img, [p] * 3, [o] * 3, mode="constant")]
        np.testing.assert_allclose(sums, g[f"{tag}.sums"], rtol=1e-12)
    assert len(grid_geometry((512, 512, 400), [128] * 3, [16] * 3)[3]) == 180      # cfg-4 tile counts
    assert len(grid_geometry((512, 512, 400), [128] * 3, [32] * 3)[3]) == 448


def test_task_modules_construct_with_reference_hparams():
    import argparse
    hp = argparse.Namespace(in_channels=1, out_channels=2, fmaps=8, learning_rate=1e-3, num_workers=0, batch_size=2,
                            loss="DICE", loss_weight=[0.05, 1.0])
    seg = SegmentationNet(hp)
    assert "loss.weight" in seg.state_dict()                                      # DiceLoss buffer key (loss.py:100)
    p = LandmarkNet.add_model_specific_args(argparse.ArgumentParser())
    a = p.parse_args([])
    assert a.fmaps == 64 and a.loss_regression_weight == [0.001, 0.015, 0.015, 0.015, 0.001, 0.001]
    a.fmaps, a.out_channels = 8, 8
    lm = LandmarkNet(a)
    assert lm.num_heatmaps == 6 and "loss_class.weight" in lm.state_dict()
    ds = SyntheticSegmentationDataset(2, (8, 8, 8), num_classes=2, num_heatmaps=3)
    item = ds[0]
    assert item["data"].shape == (1, 8, 8, 8) and item["label"].shape == (4, 8, 8, 8) and item["label"].dtype == torch.uint8


def test_fused_adam_state_dict_is_torch_adams_format_and_round_trips():
    """FusedAdam.state_dict / load_state_dict (what Trainer.fit restores on resume_from_checkpoint, the reference's
    examples/train_seg.py:122-131 through pytorch-lightning): torch.optim.Adam's own layout, one step count for the
    bucket, a single parameter group."""
    import pytest
    import torch
    from mednet_b200.optim import FusedAdam
    ps = [torch.nn.Parameter(torch.randn(3, 2)), torch.nn.Parameter(torch.randn(4))]
    opt = FusedAdam(ps, lr=2e-3, async_wgrad=False)
    assert opt.state_dict()["state"] == {}                      # nothing stepped yet, like torch
    opt._materialize()
    opt._step = 7
    opt._m.copy_(torch.arange(10.0))
    opt._v.copy_(torch.arange(10.0) * 2)
    sd = opt.state_dict()
    assert sd["param_groups"][0]["lr"] == 2e-3 and sd["param_groups"][0]["params"] == [0, 1]
    assert float(sd["state"][1]["step"]) == 7.0 and tuple(sd["state"][0]["exp_avg"].shape) == (3, 2)
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4))])
    ref.load_state_dict(sd)                                      # stock Adam accepts it
    assert torch.equal(ref.state_dict()["state"][1]["exp_avg_sq"], torch.arange(6.0, 10.0) * 2)
    opt2 = FusedAdam([torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4))], async_wgrad=False)
    opt2.load_state_dict(ref.state_dict())                       # and what stock Adam saved loads back
    assert opt2._step == 7 and torch.equal(opt2._m, torch.arange(10.0)) and opt2.param_groups[0]["lr"] == 2e-3
    with pytest.raises(NotImplementedError):
        FusedAdam([{"params": ps[:1]}, {"params": ps[1:]}])
    bad = ref.state_dict()
    bad["state"][0]["step"] = torch.tensor(3.0)
    with pytest.raises(ValueError, match="one step count"):
        opt2.load_state_dict(bad)


def test_load_from_checkpoint_reads_a_pl09_shaped_file(tmp_path):
    """examples/predict.py:46-50 loads `SegmentationNet.load_from_checkpoint(path)` on files written by pytorch-lightning
    0.9.  pytorch-lightning is absent here, so the file is assembled by hand with the keys PL 0.9's
    `Trainer.dump_checkpoint` writes (epoch, global_step, pytorch-lightning_version, checkpoint_callback_*,
    optimizer_states, lr_schedulers, state_dict, hparams_name, hyper_parameters): the loader must take the weights
    and the hyper-parameters from it and ignore the rest."""
    import argparse
    import torch
    hp = dict(in_channels=1, out_channels=2, fmaps=8, learning_rate=1e-3, num_workers=0, batch_size=2, loss="DICE",
              loss_weight=[0.05, 1.0])
    torch.manual_seed(3)
    src = SegmentationNet(argparse.Namespace(**hp))
    adam = torch.optim.Adam(src.parameters(), lr=1e-3)
    ckpt = {"epoch": 4, "global_step": 120, "pytorch-lightning_version": "0.9.0",
            "checkpoint_callback_best_model_score": torch.tensor(0.41), "checkpoint_callback_best_model_path": "epoch=3.ckpt",
            "optimizer_states": [adam.state_dict()], "lr_schedulers": [],
            "state_dict": {k: v.clone() for k, v in src.state_dict().items()},
            "hparams_name": "hparams", "hyper_parameters": dict(hp)}
    path = str(tmp_path / "epoch=4.ckpt")
    torch.save(ckpt, path)
    net = SegmentationNet.load_from_checkpoint(path)
    assert vars(net.hparams) == hp
    assert list(net.state_dict()) == list(src.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), src.state_dict().values()))
    # the older layout of the same release line: the Namespace itself under 'hparams'
    ckpt2 = dict(ckpt)
    ckpt2.pop("hyper_parameters")
    ckpt2["hparams"] = argparse.Namespace(**hp)
    torch.save(ckpt2, path)
    assert vars(SegmentationNet.load_from_checkpoint(path).hparams) == hp
