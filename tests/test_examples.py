"""The re-hosted entry points (examples/train_seg.py, train_ldmks.py, predict.py) keep the reference's flag surface
(/root/reference/examples/train_seg.py:34-59, train_ldmks.py:33-59, midasmednet/landmarks.py:191-205) and run end to end
on synthetic data."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

REFERENCE_FLAGS = ["--config", "--seed", "--neptune_project", "--experiment_name", "--data_path", "--image_group",
                   "--label_group", "--train_set", "--val_set", "--model_dir", "--log_dir", "--patch_size",
                   "--class_probabilities", "--patches_per_subject", "--data_augmentation", "--gpus", "--preload",
                   "--resume", "--max_epochs", "--log_level"]
MODEL_FLAGS = ["--learning_rate", "--fmaps", "--batch_size", "--num_workers", "--in_channels", "--out_channels",
               "--log_interval", "--log_vis_mip"]


def _flags(parser):
    return {s for a in parser._actions for s in a.option_strings}


def test_flag_surface_matches_the_reference_scripts(monkeypatch, tmp_path):
    import _common
    from mednet_b200.landmarks import LandmarkNet
    from mednet_b200.segmentation import SegmentationNet
    seg = _flags(SegmentationNet.add_model_specific_args(_common.experiment_parser("aorth")))
    ldm = _flags(LandmarkNet.add_model_specific_args(_common.experiment_parser("aorth_ldmks", heatmaps=True)))
    for f in REFERENCE_FLAGS + MODEL_FLAGS:
        assert f in seg and f in ldm, f
    assert {"--loss", "--loss_weight"} <= seg                                  # read at segmentation.py:43-49
    assert {"--heatmap_group", "--loss_class", "--loss_class_weight", "--loss_regression",
            "--loss_regression_weight"} <= ldm
    # reference defaults (train_ldmks.py:46, landmarks.py:194-205)
    p = LandmarkNet.add_model_specific_args(_common.experiment_parser("aorth_ldmks", heatmaps=True))
    d = p.parse_args([])
    assert d.patch_size == [96, 96, 96] and d.fmaps == 64 and d.batch_size == 4 and d.gpus == 1
    assert d.loss_regression_weight == [0.001, 0.015, 0.015, 0.015, 0.001, 0.001] and d.loss_class_weight == [0.05, 1.0]
    # $DATA / $MODEL substitution with quirk Q2 fixed; YAML config file values are defaults, the command line wins
    monkeypatch.setenv("DATA", "/d")
    monkeypatch.setenv("MODEL", "/m")
    assert _common.replace_env("$DATA/x:$MODEL/y") == "/d/x:/m/y"
    cfg = tmp_path / "c.yaml"
    cfg.write_text("max_epochs: 7\nseed: 3\n")
    a = _common.parse_with_config(p, ["-c", str(cfg), "--seed", "5"])
    assert a.max_epochs == 7 and a.seed == 5


@pytest.mark.gpu
def test_train_then_predict_entry_points(tmp_path):
    import predict
    import train_ldmks
    import train_seg
    mdir = str(tmp_path / "seg")
    tr = train_seg.main(["--synthetic", "4", "--patch_size", "32", "32", "32", "--batch_size", "2", "--num_workers", "0",
                         "--fmaps", "8", "--out_channels", "2", "--max_epochs", "1", "--model_dir", mdir])
    assert os.path.exists(os.path.join(mdir, "epoch=0.ckpt")) and tr.history
    out = predict.main(["--checkpoint", os.path.join(mdir, "epoch=0.ckpt"), "--model", "SegmentationNet", "--synthetic",
                        "40", "33", "48", "--patch_size", "32", "32", "32", "--patch_overlap", "4", "4", "4",
                        "--batch_size", "3", "--output", str(tmp_path / "p.npy")])
    assert tuple(out.shape) == (1, 40, 33, 48) and out.dtype.is_floating_point is False
    assert np.load(str(tmp_path / "p.npy")).max() <= 1
    ldir = str(tmp_path / "ldm")
    train_ldmks.main(["--synthetic", "2", "--patch_size", "32", "32", "32", "--batch_size", "1", "--num_workers", "0",
                      "--fmaps", "8", "--out_channels", "5", "--loss_regression_weight", "0.001", "0.015", "0.015",
                      "--max_epochs", "1", "--model_dir", ldir, "--arch", "unet3d"])
    out = predict.main(["--checkpoint", os.path.join(ldir, "epoch=0.ckpt"), "--model", "LandmarkUNet3D", "--synthetic",
                        "32", "32", "32", "--patch_size", "32", "32", "32", "--patch_overlap", "4", "4", "4", "--sigma",
                        "3", "3", "3"])
    assert tuple(out.shape) == (4, 32, 32, 32)


@pytest.mark.gpu
def test_training_entry_points_with_the_gpu_resident_sampler(tmp_path):
    """--gpu_sampler: patches drawn by the device-resident MedDataset (class-balanced positions, heatmaps + class map)."""
    import train_ldmks
    import train_seg
    tr = train_seg.main(["--synthetic", "3", "--gpu_sampler", "48", "40", "36", "--class_probabilities", "0.3", "0.7",
                         "--patches_per_subject", "2", "--patch_size", "32", "32", "32", "--batch_size", "2", "--data_augmentation",
                         "--fmaps", "8", "--out_channels", "2", "--max_epochs", "1", "--model_dir", str(tmp_path / "s")])
    assert tr.history and all(np.isfinite(list(h.values())).all() for h in tr.history)
    tr = train_ldmks.main(["--synthetic", "2", "--gpu_sampler", "40", "40", "40", "--patches_per_subject", "2",
                           "--patch_size", "32", "32", "32", "--batch_size", "2", "--fmaps", "8", "--out_channels", "4",
                           "--loss_regression_weight", "0.01", "0.01", "--max_epochs", "1", "--arch", "unet3d"])
    assert tr.history
