"""Shared by the re-hosted entry points: path bootstrap, `$DATA`/`$MODEL` substitution (fixing quirk Q2 of the
reference: examples/train_seg.py:27-31 discards the first substitution) and the experiment-level flags of
examples/train_seg.py:34-55 / train_ldmks.py:33-56 (same names, types and defaults)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-mednet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def replace_env(path_str):
    return path_str.replace("$DATA", os.getenv("DATA", "")).replace("$MODEL", os.getenv("MODEL", ""))


def experiment_parser(default_name, heatmaps=False):
    parser = argparse.ArgumentParser(add_help=False)
    parser.add_argument("-c", "--config", default=None, help="YAML file of flag values (configargparse in the reference)")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--neptune_project", type=str, default="lab-midas/mednet")      # accepted, unused (SaaS logger)
    parser.add_argument("--experiment_name", type=str, default=default_name)
    parser.add_argument("--data_path", type=replace_env)
    parser.add_argument("--image_group", type=str, default="images")
    parser.add_argument("--label_group", type=str, default="labels")
    if heatmaps:
        parser.add_argument("--heatmap_group", type=str, default="heatmaps")
    parser.add_argument("--train_set", type=str)
    parser.add_argument("--val_set", type=str)
    parser.add_argument("--model_dir", type=replace_env)
    parser.add_argument("--log_dir", type=replace_env)
    parser.add_argument("--patch_size", type=int, nargs="+", default=[96, 96, 96])
    parser.add_argument("--class_probabilities", type=float, nargs="+", default=None)
    parser.add_argument("--patches_per_subject", type=int, default=10)
    parser.add_argument("--data_augmentation", action="store_true")
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--preload", action="store_true")
    parser.add_argument("--resume", type=str)
    parser.add_argument("--max_epochs", type=int, default=100)
    parser.add_argument("--log_level", type=str, default="INFO")
    # additions of the B200 build
    parser.add_argument("--synthetic", type=int, default=0, metavar="N",
                        help="train on N synthetic patches (MedDataset contract) instead of an HDF5 file")
    parser.add_argument("--max_steps", type=int, default=None)
    parser.add_argument("--gpu_sampler", type=int, nargs=3, default=None, metavar=("X", "Y", "Z"),
                        help="with --synthetic N: N synthetic SUBJECTS of this size kept in GPU memory, patches drawn by "
                             "the device-resident MedDataset (class_probabilities / patches_per_subject as in the reference)")
    parser.add_argument("--arch", choices=["residual", "unet3d"], default="residual",
                        help="residual = what the reference task modules derive from; unet3d = the north-star UNet3D")
    return parser


def parse_with_config(parser, argv=None):
    """configargparse behaviour: values from the YAML file given with -c are defaults, the command line wins."""
    args, _ = parser.parse_known_args(argv)
    if args.config:
        import yaml
        with open(args.config) as f:
            parser.set_defaults(**(yaml.safe_load(f) or {}))
    return parser.parse_args(argv)


def require_dataset(hparams, what):
    if hparams.synthetic:
        return
    raise SystemExit(f"{what}: reading '{hparams.data_path}' needs h5py/zarr (midasmednet/dataset.py:109-207), which are "
                     "not part of this image and out of scope for the hot path (SURVEY.md section 2 row 9); "
                     "run with --synthetic N")


def seed_everything(seed):
    """examples/train_seg.py:61-65.  Weights need the same torch seed on every rank; the NumPy state that draws the
    patch positions (midasmednet/dataset.py:297-313) is offset by the rank so data-parallel ranks sample differently."""
    import numpy as np
    import torch
    torch.manual_seed(seed)
    np.random.seed(seed + int(os.environ.get("RANK", "0")))


def synthetic_cohort(n_subjects, size, in_channels, num_classes, num_heatmaps=0, seed=0):
    """Blobby synthetic subjects: images (C,X,Y,Z) f32, class maps (1,X,Y,Z) u8 (sparse foreground, so that the
    class-balanced sampling of dataset.py:297-306 has work to do), optional uint8 heatmaps."""
    import numpy as np
    rs = np.random.RandomState(seed)
    images, labels, heatmaps = [], [], []
    ax = [np.arange(s, dtype=np.float32) for s in size]
    for _ in range(n_subjects):
        lab = np.zeros(size, dtype=np.uint8)
        img = rs.randn(in_channels, *size).astype(np.float32) * 0.3
        for c in range(1, num_classes):
            ctr = [rs.uniform(0.2, 0.8) * s for s in size]
            rad = rs.uniform(0.08, 0.2) * min(size)
            r2 = ((ax[0][:, None, None] - ctr[0]) ** 2 + (ax[1][None, :, None] - ctr[1]) ** 2 +
                  (ax[2][None, None, :] - ctr[2]) ** 2)
            lab[r2 < rad ** 2] = c
            img += (r2 < rad ** 2)[None] * (0.5 * c)
        images.append(img)
        labels.append(lab[None])
        if num_heatmaps:
            pts = rs.uniform(0.1, 0.9, size=(num_heatmaps, 3)) * np.asarray(size)
            r2 = ((ax[0][None, :, None, None] - pts[:, 0, None, None, None]) ** 2 +
                  (ax[1][None, None, :, None] - pts[:, 1, None, None, None]) ** 2 +
                  (ax[2][None, None, None, :] - pts[:, 2, None, None, None]) ** 2)
            heatmaps.append((255.0 * np.exp(-r2 / (2 * 3.0 ** 2))).astype(np.uint8))
    return images, labels, (heatmaps if num_heatmaps else None)
