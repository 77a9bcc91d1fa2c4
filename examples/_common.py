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
