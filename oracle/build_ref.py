"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, made importable beside the oracle.

    python oracle/build_ref.py          # needs /root/reference (the build container); writes oracle/_ref/ only

The reference is pure Python: its path (`midasmednet/unet/{model,components,loss}.py`) "compiles" by being placed on an
import path next to a two-line `pytorch_lightning` stand-in (LightningModule = torch.nn.Module -- the package is
not installed in this image and the hot path uses nothing else of it).  The files are taken
unmodified from where they lie under /root/reference; nothing is written outside oracle/_ref/, which is git-ignored
(reference sources never enter the history) but travels to the GPU box with the snapshot, like a built .so.

Users: bench.py's `--impl reference` arm / `cpu_baseline` leg (`kind: "reference"`) and tests/test_reference_ref.py,
which checks the vendored modules against the committed golden vectors and the oracle restatement.  The product never
imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["midasmednet/unet/__init__.py", "midasmednet/unet/model.py", "midasmednet/unet/components.py",
         "midasmednet/unet/loss.py"]
STUB = '''"""Stand-in for pytorch-lightning 0.9 (requirements.txt:8), which is not installed in this image: the hot path only
derives its networks from LightningModule (midasmednet/unet/model.py:2,11,113)."""
import torch


class LightningModule(torch.nn.Module):
    pass
'''


def build(ref=REF, out=OUT):
    if not os.path.isdir(os.path.join(ref, "midasmednet", "unet")):
        return None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(ref, rel), os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    # package marker of our own: the reference's midasmednet/__init__.py is not needed (and may import more)
    open(os.path.join(out, "midasmednet", "__init__.py"), "w").write("")
    os.makedirs(os.path.join(out, "pytorch_lightning"), exist_ok=True)
    open(os.path.join(out, "pytorch_lightning", "__init__.py"), "w").write(STUB)
    json.dump({"source": ref, "sha256": manifest}, open(os.path.join(out, "MANIFEST.json"), "w"), indent=1)
    return out


def load(out=OUT):
    """Imports the vendored reference modules; returns (model_module, loss_module) or None when oracle/_ref is absent."""
    if not os.path.exists(os.path.join(out, "midasmednet", "unet", "model.py")):
        return None
    if out not in sys.path:
        sys.path.insert(0, out)
    from midasmednet.unet import loss as rloss
    from midasmednet.unet import model as rmodel
    return rmodel, rloss


if __name__ == "__main__":
    print(build() or "no /root/reference here: oracle/_ref not rebuilt")
