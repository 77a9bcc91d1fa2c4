"""CPU oracle for the torch-mednet UNet3D hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import it, and only as the checker.  The product path
(``torch-mednet_b200/mednet_b200``) never imports this package and fails loudly when
its CUDA extension is missing.

The oracle is a plain-PyTorch fp32 restatement (functional style, driven by a
``state_dict``) of the reference's algorithm:

* ``oracle.unet``    -- midasmednet/unet/model.py + components.py
* ``oracle.loss``    -- midasmednet/unet/loss.py (DiceLoss, dice_metric) and the loss
                        wiring of segmentation.py / landmarks.py
* ``oracle.steps``   -- training_step / predict epilogue restatements
* ``oracle.tiling``  -- dataset.py grid_patch_generator / add_processed_batch geometry
* ``oracle.heatmaps``-- builder-specified heatmap rendering and landmark extraction
                        (ABSENT from the reference)

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the *live reference modules* imported in the build container
(``oracle/make_golden.py``); the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py`` on every run.
The arithmetic itself lives in third-party PyTorch (requirements.txt:5 pins
torch==1.5.1; this container and the GPU box run torch 2.11.0 -- semantics of the ops
used are unchanged for these argument patterns).
"""
