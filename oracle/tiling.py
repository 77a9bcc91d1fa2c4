"""NumPy restatement of the reference sliding-window geometry (test oracle; see oracle/__init__.py).

Follows /root/reference/midasmednet/dataset.py:
  grid_patch_generator ........... dataset.py:349-389
  GridPatchSampler.add_processed_batch dataset.py:444-474 (zarr output array replaced by np.zeros)
and the inference epilogue of examples/predict.py:85-94.
"""
from __future__ import annotations

import numpy as np


def grid_geometry(img_size, patch_size, patch_overlap):
    """Tile grid of dataset.py:366-380: crop size c = P - 2o, n = ceil(S / c),
    pad low o, high o + (c - S mod c)  (a full extra c when S mod c == 0, quirk Q8)."""
    img_size = np.asarray(img_size)
    patch_size = np.asarray(patch_size)
    patch_overlap = np.asarray(patch_overlap)
    cropped = patch_size - 2 * patch_overlap
    n_patches = np.ceil(img_size / cropped).astype(int)
    overhead = cropped - img_size % cropped
    pos = [np.arange(0, n_patches[k]) * cropped[k] for k in range(3)]
    return cropped, n_patches, overhead, pos


def grid_patches(img, patch_size, patch_overlap, **pad_kwargs):
    """Yield (patch CxPxPxP, position, count) in raster order (dataset.py:380-389)."""
    patch_size = np.asarray(patch_size)
    patch_overlap = np.asarray(patch_overlap)
    _, _, overhead, pos = grid_geometry(img.shape[1:], patch_size, patch_overlap)
    pads = [[0, 0]] + [[patch_overlap[k], patch_overlap[k] + overhead[k]] for k in range(3)]
    padded = np.pad(img, pads, **pad_kwargs)
    count = -1
    for p0 in pos[0]:
        for p1 in pos[1]:
            for p2 in pos[2]:
                count += 1
                yield (padded[:, p0:p0 + patch_size[0], p1:p1 + patch_size[1], p2:p2 + patch_size[2]],
                       np.array([p0, p1, p2]), count)


def stitch_patch(result, patch, pos, patch_overlap):
    """dataset.py:452-474 -- centre crop [o:-o], clip the overhang, plain overwrite (no blending).

    The reference's axis-0 slice mixes overlap[0] and overlap[1] (quirk Q7); identical for the
    isotropic overlaps every reference caller uses, which is what is restated here per axis.
    """
    o = np.asarray(patch_overlap)
    cropped = patch[:, o[0]:patch.shape[1] - o[0], o[1]:patch.shape[2] - o[1], o[2]:patch.shape[3] - o[2]]
    pos = np.asarray(pos)
    pos_end = pos + np.array(cropped.shape[1:])
    img_size = np.array(result.shape[1:])
    crop_end = np.minimum(pos_end, img_size)
    new = np.array(cropped.shape[1:]) - np.maximum(pos_end - crop_end, 0)
    result[:, pos[0]:pos_end[0], pos[1]:pos_end[1], pos[2]:pos_end[2]] = \
        cropped[:, :new[0], :new[1], :new[2]].astype(result.dtype)


def predict_epilogue(logits, num_heatmaps):
    """examples/predict.py:88-94 on a numpy fp32 array (B, L+K, P, P, P) -> uint8 (B, L+1, P, P, P).

    argmax(softmax(x)) == first maximal index; heatmaps are clipped to [0,255] and truncated by
    ``astype(uint8)``.
    """
    import torch
    t = torch.from_numpy(np.ascontiguousarray(logits))
    cls = torch.argmax(torch.softmax(t[:, num_heatmaps:], dim=1), dim=1, keepdim=True).numpy()
    hm = np.clip(logits[:, :num_heatmaps], 0.0, 255.0)
    return np.concatenate([hm.astype(np.uint8), cls.astype(np.uint8)], axis=1)


def sliding_window_predict(volume, forward_fn, patch_size, patch_overlap, num_heatmaps, out_channels,
                           batch_size=1):
    """The loop of examples/predict.py:82-97 for one subject.  ``forward_fn`` maps a float32
    numpy batch (B,C,P,P,P) to float32 logits (B,L+K,P,P,P)."""
    result = np.zeros((out_channels,) + tuple(volume.shape[1:]), dtype=np.uint8)
    batch, poss = [], []

    def flush():
        if not batch:
            return
        out = predict_epilogue(forward_fn(np.stack(batch).astype(np.float32)), num_heatmaps)
        for b, pos in enumerate(poss):
            stitch_patch(result, out[b], pos, patch_overlap)
        batch.clear()
        poss.clear()

    for patch, pos, _ in grid_patches(volume, patch_size, patch_overlap, mode="constant"):
        batch.append(patch)
        poss.append(pos)
        if len(batch) == batch_size:
            flush()
    flush()
    return result
