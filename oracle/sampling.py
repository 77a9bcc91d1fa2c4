"""Restatement of the reference's random patch sampling (test oracle; see oracle/__init__.py).

midasmednet/dataset.py cannot be imported here (h5py / zarr / nibabel missing, and it uses the removed ``np.int``,
SURVEY.md appendix Q15), so the sampling logic is restated with the SAME sequence of ``np.random`` calls; the golden
vectors in tests/golden/sampling.npz were produced by the reference's own function bodies (extracted with ``ast`` in
oracle/make_golden.py, ``np.int`` mapped to ``int``) and pin this restatement call for call.

  get_labeled_position .......... dataset.py:18-51
  get_random_patch_indices ...... dataset.py:54-88
  MedDataset.__getitem__ ........ dataset.py:285-346 (class choice, position, crop; label map = LAST label channel)
  MedDataset.__init__ (any-maps)  dataset.py:268-279
"""
from __future__ import annotations

import numpy as np


def label_any_maps(class_map, num_classes):
    """dataset.py:274-279: per class c, np.any(label == c, axis=2) -- makes the position sampling cheap."""
    return [np.any(class_map == c, axis=2) for c in range(num_classes)]


def get_labeled_position(label, class_value, label_any=None):
    """dataset.py:18-51.  NB (preserved): the third index is drawn from ``np.argwhere(row == class_value)[0]``, i.e. from
    the FIRST matching index only (a 1-element array) -- the reference never samples along the third axis."""
    if label_any is None:
        label_any = np.any(label == class_value, axis=2)
    valid_idx = np.argwhere(label_any == True)                      # noqa: E712  (as in the reference)
    if valid_idx.size:
        rnd = np.random.randint(0, valid_idx.shape[0])
        idx = valid_idx[rnd]
        valid_idx = label[idx[0], idx[1], :]
        valid_idx = np.argwhere(valid_idx == class_value)[0]
        rnd = np.random.choice(valid_idx)
        return [idx[0], idx[1], rnd]
    return None


def get_random_patch_indices(patch_size, img_shape, pos=None):
    """dataset.py:54-88: a patch that contains ``pos`` (if given) and lies inside the image."""
    patch_size = np.asarray(patch_size)
    img_shape = np.asarray(img_shape)
    if pos:
        pos = np.array(pos, dtype=int)
        min_index = np.maximum(pos - patch_size + 1, 0)
        max_index = np.minimum(img_shape - patch_size + 1, pos + 1)
    else:
        min_index = np.array([0, 0, 0])
        max_index = img_shape - patch_size + 1
    index_ini = np.random.randint(low=min_index, high=max_index)
    return index_ini, index_ini + patch_size


def sample_patch_position(class_map, patch_size, class_probabilities=None, any_maps=None):
    """The position part of MedDataset.__getitem__ (dataset.py:297-313): returns (index_ini, selected_class)."""
    pos, selected_class = None, 0
    if class_probabilities is not None:
        p = np.asarray(class_probabilities, dtype=np.float64)
        p = p / p.sum()                                               # dataset.py:253-255
        selected_class = np.random.choice(range(len(p)), p=p)
        if selected_class > 0:
            pos = get_labeled_position(class_map, selected_class,
                                       label_any=None if any_maps is None else any_maps[selected_class])
    index_ini, _ = get_random_patch_indices(np.asarray(patch_size), np.asarray(class_map.shape), pos=pos)
    return index_ini, int(selected_class)


def crop_patch(image, label, index_ini, patch_size):
    """dataset.py:315-336: crops (C,H,W,D) image -> float32 and (L+1,H,W,D) label -> uint8."""
    a, b = np.asarray(index_ini), np.asarray(index_ini) + np.asarray(patch_size)
    sl = (slice(None), slice(a[0], b[0]), slice(a[1], b[1]), slice(a[2], b[2]))
    return image[sl].astype(np.float32), label[sl].astype(np.uint8)
