"""Sliding-window tiling geometry of the reference (midasmednet/dataset.py:349-389, :444-474) and the
synthetic datasets used for measurement.  The HDF5/zarr readers (dataset.py:109-260) are host-side I/O and out of
scope; MedDataset's random patch sampling lives in ``sampler.py`` (device-resident);
the synthetic datasets emit batches with exactly MedDataset's dict contract (dataset.py:332-346).
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset


def grid_geometry(img_size, patch_size, patch_overlap):
    """Tile grid of grid_patch_generator (dataset.py:366-389).

    Returns (cropped_patch_size, n_patches, overhead, origins) where ``origins`` is an (T, 3) int32 array
    of tile origins in PADDED coordinates, raster order over axes 0, 1, 2.  The padded image has
    ``patch_overlap`` voxels in front and ``patch_overlap + overhead`` behind (a full extra crop when the
    size divides evenly, quirk Q8 -- only positions matter, padding is applied lazily by the gather kernel).
    """
    img_size = np.asarray(img_size, dtype=np.int64)
    patch_size = np.asarray(patch_size, dtype=np.int64)
    patch_overlap = np.asarray(patch_overlap, dtype=np.int64)
    cropped = patch_size - 2 * patch_overlap
    if (cropped <= 0).any():
        raise ValueError("patch_size must exceed 2 * patch_overlap")
    n_patches = np.ceil(img_size / cropped).astype(np.int64)
    overhead = cropped - img_size % cropped
    pos = [np.arange(0, n_patches[k]) * cropped[k] for k in range(3)]
    grid = np.stack(np.meshgrid(*pos, indexing="ij"), axis=-1).reshape(-1, 3)
    return cropped, n_patches, overhead, grid.astype(np.int32)


def grid_patch_generator(img, patch_size, patch_overlap, **kwargs):
    """Host (NumPy) generator with the reference's signature and yield order; used by tests and by callers
    that want tiles on the CPU.  The GPU predictor gathers tiles directly from the device-resident volume."""
    patch_size = np.asarray(patch_size)
    patch_overlap = np.asarray(patch_overlap)
    _, _, overhead, origins = grid_geometry(img.shape[1:], patch_size, patch_overlap)
    pads = [[0, 0]] + [[int(patch_overlap[k]), int(patch_overlap[k] + overhead[k])] for k in range(3)]
    padded = np.pad(img, pads, **kwargs)
    for count, idx in enumerate(origins):
        end = idx + patch_size
        yield padded[:, idx[0]:end[0], idx[1]:end[1], idx[2]:end[2]], idx.copy(), count


class SyntheticSegmentationDataset(Dataset):
    """Random patches with MedDataset's contract: {'data': (C,H,W,D) f32, 'label': (L+1,H,W,D) u8} with the
    class map in the LAST label channel (dataset.py:322-336) and L optional uint8 Gaussian heatmaps."""

    def __init__(self, length, patch_size, in_channels=1, num_classes=2, num_heatmaps=0, sigma=3.0, seed=0):
        self.length, self.patch_size = length, tuple(patch_size)
        self.in_channels, self.num_classes, self.num_heatmaps, self.sigma, self.seed = \
            in_channels, num_classes, num_heatmaps, sigma, seed

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        data = torch.randn((self.in_channels,) + self.patch_size, generator=g)
        cls = torch.randint(0, self.num_classes, (1,) + self.patch_size, generator=g, dtype=torch.uint8)
        if self.num_heatmaps:
            pts = torch.rand(self.num_heatmaps, 3, generator=g) * torch.tensor(self.patch_size, dtype=torch.float32)
            ax = [torch.arange(s, dtype=torch.float32) for s in self.patch_size]
            r2 = ((ax[0][None, :, None, None] - pts[:, 0, None, None, None]) ** 2 +
                  (ax[1][None, None, :, None] - pts[:, 1, None, None, None]) ** 2 +
                  (ax[2][None, None, None, :] - pts[:, 2, None, None, None]) ** 2)
            hm = (255.0 * torch.exp(-r2 / (2 * self.sigma ** 2))).to(torch.uint8)
            label = torch.cat([hm, cls], dim=0)
        else:
            label = cls
        return {"data": data, "label": label}


# ------------------------------------------------------------------------------------------------------------------
# Reference-facing dataset surface (midasmednet/dataset.py:90-283): readers + MedDataset with the same constructor.
# The readers are thin adapters over the storage libraries (imported lazily; neither h5py nor zarr is part of this
# image) -- the work is in ``sampler.GpuMedDataset``, which keeps what they return in GPU memory.
# ------------------------------------------------------------------------------------------------------------------
def one_hot_to_label(data, add_background=True):
    """(C,H,W,D) one-hot -> (1,H,W,D) class values (dataset.py:90-107)."""
    data = np.asarray(data)
    if add_background:
        data = np.concatenate([~np.any(data, axis=0, keepdims=True), data], axis=0)
    return np.argmax(data, axis=0)[None]


class DataReader:
    """Reader interface of the reference (dataset.py:109-150): ``read`` yields one array per subject key."""

    def read(self, subject_keys, group, dtype=np.float16, preload=True):
        raise NotImplementedError

    def read_data_to_memory(self, subject_keys, group, dtype=np.float16, preload=True):
        return list(self.read(subject_keys, group, dtype, preload))

    def get_data_shape(self, subject_keys, group):
        return {k: tuple(np.shape(d)) for k, d in zip(subject_keys, self.read(subject_keys, group, None, False))}

    def close(self):
        pass


class DataReaderArrays(DataReader):
    """In-memory store ``{group: {key: array}}`` (or an ``.npz`` file with ``group/key`` entries)."""

    def __init__(self, path_data):
        if isinstance(path_data, (str, bytes)) or hasattr(path_data, "__fspath__"):
            npz = np.load(path_data)
            store = {}
            for name in npz.files:
                group, key = name.split("/", 1)
                store.setdefault(group, {})[key] = npz[name]
            path_data = store
        self.store = path_data

    def read(self, subject_keys, group, dtype=np.float16, preload=True):
        for k in subject_keys:
            data = self.store[group][k]
            yield np.asarray(data).astype(dtype) if (preload and dtype is not None) else data


class _HierarchicalReader(DataReader):
    """``<file>/<group>/<key>`` datasets of an HDF5 or zarr container (dataset.py:152-207)."""
    module = opener = None

    def __init__(self, path_data):
        try:
            lib = __import__(self.module)
        except ImportError as e:
            raise ImportError(f"{type(self).__name__} needs the '{self.module}' package, which is not installed; "
                              "use DataReaderArrays or pass your own DataReader subclass as ReaderClass") from e
        self.path_data = path_data
        self.root = getattr(lib, self.opener)(str(path_data), "r")

    def read(self, subject_keys, group, dtype=np.float16, preload=True):
        for k in subject_keys:
            data = self.root[f"{group}/{k}"]
            yield data[:].astype(dtype) if (preload and dtype is not None) else data

    def get_data_attribute(self, subject_keys, group, attribute):
        return {k: self.root[f"{group}/{k}"].attrs[attribute] for k in subject_keys}

    def close(self):
        if hasattr(self.root, "close"):
            self.root.close()


class DataReaderHDF5(_HierarchicalReader):
    module, opener = "h5py", "File"


class DataReaderZarr(_HierarchicalReader):
    module, opener = "zarr", "open"


def MedDataset(data_path, subject_keys, samples_per_subject, patch_size, image_group='images', label_group='labels',
               heatmap_group=None, ReaderClass=DataReaderHDF5, class_probabilities=None, preload=True, transform=None,
               **device_options):
    """The reference's constructor (dataset.py:211-260) returning the device-resident dataset: the groups are read once
    through ``ReaderClass`` with the reference's storage dtypes (images float16, labels / heatmaps uint8), moved to GPU
    memory and sampled there (``sampler.GpuMedDataset``; ``preload`` is implied).  ``transform``: an
    ``IntensityAugmentation`` runs as CUDA kernels; any other callable receives the batch dict of device tensors."""
    from .sampler import GpuMedDataset, IntensityAugmentation
    reader = ReaderClass(data_path)
    images = reader.read_data_to_memory(subject_keys, image_group, dtype=np.float16, preload=True)
    labels = reader.read_data_to_memory(subject_keys, label_group, dtype=np.uint8, preload=True)
    heatmaps = reader.read_data_to_memory(subject_keys, heatmap_group, dtype=np.uint8, preload=True) if heatmap_group else None
    reader.close()
    assert len(images) == len(labels)                                      # dataset.py:263
    augmentation = transform if isinstance(transform, IntensityAugmentation) else None
    return GpuMedDataset([np.asarray(i, dtype=np.float32) for i in images], list(labels),
                         samples_per_subject, patch_size, heatmaps=heatmaps, class_probabilities=class_probabilities,
                         subject_keys=subject_keys, transform=None if augmentation is not None else transform,
                         augmentation=augmentation, **device_options)
