"""Restatement of the intensity augmentation chain of the training scripts (test oracle; see oracle/__init__.py).

PARITY UNPINNED.  The chain (examples/train_seg.py:82-86, train_ldmks.py:82-84) is composed from ``batchgenerators``
(requirements.txt:7, no version pinned), a third-party package that is neither vendored in /root/reference nor installed
in this image, so no output of the real library could be generated here.  What follows restates the published algorithm
of ``batchgenerators.transforms.color_transforms`` / ``augmentations.color_augmentations`` (BrightnessTransform ->
augment_brightness_additive, GammaTransform -> augment_gamma, ContrastAugmentationTransform -> augment_contrast) with the
arguments of the reference's call sites and the library defaults for the rest (per_channel=True / p_per_channel=1 for
brightness, per_channel=False / invert_image=False / retain_stats=False / epsilon=1e-7 for gamma, per_channel=True /
preserve_range=True for contrast, p_per_sample=1 everywhere), in float32 as the dataset hands the patch over
(midasmednet/dataset.py:315-317), with one np.random draw where the library makes one.
"""
from __future__ import annotations

import numpy as np

BRIGHTNESS = dict(mu=0.0, sigma=0.3)                 # train_seg.py:84
GAMMA_RANGE = (0.7, 1.3)                             # train_seg.py:85
CONTRAST_RANGE = (0.3, 1.7)                          # train_seg.py:86


def _two_sided(rng_range):
    """The library's way of drawing a factor around 1: half of the draws below 1 when the range allows it."""
    if np.random.random() < 0.5 and rng_range[0] < 1:
        return np.random.uniform(rng_range[0], 1)
    return np.random.uniform(max(rng_range[0], 1), rng_range[1])


def brightness_additive(sample, mu, sigma, p_per_channel=1.0):
    for c in range(sample.shape[0]):
        if np.random.uniform() <= p_per_channel:
            sample[c] += np.random.normal(mu, sigma)
    return sample


def gamma(sample, gamma_range, epsilon=1e-7):
    g = _two_sided(gamma_range)
    minm = sample.min()
    rnge = sample.max() - minm
    return np.power((sample - minm) / np.float32(rnge + np.float32(epsilon)), np.float32(g)) * rnge + minm


def contrast(sample, contrast_range):
    for c in range(sample.shape[0]):
        mn = sample[c].mean()
        minm, maxm = sample[c].min(), sample[c].max()
        factor = np.float32(_two_sided(contrast_range))
        sample[c] = (sample[c] - mn) * factor + mn
        sample[c][sample[c] < minm] = minm
        sample[c][sample[c] > maxm] = maxm
    return sample


def augment_patch(data, brightness=BRIGHTNESS, gamma_range=GAMMA_RANGE, contrast_range=CONTRAST_RANGE, p_per_sample=1.0):
    """data: (C, H, W, D) float32, one patch (the dataset calls the chain with a batch of one, dataset.py:332-341).
    Returns the augmented float32 patch; draws from the global NumPy state in the library's order."""
    x = np.array(data, dtype=np.float32, copy=True)
    if np.random.uniform() < p_per_sample:
        x = brightness_additive(x, brightness["mu"], brightness["sigma"])
    if np.random.uniform() < p_per_sample:
        x = gamma(x, gamma_range).astype(np.float32)
    if np.random.uniform() < p_per_sample:
        x = contrast(x, contrast_range)
    return x
