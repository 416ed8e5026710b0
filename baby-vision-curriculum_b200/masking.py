"""Host-side mirror of the reference's mask samplers (pretraining/generative/mask.py:3-46): same class names,
constructor arguments, attributes, numpy-global-RNG consumption and float64 {0,1} output, so the reference loop's
`mask_gen = TubeMaskingGenerator((8,14,14), 0.9)` / `bool_masked[i,:] = mask_gen()` lines run unchanged.  Both samplers
are one shuffled {0,1} vector; they differ in what is shuffled (one frame's patches, tiled over time -- or all patches)."""
import numpy as np
import torch


class _ShuffledMask:
    def __init__(self, input_size):
        self.frames, self.height, self.width = input_size

    @staticmethod
    def _draw(n, n_masked):
        """n - n_masked zeros followed by n_masked ones, shuffled in place by numpy's GLOBAL generator (which the
        reference uses and never seeds): one np.random.shuffle of an n-vector per call."""
        v = np.zeros(n)
        v[n - n_masked:] = 1.0
        np.random.shuffle(v)
        return v

    def __repr__(self):
        return f"Maks: total patches {self.total_patches}, mask patches {self.total_masks}"  # (sic)


class TubeMaskingGenerator(_ShuffledMask):
    """mask.py:3-24: the same spatial pattern in every temporal slot (a "tube")."""

    def __init__(self, input_size, mask_ratio):
        super().__init__(input_size)
        self.num_patches_per_frame = self.height * self.width
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_patches = self.frames * self.num_patches_per_frame
        self.total_masks = self.frames * self.num_masks_per_frame

    def __call__(self):
        frame = self._draw(self.num_patches_per_frame, self.num_masks_per_frame)
        return np.tile(frame, (self.frames, 1)).flatten()


class RandomMaskingGenerator(_ShuffledMask):
    """mask.py:26-46: every spatio-temporal patch masked independently of its neighbours in time."""

    def __init__(self, input_size, mask_ratio):
        super().__init__(input_size)
        self.total_patches = self.frames * self.height * self.width
        self.total_masks = int(mask_ratio * self.total_patches)

    def __call__(self):
        return self._draw(self.total_patches, self.total_masks)


def batch_masks(mask_gen, batch_size):
    """pretrain_videomae.py:294-297: one generator call per sample into a float64 array, then .bool()."""
    rows = np.stack([mask_gen() for _ in range(batch_size)])
    return torch.from_numpy(rows).bool()
