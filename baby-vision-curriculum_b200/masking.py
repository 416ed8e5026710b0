"""Host-side mirror of the reference's mask samplers (pretraining/generative/mask.py:3-46): same class names,
constructor arguments, numpy-global-RNG behaviour and float64 {0,1} output, so the reference loop's
`mask_gen = TubeMaskingGenerator((8,14,14), 0.9)` / `bool_masked[i,:] = mask_gen()` lines run unchanged."""
import numpy as np
import torch


class TubeMaskingGenerator:
    def __init__(self, input_size, mask_ratio):
        self.frames, self.height, self.width = input_size
        self.num_patches_per_frame = self.height * self.width
        self.total_patches = self.frames * self.num_patches_per_frame
        self.num_masks_per_frame = int(mask_ratio * self.num_patches_per_frame)
        self.total_masks = self.frames * self.num_masks_per_frame

    def __repr__(self):
        return "Maks: total patches {}, mask patches {}".format(self.total_patches, self.total_masks)

    def __call__(self):
        per_frame = np.hstack([np.zeros(self.num_patches_per_frame - self.num_masks_per_frame),
                               np.ones(self.num_masks_per_frame)])
        np.random.shuffle(per_frame)  # numpy's global RNG, like the reference (which never seeds it)
        return np.tile(per_frame, (self.frames, 1)).flatten()


class RandomMaskingGenerator:
    def __init__(self, input_size, mask_ratio):
        self.frames, self.height, self.width = input_size
        self.total_patches = self.frames * self.height * self.width
        self.total_masks = int(mask_ratio * self.total_patches)

    def __repr__(self):
        return "Maks: total patches {}, mask patches {}".format(self.total_patches, self.total_masks)

    def __call__(self):
        mask = np.hstack([np.zeros(self.total_patches - self.total_masks), np.ones(self.total_masks)])
        np.random.shuffle(mask)
        return mask


def batch_masks(mask_gen, batch_size):
    """pretrain_videomae.py:294-297: one generator call per sample into a float64 array, then .bool()."""
    out = np.zeros((batch_size, mask_gen.total_patches))
    for i in range(batch_size):
        out[i, :] = mask_gen()
    return torch.from_numpy(out).bool()
