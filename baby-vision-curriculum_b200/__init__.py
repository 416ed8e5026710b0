"""bvc-b200: the VideoMAE pretraining step of ssheybani/baby-vision-curriculum on hand-written sm_100a kernels.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it through the
`bvc_b200` shim at the repository root:  `import bvc_b200 as bvc`.
"""
from . import _lib  # noqa: F401
from ._lib import BvcError, load as load_library  # noqa: F401
from .modeling_videomae import (  # noqa: F401
    VideoMAEConfig,
    VideoMAEForPreTraining,
    VideoMAEForPreTrainingOutput,
    get_sinusoid_encoding_table,
)
from .masking import TubeMaskingGenerator, RandomMaskingGenerator, batch_masks  # noqa: F401
from .ddputils import AllReduce, AllGather  # noqa: F401
from .optim import FusedSGD, FusedAdam, FusedAdamW  # noqa: F401
from .ddp import DistributedDataParallel  # noqa: F401
from .graphed import GraphedTrainStep  # noqa: F401
from .simclr import info_nce_loss, get_special_matrix, make_masks as make_simclr_masks  # noqa: F401
from .jepa import (apply_masks, repeat_interleave_batch, jepa_targets, smooth_l1_loss, ema_update,  # noqa: F401
                   MaskCollator, update_masks)
from . import jepa_vit  # noqa: F401,E402  (Block / convert_blocks: the predictive path's ViT block on these kernels)
