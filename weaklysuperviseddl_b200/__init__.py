"""weaklysuperviseddl_b200 -- B200-native (sm_100a) drop-in for the weak-supervision hot path of
alexncoleman/WeaklySupervisedDL: LayerCAM -> pseudo-mask -> pairwise cut / boundary regularisers.

Module names mirror the reference's TraditionalModel/ files so `from LayerCAM import LayerCAMGenerator`
becomes `from weaklysuperviseddl_b200.LayerCAM import LayerCAMGenerator`.  All arithmetic runs in
libwsdl_b200.so (include/wsdl_b200.h); importing this package does not need a GPU, calling it does."""
from . import _native  # noqa: F401
from .AlternatingDirectionBoundaryLoss import ConstrainToBoundaryLossSingle  # noqa: F401
from .AlternatingDirectionCutLoss import (  # noqa: F401
    LocalNormalizedCutLoss, compute_affinities, refine_pseudo_mask, refine_pseudo_masks_batched)
from .ExtraUtilities import compute_iou_and_acc, compute_iou_and_acc_batched  # noqa: F401
from .LayerCAM import LayerCAMGenerator, evaluate_layercam_on_test_set  # noqa: F401
from .PsuedoMasks import generate_pseudo_masks, generate_pseudo_masks_sharded, keep_largest  # noqa: F401
from .WeakSupervisionLoss import WeakSupervisionLoss  # noqa: F401

__version__ = "0.1.0"
