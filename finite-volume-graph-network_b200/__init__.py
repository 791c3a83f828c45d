"""fvgn_b200 — B200-native implementation of FVGN's encode-process-decode message-passing step
(Litianyu141/Finite-Volume-Graph-Network) behind the reference's own module API.  See DESIGN.md.

    from fvgn_b200 import FVGN, Data, Batch, dataset
The arithmetic lives in libfvgn_b200.so (sm_100a CUDA, C-ABI in include/fvgn_b200.h); there is no CPU fallback."""
from . import _lib, ops, synthetic, parallel, domain  # noqa: F401
from .data import Data, Batch  # noqa: F401
from .topology import MeshTopology, topology_of  # noqa: F401
from .FVGN_model import (FVGN, EncoderProcesserDecoder, Encoder, GnBlock, Decoder, Intergrator, build_mlp, CellBlock, EdgeBlock)  # noqa: F401
from .utils.normalization import Normalizer  # noqa: F401
from .utils.utilities import NodeType  # noqa: F401
from .dataset import Load_mesh as dataset  # noqa: F401
from .utils.loss_compute import compute_FVM_loss  # noqa: F401
from .rollout import GraphedRollout  # noqa: F401
from .training import GraphedTrainStep  # noqa: F401
from .renumber import renumber_mesh, inverse_permutation  # noqa: F401

__version__ = "0.1.0"
