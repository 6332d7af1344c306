"""superpoint_transformer_b200 — B200-native (sm_100a) hot path of the Superpoint
Transformer: superpoint-graph self-attention + segment pooling, behind the
reference's SPT / Stage / SelfAttentionBlock / *Pool / Data / NAG / Cluster API.

Host code is Python/PyTorch (plumbing); the arithmetic is hand-written CUDA in
libspt_b200.so (C ABI: include/spt_b200.h), loaded through ctypes.  Importing the
package does not need a GPU; calling any op does, and fails loudly otherwise.
"""
from . import ops  # noqa: F401
from . import nn  # noqa: F401
from . import data  # noqa: F401
from . import transforms  # noqa: F401
from .data import Data, Batch, NAG, NAGBatch, Cluster, CSRData  # noqa: F401
from .spt import SPT  # noqa: F401
from .utils.nn import init_weights  # noqa: F401

__version__ = '0.1.0'
