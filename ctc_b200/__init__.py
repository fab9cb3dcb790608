"""ctc_b200 -- B200-native (sm_100a) no-blank CTC losses, drop-in for gotaku6629/CTC's
``NoBlankCTC`` / ``NoBlankBinaryCTC`` (the one hot path this repository accelerates).

CUDA only: importing is cheap, but the first call loads ``ctc_b200/libnbctc.so`` and raises if
it has not been built (``python -m ctc_b200.build``).  There is no CPU fallback.
"""
from ._ffi import FLAG_DEFAULT, FLAG_GENERIC, FLAG_LOCKSTEP, FLAG_NO_GRAD, FLAG_SEQWARP, FLAG_SUM_WEIGHTED, NbctcError, launch_count  # noqa: F401
from .function import best_path, ctc_plus_cross_entropy, no_blank_binary_ctc_loss, no_blank_ctc_loss  # noqa: F401
from .modules import NoBlankBinaryCTC, NoBlankCTC  # noqa: F401
from .dist import ShardedLoss, all_reduce_sum, all_reduce_sum_async, shard_batch  # noqa: F401

__all__ = [
    "NoBlankCTC", "NoBlankBinaryCTC", "no_blank_ctc_loss", "no_blank_binary_ctc_loss", "best_path", "ctc_plus_cross_entropy",
    "ShardedLoss", "all_reduce_sum", "all_reduce_sum_async", "shard_batch", "NbctcError", "launch_count",
    "FLAG_DEFAULT", "FLAG_GENERIC", "FLAG_LOCKSTEP", "FLAG_NO_GRAD", "FLAG_SEQWARP",
]
