"""lcasr_b200 — B200-native (sm_100a) drop-in for the encoder + CTC hot path of
robflynnyh/long-context-asr: ``SCConformerXL`` (lcasr/models/sconformer_xl.py), ``GreedyCTCDecoder``
(lcasr/decoding/greedy.py) and the ``torch.nn.CTCLoss`` call sites (exp/train.py:104,249).

Importing this package loads ``liblcasr_b200.so`` (hand-written CUDA behind a C ABI, see
include/lcasr_b200.h) and fails loudly if it has not been built.  There is no CPU fallback.
"""
from . import _lib, ops, seqpar, train_ops, training, longform, optim, frontend, augmentation
from .model import SCConformerXL, ConformerLayer, RMSNorm, BatchRenorm1d
from .decoding import GreedyCTCDecoder
from .losses import CTCLoss

__all__ = ["SCConformerXL", "ConformerLayer", "RMSNorm", "BatchRenorm1d", "GreedyCTCDecoder", "CTCLoss", "ops", "seqpar"]
__version__ = "0.1.0"
