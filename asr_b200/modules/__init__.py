from .blocks import BatchRNN, InferenceBatchSoftmax, Lookahead, MaskConv, SequenceWise
from .deepspeech import DeepSpeech

__all__ = ["BatchRNN", "InferenceBatchSoftmax", "Lookahead", "MaskConv", "SequenceWise", "DeepSpeech"]
