from .greedy_decoder import Decoder, GreedyDecoder

__all__ = ["Decoder", "GreedyDecoder"]
