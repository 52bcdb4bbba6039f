from .deepspeech_trainer import CTCLoss, DeepSpeechStep, fit

__all__ = ["CTCLoss", "DeepSpeechStep", "fit"]
