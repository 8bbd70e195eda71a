from .config import ModelConfig, PerceiverConfig, TextConfig, VisionConfig
from .idefics2 import Model
from .language import LanguageModel
from .vision import VisionModel
