from .config import ModelConfig, TextConfig, VisionConfig
from .idefics3 import Model
from .language import LanguageModel
from .vision import VisionModel
