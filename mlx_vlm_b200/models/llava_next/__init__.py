from .config import ModelConfig, TextConfig, VisionConfig
from .language import LanguageModel
from .llava_next import Model
from .vision import VisionModel
