from .config import ModelConfig, TextConfig, VisionConfig
from .language import LanguageModel
from .llava import Model
from .vision import VisionModel
