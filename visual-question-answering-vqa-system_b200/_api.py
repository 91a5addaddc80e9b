"""Public names of the package (import-light: the CUDA library is loaded on first use)."""
from .batcher import MicroBatcher
from .inference import VQAInference, get_inference_engine
from .model import VQAModel, create_vqa_model, load_vqa_model
from .text import AnswerVocabulary, Tokenizer

__all__ = ["VQAModel", "create_vqa_model", "load_vqa_model", "VQAInference", "get_inference_engine",
           "Tokenizer", "AnswerVocabulary", "MicroBatcher"]
