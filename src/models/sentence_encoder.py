"""Drop-in module path of the reference (src/models/sentence_encoder.py)."""
from text_similarity_b200.encoder import (BaseEncoderModel, OnnxSentenceTransformerWrapper,  # noqa: F401
                                          SentenceTransformerWrapper)
