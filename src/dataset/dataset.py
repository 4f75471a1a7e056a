"""Drop-in module path of the reference (src/dataset/dataset.py): the feature carrier of the path."""
from text_similarity_b200.features import EmbeddingsFeatures  # noqa: F401
