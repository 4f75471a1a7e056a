"""Drop-in module path of the reference (src/modules/modules.py): the poolers on the search path."""
from text_similarity_b200.pooling import AvgPoolingStrategy, LearningStrategy, PoolingStrategy  # noqa: F401
