"""Drop-in module path of the reference (src/pipeline/clustering.py)."""
from text_similarity_b200.ranking import ClusteringPipeline  # noqa: F401
