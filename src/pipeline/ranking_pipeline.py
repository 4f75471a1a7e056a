"""Drop-in module path of the reference (src/pipeline/ranking_pipeline.py)."""
from text_similarity_b200.ranking import RankingPipeline  # noqa: F401
