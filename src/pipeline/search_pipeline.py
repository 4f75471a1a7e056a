"""Drop-in module path of the reference (src/pipeline/search_pipeline.py): re-exports the
B200-native pipelines."""
from text_similarity_b200.pipeline import (APISearchPipeline, Pipeline, SearchPipeline,  # noqa: F401
                                           SemanticSearchPipeline, SentenceMiningPipeline)
