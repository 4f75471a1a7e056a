"""Drop-in module path of the reference (src/evaluation/eval_sentence_mining.py): the pipeline-agreement check."""
from text_similarity_b200.ranking import compare_models  # noqa: F401
