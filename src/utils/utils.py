"""Drop-in module path of the reference (src/utils/utils.py): the similarity helper on the search path."""
from text_similarity_b200.ranking import most_similar_vectors  # noqa: F401
