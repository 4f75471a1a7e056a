"""Drop-in module path of the reference (src/configurations/config.py)."""
from text_similarity_b200.config import Configuration, ModelParameters, SearchConfiguration  # noqa: F401
