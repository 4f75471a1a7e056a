"""Drop-in module path of the reference (src/utils/metrics.py): cos_sim and its fused top-k form."""
from text_similarity_b200.metrics import cos_sim, cos_sim_topk  # noqa: F401
