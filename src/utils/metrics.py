"""Drop-in module path of the reference (src/utils/metrics.py): cos_sim, its fused top-k form and the
retrieval-accuracy meter built on it."""
from text_similarity_b200.metrics import AverageMeter, RetrievalAccuracyMeter, cos_sim, cos_sim_topk  # noqa: F401
