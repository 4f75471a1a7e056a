"""B200-native semantic-search hot path (pool -> normalise -> cosine -> top-k -> sharded merge)
behind the Python surface of cr1m5onk1ng/text_similarity.  See DESIGN.md."""
from . import _lib  # noqa: F401

__all__ = ["ops", "_lib"]
