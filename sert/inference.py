"""Drop-in for the reference's sert/inference.py."""
from sert_b200.inference import create, WordBatcher, EmbeddingMapper, aggregate_distribution  # noqa: F401
