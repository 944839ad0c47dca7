"""models.model_tools -- see seed_b200/llama.py (mirror of the reference module of this name)."""
from seed_b200.llama import get_pretrained_llama_causal_model  # noqa: F401
