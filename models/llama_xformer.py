"""models.llama_xformer -- see seed_b200/llama.py (mirror of the reference module of this name)."""
from seed_b200.llama import LlamaForCausalLM  # noqa: F401
