"""models.seed_llama_tokenizer -- see seed_b200/tokenizer.py (mirror of the reference module of this name)."""
from seed_b200.tokenizer import (DIFFUSION_NAME, WEIGHTS_NAME, ImageTokenizer, SeedImageTokenMixin,  # noqa: F401
                                 SeedLlamaTokenizer, all_gather_ids)
