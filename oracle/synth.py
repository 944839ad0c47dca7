"""Re-export of seed_b200/synth.py (deterministic synthetic weights / inputs) for the oracle-side scripts."""
from seed_b200.synth import *  # noqa: F401,F403
from seed_b200.synth import CODEBOOK_STD, encoder_state_dict, images, llama_state_dict, prompt_ids  # noqa: F401
