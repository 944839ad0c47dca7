"""models.seed_qformer.qformer_quantizer -- see seed_b200/qformer_quantizer.py."""
from seed_b200.qformer_quantizer import Blip2QformerQuantizer  # noqa: F401
