"""seed_b200 -- Blackwell (sm_100a) implementation of the SEED visual-tokenizer encode path and the
llama_xformer forward path behind a C ABI (include/seedb200.h).  See DESIGN.md."""
__version__ = "0.1.0"
