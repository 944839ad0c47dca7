"""Drop-in replacements for the reference's `models` package (same dotted paths as the hydra `_target_`
strings in /root/reference/configs/**): each module re-exports the seed_b200 mirror of the reference module of
the same name."""
