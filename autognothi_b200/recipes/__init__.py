"""Host-side mirror of the reference's `recipes/` plugin boundary (reference recipes/types.py:96-162)."""
