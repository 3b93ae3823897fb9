"""llmseg_b200 — B200-native (sm_100a) implementation of the LLM-Seg inference forward path."""
__version__ = "0.1.0"
