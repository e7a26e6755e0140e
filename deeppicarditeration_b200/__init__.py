"""deeppicarditeration_b200 — B200-native (sm_100a) Monte Carlo label generator for Deep Picard Iteration.

Only the hot path of the reference (picard/data.py label generation and what it calls) lives here, behind the
reference's own Python interfaces; see DESIGN.md.  The CUDA library is mandatory: nothing falls back to CPU.
"""
__version__ = "0.1.0"
