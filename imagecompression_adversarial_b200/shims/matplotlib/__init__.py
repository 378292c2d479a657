"""Inert stand-in for the plotting imports of the reference scripts (attack_rd.py:18); no hot-path use."""
