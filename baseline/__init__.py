"""The reference arm of bench.py: install recipe and timing driver for the UNMODIFIED reference (baseline/_ref)."""
