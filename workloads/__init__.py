"""Synthetic workloads for tests, parity tools and bench.py (test / measurement infrastructure, not product code):
seeded Khmer text-line images composed from a committed bank of rendered pseudo-words."""
