"""Alphabet tags, same integer values as /root/reference/phylo_utils/alignment/alphabets.py:1-3 (CODON is new)."""
DNA = 0
PROTEIN = 1
BINARY = 2
CODON = 3
