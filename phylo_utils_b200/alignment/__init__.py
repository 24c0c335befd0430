from . import alphabets, charmaps
from .alignment import (alignment_to_numpy, alignment_to_codes, compress_codes, compress_codes_gpu, seq_to_partials,
                        seq_to_codes, invariant_sites, read_alignment, codebook_for, SeqRecord)
