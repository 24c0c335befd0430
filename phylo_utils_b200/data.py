"""
Static model constants under the reference's names (/root/reference/phylo_utils/data.py:4-112):
``lg_rates, lg_freqs, wag_*, jtt_*, dayhoff_*`` (read-only 20x20 / 20 arrays, amino-acid order
ARNDCQEGHILKMFPSTWYV) and the equal nucleotide rates / frequencies.

The empirical matrices live next to this file as PAML-style ``.dat`` text
(tools/make_protein_data.py wrote them from the reference's own arrays, so the numbers are
identical).
"""
import os

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _read_paml_dat(name):
    with open(os.path.join(_HERE, name + ".dat")) as fh:
        values = [float(tok) for tok in fh.read().split()]
    n = 20
    n_tri = n * (n - 1) // 2
    if len(values) != n_tri + n:
        raise ValueError("{}.dat: expected {} numbers, found {}".format(name, n_tri + n, len(values)))
    rates = np.zeros((n, n), dtype=np.double)
    rates[np.tril_indices(n, -1)] = values[:n_tri]
    rates = np.ascontiguousarray(rates + rates.T)
    freqs = np.ascontiguousarray(values[n_tri:], dtype=np.double)
    rates.setflags(write=False)
    freqs.setflags(write=False)
    return rates, freqs


lg_rates, lg_freqs = _read_paml_dat("lg")
wag_rates, wag_freqs = _read_paml_dat("wag")
jtt_rates, jtt_freqs = _read_paml_dat("jtt")
dayhoff_rates, dayhoff_freqs = _read_paml_dat("dayhoff")

fixed_equal_nucleotide_rates = np.ascontiguousarray(np.ones((4, 4)) - np.eye(4))
fixed_equal_nucleotide_rates.setflags(write=False)
fixed_equal_nucleotide_frequencies = np.full(4, 0.25)
fixed_equal_nucleotide_frequencies.setflags(write=False)
