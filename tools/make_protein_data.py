"""
One-off generator for phylo_utils_b200/data/*.dat.

The empirical amino-acid models (LG, WAG, JTT, Dayhoff) are published constants; the
reference keeps them as numpy literals in phylo_utils/data.py:4-101.  To guarantee that
both sides use the very same numbers, this script reads the reference's arrays (it only
runs in the build container, where /root/reference is mounted) and re-emits them in the
customary PAML ``.dat`` form: 19 lower-triangle rows of exchangeabilities followed by the
20 equilibrium frequencies, amino-acid order ARNDCQEGHILKMFPSTWYV.

    python tools/make_protein_data.py
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/phylo_utils/data.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "phylo_utils_b200", "data")


def main():
    spec = importlib.util.spec_from_file_location("_ref_data", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("lg", "wag", "jtt", "dayhoff"):
        rates = np.asarray(getattr(mod, name + "_rates"), dtype=np.double)
        freqs = np.asarray(getattr(mod, name + "_freqs"), dtype=np.double)
        assert rates.shape == (20, 20) and freqs.shape == (20,)
        assert np.array_equal(rates, rates.T), name
        path = os.path.join(OUT, name + ".dat")
        with open(path, "w") as fh:
            for i in range(1, 20):
                fh.write(" ".join(repr(float(rates[i, j])) for j in range(i)) + "\n")
            fh.write("\n")
            fh.write(" ".join(repr(float(f)) for f in freqs) + "\n")
        print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
