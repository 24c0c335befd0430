"""
Substitution models.  First block mirrors /root/reference/tests/test_substitution_models.py
(same instances, same assertions); second block pins Q, frequencies and P / dP / d2P against values
produced by the reference itself (tests/golden/models.npz).
"""
import numpy as np
import pytest

from phylo_utils_b200.substitution_models import (JC69, K80, F81, F84, HKY85, TN93, GTR, Strsym, Unrest, WAG, LG, JTT,
                                                  Dayhoff, GY94)
from phylo_utils_b200.substitution_models import utils as smu
from phylo_utils_b200.substitution_models.abstract import Model
from helpers import load

F4 = [0.1, 0.2, 0.3, 0.4]
UNREST_RATES = [[0., 1., 2., 3.], [4., 0., 5., 6.], [7., 8., 0., 9.], [10., 11., 12., 0.]]
REVERSIBLE = {
    "JC69": lambda: JC69(), "K80": lambda: K80(1.5), "F81": lambda: F81(F4), "F84": lambda: F84(1.5, F4),
    "HKY85": lambda: HKY85(1.5, F4), "TN93": lambda: TN93(2.5, 2.4, freqs=F4),
    "GTR": lambda: GTR([6., 5., 4., 3., 2., 1.], F4), "WAG": lambda: WAG(), "LG": lambda: LG(), "JTT": lambda: JTT(),
    "Dayhoff": lambda: Dayhoff(),
}
NONREV = {"Strsym": lambda: Strsym([1., 2., 3., 4., 5., 6.]), "Unrest": lambda: Unrest(rates=UNREST_RATES)}


def test_check_frequencies():
    good = np.array([0.25, 0.25, 0.25, 0.25])
    with pytest.raises(ValueError):
        smu.check_frequencies(good, 5)
    with pytest.raises(ValueError):
        smu.check_frequencies(np.array([0.250001, 0.25, 0.25, 0.25]), 4)
    with pytest.raises(ValueError):
        smu.check_frequencies(np.array([-0.25, 0.75, 0.25, 0.25]), 4)
    assert np.allclose(good, smu.check_frequencies(good, 4))


def test_check_rates():
    with pytest.raises(ValueError):
        smu.check_rates(np.ones((3, 4)), 4)
    with pytest.raises(ValueError):
        smu.check_rates(-np.ones((4, 4)), 4)
    with pytest.raises(ValueError):
        smu.check_rates(np.triu(np.ones((4, 4))), 4)
    smu.check_rates(np.triu(np.ones((4, 4))), 4, symmetry=False)


@pytest.mark.parametrize("name", sorted(REVERSIBLE))
def test_reversible_detailed_balance_and_scale(name):
    m = REVERSIBLE[name]()
    assert m.detailed_balance()
    assert abs(m.freqs.T.dot(-np.diag(m.q())) - 1.0) < 1e-7
    e = m.eigen
    assert np.allclose((e.evecs * e.evals).dot(e.ivecs), m.q(), atol=1e-12)
    assert e.ivecs.flags.f_contiguous and e.evecs.flags.c_contiguous


@pytest.mark.parametrize("name", sorted(NONREV))
def test_nonreversible(name):
    m = NONREV[name]()
    assert not m.detailed_balance()
    assert abs(m.freqs.T.dot(-np.diag(m.q())) - 1.0) < 1e-7
    assert not m.has_real_eigensystem


def test_strsym_frequency_constraints():
    m = NONREV["Strsym"]()
    assert abs(m.freqs[0] - m.freqs[3]) < 1e-12 and abs(m.freqs[1] - m.freqs[2]) < 1e-12


@pytest.mark.parametrize("name", sorted(list(REVERSIBLE) + list(NONREV)))
def test_against_reference_values(name):
    g = load("models")
    m = (REVERSIBLE.get(name) or NONREV[name])()
    t, rates = float(g["t"]), g["rates"]
    assert np.allclose(m.q(), g[name + "_q"], rtol=1e-13, atol=1e-15)
    assert np.allclose(m.freqs, g[name + "_freqs"], rtol=1e-13, atol=1e-15)
    # P is invariant to the eigenvector sign / ordering ambiguity of LAPACK; compare P, never V
    assert np.allclose(m.p(t, rates), g[name + "_p"], rtol=1e-11, atol=1e-14)
    assert np.allclose(m.dp_dt(t, rates), g[name + "_dp"], rtol=1e-10, atol=1e-13)
    assert np.allclose(m.d2p_dt2(t, rates), g[name + "_d2p"], rtol=1e-10, atol=1e-12)
    assert m.p(t, rates).flags.c_contiguous and m.p(t, rates).shape == (4, m.size, m.size)
    if name == "JC69":
        assert np.allclose(m.p(t), g["JC69_p_closed"], rtol=1e-14)
        assert np.allclose(Model.p(m, t, rates), m.p(t, rates), atol=1e-15)


def test_dp_dt_convention_is_derivative_in_scaled_time():
    # reference quirk (abstract.py:61-77): no r_k chain-rule factor
    m, rates, t, h = REVERSIBLE["GTR"](), np.array([0.5, 2.0]), 0.2, 1e-6
    fd = (m.p(t + h, rates) - m.p(t - h, rates)) / (2 * h)
    assert np.allclose(fd, m.dp_dt(t, rates) * rates[:, None, None], atol=1e-8)


def test_p_rows_sum_to_one_and_p0_is_identity():
    for name, make in REVERSIBLE.items():
        m = make()
        p = m.p(0.37, [0.1, 1.0, 3.0])
        assert np.allclose(p.sum(axis=2), 1.0, atol=1e-12), name
        assert np.allclose(m.p(0, [1.0])[0], np.eye(m.size), atol=1e-12), name


def test_gtr_input_forms():
    a = GTR([6., 5., 4., 3., 2., 1.], F4)
    five = [6., 5., 4., 3., 2.]
    b = GTR(five, F4)
    assert five == [6., 5., 4., 3., 2.]          # caller's list is not mutated
    assert np.allclose(a.q(), b.q())
    c = GTR(a.rates, F4)                          # 4x4 matrix accepted
    assert np.allclose(a.q(), c.q())
    assert np.allclose(GTR().q(), JC69().q())


def test_gy94_structure():
    rng = np.random.default_rng(0)
    pf = rng.dirichlet(np.ones(4) * 5, size=3)
    from phylo_utils_b200.substitution_models.codon import f3x4, SENSE_CODONS, GENETIC_CODE
    m = GY94(2.0, 0.2, f3x4(pf))
    assert m.size == 61 and len(SENSE_CODONS) == 61 and "TAA" not in SENSE_CODONS
    assert m.detailed_balance()
    assert abs(m.freqs.dot(-np.diag(m.q())) - 1.0) < 1e-10
    q = m.q()
    i, j = SENSE_CODONS.index("AAA"), SENSE_CODONS.index("AAG")      # Lys -> Lys, transition, synonymous
    k = SENSE_CODONS.index("AAC")                                    # Lys -> Asn, transversion, non-synonymous
    l = SENSE_CODONS.index("CCC")                                    # three differences
    assert GENETIC_CODE["AAA"] == GENETIC_CODE["AAG"] != GENETIC_CODE["AAC"]
    assert np.isclose(q[i, j] / m.freqs[j], 2.0 * q[i, k] / m.freqs[k] / 0.2)
    assert q[i, l] == 0
    p = m.p(0.3, [0.5, 1.5])
    assert np.allclose(p.sum(axis=2), 1.0, atol=1e-11) and p.min() > -1e-12


# The standard genetic code written out per amino acid (NCBI translation table 1) - deliberately NOT the packed
# 64-character string codon.py uses, so that the two encodings check each other.
_CODE_BY_AA = {
    "F": "TTT TTC", "L": "TTA TTG CTT CTC CTA CTG", "I": "ATT ATC ATA", "M": "ATG", "V": "GTT GTC GTA GTG",
    "S": "TCT TCC TCA TCG AGT AGC", "P": "CCT CCC CCA CCG", "T": "ACT ACC ACA ACG", "A": "GCT GCC GCA GCG",
    "Y": "TAT TAC", "H": "CAT CAC", "Q": "CAA CAG", "N": "AAT AAC", "K": "AAA AAG", "D": "GAT GAC", "E": "GAA GAG",
    "C": "TGT TGC", "W": "TGG", "R": "CGT CGC CGA CGG AGA AGG", "G": "GGT GGC GGA GGG",
}


def _textbook_gy94(kappa, omega, pi, codons):
    """Goldman & Yang (1994) / Yang (2006, eq. 2.7) rate matrix built entry by entry from the definition."""
    aa = {cod: a for a, cods in _CODE_BY_AA.items() for cod in cods.split()}
    assert len(aa) == 61 and sorted(aa) == sorted(codons)
    purines = set("AG")
    n = len(codons)
    q = np.zeros((n, n))
    for i, ci in enumerate(codons):
        for j, cj in enumerate(codons):
            where = [p for p in range(3) if ci[p] != cj[p]]
            if i == j or len(where) != 1:
                continue
            x, y = ci[where[0]], cj[where[0]]
            rate = pi[j]
            if (x in purines) == (y in purines):       # purine <-> purine or pyrimidine <-> pyrimidine: transition
                rate *= kappa
            if aa[ci] != aa[cj]:
                rate *= omega
            q[i, j] = rate
        q[i, i] = -q[i].sum()
    return q / -(pi * np.diag(q)).sum()                # one expected substitution per codon per unit time


def test_gy94_q_and_p_against_the_textbook_definition():
    """Independent pin for the codon model (it has no counterpart in the reference): Q from the published definition,
    P = expm(Q t r) from scipy, against GY94.q() and the eigen-decomposition route of Model.p."""
    from scipy.linalg import expm
    from phylo_utils_b200.substitution_models.codon import f3x4, SENSE_CODONS
    rng = np.random.default_rng(4)
    for kappa, omega in ((2.0, 0.2), (1.0, 1.0), (5.5, 1.7)):
        pi = f3x4(rng.dirichlet(np.ones(4) * 5, size=3))
        m = GY94(kappa, omega, pi)
        q = _textbook_gy94(kappa, omega, pi, SENSE_CODONS)
        assert np.allclose(m.q(), q, rtol=1e-13, atol=1e-16)
        rates = np.array([0.03, 0.25, 0.82, 2.89])
        for t in (0.01, 0.3, 2.0):
            want = np.stack([expm(q * t * r) for r in rates])
            got = m.p(t, rates)
            assert np.allclose(got, want, rtol=1e-9, atol=1e-13)
    # f3x4 itself: product of the position frequencies over the sense codons, renormalised
    pf = rng.dirichlet(np.ones(4) * 5, size=3)
    want = np.array([pf[0]["ACGT".index(c[0])] * pf[1]["ACGT".index(c[1])] * pf[2]["ACGT".index(c[2])] for c in SENSE_CODONS])
    assert np.allclose(f3x4(pf), want / want.sum(), rtol=1e-14)
