"""
Empirical amino-acid models (reference: substitution_models/lg.py:6-15 and the identical
wag.py / jtt.py / dayhoff.py): fixed exchangeabilities, frequencies either the model's own or
user supplied ("+F").
"""
from .. import data
from .abstract import Eigen, ProteinModel
from .utils import check_frequencies, compute_q_matrix, get_eigen


class _Empirical(ProteinModel):
    _default_freqs = None

    def __init__(self, freqs=None, rates=None):
        # ``rates`` is accepted and ignored so that the CLI's uniform
        # ``cls(rates=..., freqs=...)`` call (reference: bin/phy.py:129) works for protein models too
        if freqs is None:
            self._freqs = self._default_freqs.copy()
        else:
            self._freqs = check_frequencies(freqs, self.size)
        self._q_mtx = compute_q_matrix(self._rates, self._freqs)
        self.eigen = Eigen(*get_eigen(self._q_mtx, self._freqs))


class LG(_Empirical):
    _name = 'LG'
    _rates = data.lg_rates.copy()
    _default_freqs = data.lg_freqs


class WAG(_Empirical):
    _name = 'WAG'
    _rates = data.wag_rates.copy()
    _default_freqs = data.wag_freqs


class JTT(_Empirical):
    _name = 'JTT'
    _rates = data.jtt_rates.copy()
    _default_freqs = data.jtt_freqs


class Dayhoff(_Empirical):
    _name = 'Dayhoff'
    _rates = data.dayhoff_rates.copy()
    _default_freqs = data.dayhoff_freqs
