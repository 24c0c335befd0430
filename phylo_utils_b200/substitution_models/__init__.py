from .dna import JC69, K80, F81, F84, HKY85, TN93, GTR, Strsym, Unrest
from .protein import LG, WAG, JTT, Dayhoff
from .codon import GY94
from .abstract import Model, Eigen

__all__ = ['JC69', 'K80', 'F81', 'F84', 'HKY85', 'TN93', 'GTR', 'Strsym', 'Unrest',
           'LG', 'WAG', 'JTT', 'Dayhoff', 'GY94']
