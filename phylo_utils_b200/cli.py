"""
Command-line front end: ``phy -t tree.nwk -s aln.fasta -m "GTR{6.0,5.0,4.0,3.0,2.0,1.0}+F{0.1,0.2,0.3,0.4}+G4{0.5}"``
prints ``lnL = ...`` like the reference's bin/phy.py (lines 13-146), without dendropy / Biopython / pyparsing:
trees come from the built-in Newick reader, alignments from the FASTA / PHYLIP reader, and the model
string ``Model{params}+F{freqs}+G<ncat>{alpha}`` is parsed with a small regular-expression grammar
(same tokens and meaning as bin/phy.py:41-56).
"""
import argparse
import os
import re
import sys

from . import rate_models, substitution_models
from .alignment import alphabets
from .alignment.alignment import read_alignment
from .tree import Tree

MODELS = {
    'GTR': substitution_models.GTR, 'HKY': substitution_models.HKY85, 'HKY85': substitution_models.HKY85,
    'K80': substitution_models.K80, 'F81': substitution_models.F81, 'F84': substitution_models.F84,
    'JC': substitution_models.JC69, 'JC69': substitution_models.JC69, 'TN93': substitution_models.TN93,
    'LG': substitution_models.LG, 'WAG': substitution_models.WAG, 'JTT': substitution_models.JTT,
    'Dayhoff': substitution_models.Dayhoff,
}
PROTEIN_MODELS = ("JTT", "Dayhoff", "WAG", "LG")

_REAL = r"\d+\.\d+"
_LIST = r"\{\s*(" + _REAL + r"(?:\s*,\s*" + _REAL + r")*)\s*\}"
_TOKEN = re.compile(r"^(?:(?P<freq>F)(?:" + _LIST + r")?|(?P<gamma>G)(?P<ncat>\d+)(?:" + _LIST.replace("(", "(?P<alpha>", 1) + r")?)$")
_HEAD = re.compile(r"^(?P<name>[A-Za-z]+\d*)(?:" + _LIST + r")?$")


def _floats(text):
    return [float(v) for v in text.split(",")] if text else None


def _unpack(vals):
    if vals is None:
        return None
    return vals[0] if len(vals) == 1 else vals


def parse_model_string(model_string):
    """-> dict(subs_model, model_params, freq_params, rate_model, rate_cats, rate_param); raises ValueError on bad syntax."""
    parts = [p.strip() for p in model_string.split("+")]
    head = _HEAD.match(parts[0])
    if not head:
        raise ValueError("cannot parse substitution model in {!r}".format(model_string))
    out = dict(subs_model=head.group("name"), model_params=_floats(head.group(2)), freq_params=None, rate_model=None,
               rate_cats=None, rate_param=None)
    if len(parts) > 3:
        raise ValueError("at most two '+' components are allowed (frequencies and a rate model)")
    for comp in parts[1:]:
        m = _TOKEN.match(comp)
        if not m:
            raise ValueError("cannot parse model component {!r}".format(comp))
        if m.group("freq"):
            out["freq_params"] = _floats(m.group(2))
        else:
            out["rate_model"] = "G"
            out["rate_cats"] = int(m.group("ncat"))
            out["rate_param"] = _floats(m.group("alpha"))
    return out


def build_models(desc):
    name = desc["subs_model"]
    if name not in MODELS:
        raise ValueError("Unrecognised model {}. Valid options are {}".format(name, ", ".join(MODELS)))
    cls = MODELS[name]
    params, freqs = _unpack(desc["model_params"]), desc["freq_params"]
    if name in ("JC", "JC69"):
        model = cls()
    elif name in PROTEIN_MODELS:
        model = cls(freqs=freqs)
    elif name == "K80":
        model = cls(params if params is not None else 1.0)
    elif name == "F81":
        model = cls(freqs if freqs is not None else [0.25] * 4)
    elif name in ("HKY", "HKY85", "F84"):
        model = cls(params if params is not None else 1.0, freqs if freqs is not None else [0.25] * 4)
    elif name == "TN93":
        p = desc["model_params"] or [1.0, 1.0]
        model = cls(p[0], p[1], p[2] if len(p) > 2 else 1.0, freqs)
    else:
        model = cls(rates=desc["model_params"], freqs=freqs)
    if desc["rate_model"] == "G":
        alpha = _unpack(desc["rate_param"])
        rate = rate_models.GammaRateModel(desc["rate_cats"] or 4, alpha if alpha is not None else 0.5)
    else:
        rate = rate_models.UniformRateModel()
    alphabet = alphabets.PROTEIN if name in PROTEIN_MODELS else alphabets.DNA
    return model, rate, alphabet


def parse_cli(argv=None):
    parser = argparse.ArgumentParser('phy - calculate likelihood of an alignment given a phylogenetic model')
    parser.add_argument('-t', '--tree', type=str, help='File path to a tree in newick format')
    parser.add_argument('-s', '--alignment', type=str, help='File path to an alignment in fasta (or phylip) format')
    parser.add_argument('-m', '--model', type=str, help='Model specification string', default='JC')
    parser.add_argument('--format', type=str, default='fasta')
    parser.add_argument('--device', type=int, default=0)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_cli(argv)
    for label, path in (("Tree", args.tree), ("Alignment", args.alignment)):
        if path is None:
            sys.stderr.write('ERROR: {} file not specified\n'.format(label))
            return 1
        if not os.path.exists(path):
            sys.stderr.write('ERROR: {} file {} does not exist\n'.format(label, path))
            return 1
    try:
        model, rate, alphabet = build_models(parse_model_string(args.model))
        tree = Tree.get_from_path(args.tree, schema='newick')
    except ValueError as exc:
        sys.stderr.write('ERROR: {}\n'.format(exc))
        return 1
    from .tree_model import TreeModel
    tm = TreeModel(device=args.device)
    tm.set_tree(tree)
    tm.set_alignment(read_alignment(args.alignment, args.format), alphabet)
    tm.set_rate_model(rate)
    tm.set_substitution_model(model)
    tm.initialise()
    print('lnL = {}'.format(tm.compute_likelihood_at_edge(*tm.traversal.root_edge).sum()))
    return 0
