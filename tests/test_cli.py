"""Model-string grammar and the command-line front end (reference: bin/phy.py:41-146)."""
import numpy as np
import pytest

from phylo_utils_b200 import cli, rate_models, substitution_models as sm
from helpers import load


def test_model_string_grammar():
    d = cli.parse_model_string("GTR{6.0,5.0,4.0,3.0,2.0,1.0}+F{0.1,0.2,0.3,0.4}+G4{0.5}")
    assert d["subs_model"] == "GTR" and d["model_params"] == [6., 5., 4., 3., 2., 1.]
    assert d["freq_params"] == [0.1, 0.2, 0.3, 0.4] and d["rate_model"] == "G" and d["rate_cats"] == 4 and d["rate_param"] == [0.5]
    d = cli.parse_model_string("HKY{2.5}+G8")
    assert d["subs_model"] == "HKY" and d["model_params"] == [2.5] and d["rate_cats"] == 8 and d["rate_param"] is None
    assert cli.parse_model_string("JC")["rate_model"] is None
    assert cli.parse_model_string("LG+G4{0.8}+F{%s}" % ",".join(["0.05"] * 20))["freq_params"] == [0.05] * 20
    for bad in ("GTR{6,5}", "GTR+X4", "GTR+G4+F+G4", "{1.0}"):
        with pytest.raises(ValueError):
            cli.parse_model_string(bad)


def test_models_are_built_like_the_reference_cli_does():
    m, r, alpha = cli.build_models(cli.parse_model_string("GTR{6.0,5.0,4.0,3.0,2.0,1.0}+F{0.1,0.2,0.3,0.4}+G4{0.5}"))
    assert isinstance(m, sm.GTR) and isinstance(r, rate_models.GammaRateModel) and r.ncat == 4 and alpha == 0
    assert np.allclose(m.q(), sm.GTR([6., 5., 4., 3., 2., 1.], [0.1, 0.2, 0.3, 0.4]).q())
    m, r, alpha = cli.build_models(cli.parse_model_string("WAG"))
    assert isinstance(m, sm.WAG) and isinstance(r, rate_models.UniformRateModel) and alpha == 1
    m, r, _ = cli.build_models(cli.parse_model_string("K80{2.0}+G4"))
    assert isinstance(m, sm.K80) and r.alpha == 0.5
    with pytest.raises(ValueError):
        cli.build_models(cli.parse_model_string("XYZ"))


def test_missing_files_are_reported(tmp_path, capsys):
    assert cli.main(["-t", str(tmp_path / "none.nwk"), "-s", str(tmp_path / "none.fa")]) == 1
    assert "does not exist" in capsys.readouterr().err


@pytest.mark.gpu
def test_cli_end_to_end_matches_reference(tmp_path, capsys):
    g = load("cfg1_gtr_g4")
    (tmp_path / "t.nwk").write_text(str(g["newick"]))
    with open(tmp_path / "a.fa", "w") as fh:
        for name, row in zip(g["names"], g["seqs"]):
            fh.write(">{}\n{}\n".format(name, bytes(row).decode()))
    rc = cli.main(["-t", str(tmp_path / "t.nwk"), "-s", str(tmp_path / "a.fa"), "-m",
                   "GTR{6.0,5.0,4.0,3.0,2.0,1.0}+F{0.1,0.2,0.3,0.4}+G4{0.5}"])
    assert rc == 0
    out = capsys.readouterr().out.strip()
    assert out.startswith("lnL = ")
    assert abs(float(out.split("=")[1]) - float(g["total_lnl"])) <= 1e-10 * abs(float(g["total_lnl"]))
