"""Text parameter readers and the command-line front end (reference src/IO.py:4-142,
VGsim_cmd.py).  Host logic only on CPU; one end-to-end CLI run on the GPU."""
import importlib.util
import os

import numpy as np
import pytest

from vgsim_b200 import io as vio

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
REF_IO = "/root/reference/src/IO.py"
REF_EX = "/root/reference/testing/cmd_example/example"


def test_read_rates():
    b, d, s, m = vio.read_rates(os.path.join(DATA, "model.rt"))
    assert b == [0.30, 0.35, 0.25, 0.40] and d == [0.10, 0.10, 0.12, 0.09] and s == [0.002, 0.002, 0.001, 0.003]
    # a 0 is inserted at the haplotype's own allele (A, T, C, G = 0..3): [rate, wA, wT, wC, wG]
    assert m[0][0] == [2e-4, 0, 1.0, 2.0, 3.0]
    assert m[1][0] == [3e-4, 1.0 / 3.0, 0, 1.0 / 3.0, 1.0 / 3.0]
    assert m[2][0] == [1e-4, 1.0, 1.0, 0, 1.0]
    assert m[3][0] == [5e-4, 2.0, 1.0, 1.0, 0]


def test_read_rates_sampling_probability():
    b, d, s, m = vio.read_rates(os.path.join(DATA, "model_sp.rt"))
    assert b == [0.5] and d == [0.2 * (1 - 0.25)] and s == [0.2 * 0.25] and m == [[]]


def test_read_susceptibility_skips_comments():
    sus, typ = vio.read_susceptibility(os.path.join(DATA, "model.su"))
    assert typ == [1, 1, 0, 1]
    assert sus == [["1.0", "0.2"], ["1.0", "0.1"], ["1.0", "0.3"], ["1.0", "0.0"]]


def test_read_populations_optional_columns():
    sizes, cd, after, start, end, mult = vio.read_populations(os.path.join(DATA, "model.pp"))
    assert sizes == [200000, 100000, 50000, 40000] and cd == [1.0, 0.9, 1.1, 1.2]
    assert after == [0.2, 0.3, 0, 0.4] and start == [0.02, 0.03, 1.0, 0.05] and end == [0.004, 0.005, 1.0, 0.01]
    assert mult == [1.5, 2.0, 0.5, 1]


def test_read_matrix():
    m = vio.read_matrix(os.path.join(DATA, "model.mg"))
    assert len(m) == 4 and m[1] == [0.01, 0.0, 0.01, 0.01]
    assert vio.read_matrix(os.path.join(DATA, "model.st")) == [[0.0, 0.0], [0.02, 0.0]]


def test_malformed_rates_raise(tmp_path):
    f = tmp_path / "bad.rt"
    f.write_text("#v\nH B D S M0\nA 1 1 1 1e-3,1,2\n")
    with pytest.raises(ValueError):
        vio.read_rates(str(f))
    f.write_text("#v\nH B D S\nA 1 1 1\nT 1 1 1\n")
    with pytest.raises(ValueError):
        vio.read_rates(str(f))


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="reference sources not mounted")
def test_readers_equal_reference_on_its_example_files():
    spec = importlib.util.spec_from_file_location("ref_io", REF_IO)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    assert vio.read_rates(REF_EX + ".rt") == ref.read_rates(REF_EX + ".rt")
    assert vio.read_susceptibility(REF_EX + ".su") == ref.read_susceptibility(REF_EX + ".su")
    assert vio.read_populations(REF_EX + ".pp") == ref.read_populations(REF_EX + ".pp")
    assert vio.read_matrix(REF_EX + ".mg") == ref.read_matrix(REF_EX + ".mg")
    assert vio.read_matrix(REF_EX + ".st") == ref.read_matrix(REF_EX + ".st")
    for name in ("model.rt", "model_sp.rt", "model.mg", "model.st"):
        fn = os.path.join(DATA, name)
        rd = "read_rates" if name.endswith(".rt") else "read_matrix"
        assert getattr(vio, rd)(fn) == getattr(ref, rd)(fn)


def test_cli_configures_the_engine_from_files():
    """No GPU needed: parameter setters are host-side; the handle is only created by simulate()."""
    from vgsim_b200 import cli
    args = cli.build_parser().parse_args(["-rt", os.path.join(DATA, "model.rt"), "-pm", os.path.join(DATA, "model.pp"),
                                          os.path.join(DATA, "model.mg"), "-su", os.path.join(DATA, "model.su"),
                                          "-st", os.path.join(DATA, "model.st"), "-seed", "5"])
    sim, seed = cli.configure(args)
    e = sim.simulation
    assert seed == 5 and (e.sites, e.popNum, e.susNum) == (1, 4, 2)
    np.testing.assert_allclose(e.transmission_rate, [0.30, 0.35, 0.25, 0.40])
    np.testing.assert_allclose(e.recovery_rate, [0.10, 0.10, 0.12, 0.09])
    np.testing.assert_allclose(e.mutation_rate[:, 0], [2e-4, 3e-4, 1e-4, 5e-4])
    np.testing.assert_allclose(e.susceptibility[:, 1], [0.2, 0.1, 0.3, 0.0])
    np.testing.assert_array_equal(e.susceptibility_type, [1, 1, 0, 1])
    np.testing.assert_array_equal(e.population_size, [200000, 100000, 50000, 40000])
    np.testing.assert_allclose(e.sampling_multiplier, [1.5, 2.0, 0.5, 1.0])
    np.testing.assert_allclose(e.immunity_transition, [[0.0, 0.0], [0.02, 0.0]])
    mig = np.asarray(e.migration_probability)
    assert mig[0, 1] == 0.01 and mig[2, 3] == 0.005 and abs(mig[1].sum() - 1.0) < 1e-12


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    from vgsim_b200 import cli
    out = str(tmp_path / "run")
    rc = cli.main(["-rt", os.path.join(DATA, "model.rt"), "-pm", os.path.join(DATA, "model.pp"), os.path.join(DATA, "model.mg"),
                   "-su", os.path.join(DATA, "model.su"), "-st", os.path.join(DATA, "model.st"), "-it", "20000", "-s", "300",
                   "-seed", "11", "-nwk", out, "-tsv", out, "--writeMigrations", out + "_mig", "--output_chain_events", out])
    assert rc == 0
    nwk = open(out + "_tree.nwk").read().strip()
    assert nwk.endswith(";") and nwk.count("(") == nwk.count(")") >= 50
    assert os.path.exists(out + ".tsv") and os.path.exists(out + "_sample_population.tsv")
    assert open(out + "_mig.tsv").read().startswith("Node\tTime\tOld_population\tNew_population")
    chain = np.load(out + ".npy")
    # a binary tree over every sampled individual: one '(' per internal node
    assert chain.shape[0] == 6 and (chain[1] == 2).sum() == nwk.count("(") + 1


def test_export_settings_round_trip(tmp_path, capsys):
    """export_settings -> the readers -> cli.configure gives back the same model (host side only)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from scenarios import SCENARIOS
    from vgsim_b200 import cli
    from vgsim_b200._engine import BirthDeathModel as Eng
    for name in ("example", "s9", "s1"):
        (U, K, S), setup = SCENARIOS[name]
        e = Eng(U, K, S, 3, False, False, int(1e6), 0.0)
        setup(e)
        d = str(tmp_path / ("model_" + name))
        e.export_settings(d)
        base = os.path.join(d, "model_" + name)
        assert "Command line command: " + base + ".rt -pm " in capsys.readouterr().out
        args = cli.build_parser().parse_args(["-rt", base + ".rt", "-pm", base + ".pp", base + ".mg", "-su", base + ".su",
                                              "-st", base + ".st", "-seed", "3"])
        sim, _ = cli.configure(args)
        a, b = e.param_arrays(), sim.simulation.param_arrays()
        for k in a:
            if k in ("cd", "cdBefore"):
                continue
            np.testing.assert_allclose(np.asarray(b[k], float), np.asarray(a[k], float), rtol=1e-15, atol=0, err_msg=name + ":" + k)
        np.testing.assert_allclose(np.asarray(b["cd"], float), np.asarray(a["cd"], float), rtol=1e-15)
