"""BASELINE.json configs[0]: the reference's example script (testing/example.py:5-83) run end to end on the device --
direct Gillespie to t = 110, parameter changes, the tau call (which takes 0 leaps: quirk Q5, the default
sample_size = iterations = 1000 is already exceeded), genealogy, and the three exports -- once through
`vgsim_b200.Simulator` and once through the REFERENCE'S OWN `Simulator` class (src/_interface.py, byte-compiled into
oracle/_ref/VGsim/_interface.pyc.bin by oracle/build_ref.py) with the one-line engine substitution INTEGRATION.md describes.
Both runs use the same seed and the same engine, so their output files must be identical: that is the drop-in claim."""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "oracle", "_ref", "VGsim")


def run_example(VGsim, outdir):
    """testing/example.py:5-83, statement for statement (plot calls are commented out upstream too)."""
    number_of_sites = 2
    populations_number = 3
    number_of_susceptible_groups = 3
    simulator = VGsim.Simulator(number_of_sites, populations_number, number_of_susceptible_groups, seed=1234)

    # Set epidemiological parameters
    simulator.set_transmission_rate(0.25)
    simulator.set_transmission_rate(0.5, haplotype="GG")
    simulator.set_recovery_rate(0.099)
    simulator.set_sampling_rate(0.001)

    # Set mutation rates
    mutation_rate = 0.00003
    substitution_weights = [1, 1, 1, 2]  # ATCG
    simulator.set_mutation_rate(mutation_rate)
    simulator.set_mutation_probabilities(substitution_weights)
    simulator.set_mutation_rate(3 * mutation_rate, haplotype='G*', mutation=1)

    # Set host immunity types triggered by infection
    simulator.set_susceptibility_type(1)
    simulator.set_susceptibility_type(2, haplotype='G*')
    simulator.set_susceptibility(0.1, susceptibility_type=1)
    simulator.set_susceptibility(0.5, susceptibility_type=1, haplotype='G*')
    simulator.set_susceptibility(0.0, susceptibility_type=2)

    # Set loss of immunity
    simulator.set_immunity_transition(1 / 90, source=1, target=0)
    simulator.set_immunity_transition(1 / 180, source=2, target=0)

    # Set host population structure with three populations and migration
    simulator.set_population_size(10000000, population=0)
    simulator.set_population_size(5000000, population=1)
    simulator.set_population_size(1000000, population=2)
    simulator.set_migration_probability(10 / 365 / 2)

    # Set specific sampling efforts in different populations
    simulator.set_sampling_multiplier(3, population=1)
    simulator.set_sampling_multiplier(0, population=2)

    # Set non-pharmasutical interventions (same for each population)
    simulator.set_npi([0.1, 0.01, 0.002])

    # Run simulation with the exact algorithm
    simulator.simulate(10000000, epidemic_time=110)

    # Change some of parameters
    simulator.set_immunity_transition(0.05, source=0, target=1)
    simulator.set_immunity_transition(0.05, source=0, target=2)
    simulator.set_contact_density(0.7, population=0)
    simulator.set_contact_density(0.7, population=1)
    simulator.set_migration_probability(2 / 365 / 2, source=0, target=2)
    simulator.set_migration_probability(2 / 365 / 2, source=1, target=2)

    # Run simulation with the approximate tau-leaping algorithm
    simulator.simulate(1000, epidemic_time=210, method='tau')

    # Simulate genealogy of sampled individuals
    simulator.genealogy()

    # Create directory to save results
    cwd = os.getcwd()
    os.makedirs(outdir, exist_ok=True)
    os.chdir(outdir)
    try:
        # Save the genealogy with mutations and migrations
        simulator.export_newick()
        simulator.export_mutations('mutations')
        simulator.export_migrations('migrations')
    finally:
        os.chdir(cwd)
    return simulator


def read_all(d):
    return {f: open(os.path.join(d, f)).read() for f in ("tree.nwk", "sample_population.tsv", "mutations.tsv", "migrations.tsv")}


def test_example_script_runs_on_the_device(tmp_path, capsys):
    import vgsim_b200 as VGsim
    sim = run_example(VGsim, str(tmp_path / "example_output"))
    c = sim.simulation.counters()
    # the reference's own run of this script (probe build, PCG64 stream): 741,080 events, 4,513 samples
    assert 110.0 <= c["time"][0] < 111.0
    assert 1e5 < c["events"][0] < 5e6 and 500 < c["sCounter"][0] < 50000
    assert c["leaps"][0] == 0, "the tau call must stop at once: sample_size defaults to iterations = 1000 (quirk Q5)"
    files = read_all(str(tmp_path / "example_output"))
    n = int(c["sCounter"][0])
    nwk = files["tree.nwk"]
    assert nwk.endswith(";") and nwk.count("(") == n - 1 and nwk.count(",") == n - 1
    pops = [l.split("\t") for l in files["sample_population.tsv"].splitlines()]
    assert len(pops) == 2 * n - 1 and {int(p[1]) for p in pops} <= {0, 1, 2}
    assert files["migrations.tsv"].startswith("Node\tTime\tOld_population\tNew_population\n")
    tree, times = sim.get_tree()
    assert (tree == -1).sum() == 1 and len(tree) == 2 * n - 1
    # sampling multiplier 0 in deme 2: no leaf there
    leaf_pop = {int(p[0]): int(p[1]) for p in pops}
    kids = np.bincount(tree[tree >= 0], minlength=len(tree))
    assert all(leaf_pop[i] != 2 for i in np.flatnonzero(kids == 0))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_PKG, "_interface.pyc.bin")), reason="oracle/_ref (reference build) not present")
def test_reference_simulator_class_over_the_new_engine(tmp_path):
    """The reference's Simulator (unmodified bytecode of src/_interface.py + src/IO.py) with `_BirthDeath` replaced by the
    three-line stub of INTEGRATION.md section 1 gives the same files as vgsim_b200.Simulator."""
    pkg = tmp_path / "VGsim_dropin"
    pkg.mkdir()
    shutil.copy(os.path.join(REF_PKG, "_interface.pyc.bin"), pkg / "_interface.pyc")
    shutil.copy(os.path.join(REF_PKG, "IO.pyc.bin"), pkg / "IO.pyc")
    (pkg / "_BirthDeath.py").write_text(
        "# INTEGRATION.md section 1: the engine the reference's wrapper binds\n"
        "from vgsim_b200._engine import BirthDeathModel  # ctypes -> libvgsim_b200.so (sm_100a CUDA)\n")
    (pkg / "__init__.py").write_text("from ._interface import Simulator\n")
    stubs = os.path.join(ROOT, "oracle", "_ref")       # import stub for matplotlib.pyplot (plotting is out of scope)
    sys.path[:0] = [str(tmp_path)]
    added_stub = False
    try:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            sys.path.append(stubs)
            added_stub = True
        dropin = importlib.import_module("VGsim_dropin")
        assert type(dropin.Simulator).__module__ == "builtins" and dropin.Simulator.__module__ == "VGsim_dropin._interface"
        a = run_example(dropin, str(tmp_path / "out_reference_wrapper"))
        from vgsim_b200._engine import BirthDeathModel
        assert isinstance(a.simulation, BirthDeathModel)
    finally:
        sys.path.remove(str(tmp_path))
        if added_stub:
            sys.path.remove(stubs)
    import vgsim_b200
    run_example(vgsim_b200, str(tmp_path / "out_own_wrapper"))
    fa, fb = read_all(str(tmp_path / "out_reference_wrapper")), read_all(str(tmp_path / "out_own_wrapper"))
    for k in fa:
        assert fa[k] == fb[k], k
    assert len(fa["tree.nwk"]) > 1000
