"""Boundary conformance of the host side: `Simulator` keeps the reference's setter semantics, stored
arrays and exact exception types/messages (reference src/_interface.py:9-883, engine setters
src/_BirthDeath.pyx:1380-1702).  Runs without a GPU: parameter handling is plain host Python.

Two layers:
  * test_reference_suite_passes_unchanged: where the reference checkout is present (this container, not
    the GPU box) its own tests/test_interface.py (278 cases) is executed UNMODIFIED against this package
    through a `VGsim` alias module.
  * the cases below: our own vectors for the same API, always run.
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
from numpy.testing import assert_allclose

from vgsim_b200 import Simulator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests/test_interface.py"


def sim(**kw):
    kw.setdefault("verbose", False)
    kw.setdefault("seed", 1)
    return Simulator(**kw)


@pytest.mark.skipif(not os.path.exists(REF_TESTS), reason="reference checkout not present")
def test_reference_suite_passes_unchanged(tmp_path):
    alias = tmp_path / "VGsim"
    alias.mkdir()
    (alias / "__init__.py").write_text("from vgsim_b200 import Simulator\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=str(tmp_path))
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and "278 passed" in tail, r.stdout[-2000:]


# ---------------------------------------------------------------------------------------- constructor
def test_constructor_shapes_and_defaults():
    m = sim(number_of_sites=2, populations_number=3, number_of_susceptible_groups=2)
    assert (m.number_of_sites, m.haplotypes_number, m.populations_number, m.number_of_susceptible_groups) == (2, 16, 3, 2)
    assert_allclose(m.transmission_rate, np.full(16, 2.0))
    assert_allclose(m.recovery_rate, np.full(16, 1.0))
    assert_allclose(m.sampling_rate, np.full(16, 0.01))
    assert_allclose(m.mutation_rate, np.full((16, 2), 0.01))
    assert_allclose(m.mutation_probabilities, np.ones((16, 2, 3)))
    assert_allclose(m.susceptibility, np.tile([1.0, 0.0], (16, 1)))
    assert_allclose(m.population_size, np.full(3, 1000000))
    assert_allclose(m.susceptible, np.tile([1000000, 0], (3, 1)))
    assert_allclose(m.infectious, np.zeros((3, 16)))
    assert_allclose(m.migration_probability, np.zeros((3, 3)))
    assert_allclose(m.contact_density, np.ones(3))
    after, start, end = m.npi
    assert_allclose(after, np.zeros(3)); assert_allclose(start, np.ones(3)); assert_allclose(end, np.ones(3))


@pytest.mark.parametrize("kw,err,text", [
    (dict(number_of_sites=1.5), TypeError, "Incorrect type of number of sites. Type should be int."),
    (dict(number_of_sites=-1), ValueError, "Incorrect value of number of sites. Value should be more or equal 0."),
    (dict(populations_number=0), ValueError, "Incorrect value of populations number. Value should be more 0."),
    (dict(number_of_susceptible_groups="2"), TypeError, "Incorrect type of number of susceptible groups. Type should be int."),
    (dict(seed=-5), ValueError, "Incorrect value of seed. Value should be more or equal 0."),
    (dict(sampling_probability=3), ValueError, "Incorrect value of sampling probability. Value of sampling probability should be True or False."),
    (dict(memory_optimization="no"), ValueError, "Incorrect value of memory optimization. Value of memory optimization should be True or False."),
    (dict(number_of_sites=3, genome_length=2), ValueError, "Incorrect value of number of sites or genome length. Genome length should be more or equal number of sites."),
    (dict(recombination_probability=1.5), ValueError, "Incorrect value of recombination probability. Value should be more or equal 0 and equal or less 1."),
])
def test_constructor_errors(kw, err, text):
    with pytest.raises(err, match=text):
        sim(**kw)


# ---------------------------------------------------------------------------------------- haplotype addressing
def test_wildcards_and_lists_address_the_same_cells():
    m = sim(number_of_sites=2)
    m.set_transmission_rate(3.5, "T*")            # T* = haplotypes 4..7 (A,T,C,G = 0..3, site 0 most significant)
    want = np.full(16, 2.0); want[4:8] = 3.5
    assert_allclose(m.transmission_rate, want)
    m.set_transmission_rate(0.5, [0, 15, "*G"])   # *G = 3, 7, 11, 15
    want[[0, 15, 3, 7, 11]] = 0.5
    assert_allclose(m.transmission_rate, want)
    m.set_transmission_rate(9.0)                  # None = every haplotype
    assert_allclose(m.transmission_rate, np.full(16, 9.0))


@pytest.mark.parametrize("hap,err,text", [
    (16, IndexError, "There are no such haplotype!"),
    (-1, IndexError, "There are no such haplotype!"),
    ("AAA", ValueError, "Incorrect haplotype. Haplotype should contain only"),
    ("AX", ValueError, "Incorrect haplotype. Haplotype should contain only"),
    (1.0, TypeError, "Incorrect type of haplotype. Type should be int or str or None."),
])
def test_bad_haplotype(hap, err, text):
    m = sim(number_of_sites=2)
    with pytest.raises(err, match=text):
        m.set_recovery_rate(1.0, hap)


# ---------------------------------------------------------------------------------------- rates
def test_rates_and_sampling_probability_mode():
    m = sim(number_of_sites=1)
    m.set_recovery_rate(0.4, "C")
    m.set_sampling_rate(0.1, 2)
    assert_allclose(m.recovery_rate, [1, 1, 0.4, 1]); assert_allclose(m.sampling_rate, [0.01, 0.01, 0.1, 0.01])
    p = sim(number_of_sites=1, sampling_probability=True)
    p.set_recovery_rate(2.0, None)
    p.set_sampling_rate(0.25, 1)                  # splits d+s = 2.01 into recovery 75 % / sampling 25 %
    assert_allclose(p.recovery_rate, [2, 0.75 * 2.01, 2, 2]); assert_allclose(p.sampling_rate, [0.01, 0.25 * 2.01, 0.01, 0.01])
    with pytest.raises(ValueError, match="Incorrect value of sampling probability. Value should be more or equal 0 and equal or less 1."):
        p.set_sampling_rate(1.2, None)
    with pytest.raises(ValueError, match="Incorrect value of transmission rate. Value should be more or equal 0."):
        m.set_transmission_rate(-1, None)
    with pytest.raises(TypeError, match="Incorrect type of recovery rate. Type should be int or float."):
        m.set_recovery_rate("1", None)


def test_mutation_rate_and_probabilities_delete_own_allele():
    m = sim(number_of_sites=2)
    m.set_mutation_rate(0.3, "A*", 1)
    want = np.full((16, 2), 0.01); want[0:4, 1] = 0.3
    assert_allclose(m.mutation_rate, want)
    m.set_mutation_probabilities([5, 6, 7, 8], "TC", None)   # haplotype 6: site0 = T(1), site1 = C(2)
    w = np.ones((16, 2, 3)); w[6, 0] = [5, 7, 8]; w[6, 1] = [5, 6, 8]
    assert_allclose(m.mutation_probabilities, w)
    with pytest.raises(ValueError, match="Incorrect probabilities list. The sum of three elements without mutation allele should be more 0."):
        m.set_mutation_probabilities([4, 0, 0, 0], 0, 0)
    with pytest.raises(ValueError, match="Incorrect length of probabilities list. Length should be equal 4."):
        m.set_mutation_probabilities([1, 1, 1], None, None)
    with pytest.raises(TypeError, match="Incorrect type of probabilities list. Type should be list."):
        m.set_mutation_probabilities((1, 1, 1, 1), None, None)
    with pytest.raises(IndexError, match="There are no such mutation site!"):
        m.set_mutation_rate(0.1, None, 2)


def test_mutation_position_and_genome_length():
    m = sim(number_of_sites=3, genome_length=100)
    assert list(m.mutation_position) == [0, 50, 100]
    m.set_mutation_position(1, 10)
    assert list(m.mutation_position) == [0, 10, 100]
    with pytest.raises(IndexError, match="Incorrect value of position. Two mutations can't have the same position."):
        m.set_mutation_position(2, 10)
    m.set_genome_length(1000)
    assert m.genome_length == 1000 and list(m.mutation_position) == [0, 500, 1000]


# ---------------------------------------------------------------------------------------- immunity
def test_susceptibility_model():
    m = sim(number_of_sites=1, number_of_susceptible_groups=3)
    m.set_susceptibility_type(2, "G")
    assert list(m.susceptibility_type) == [0, 0, 0, 2]
    m.set_susceptibility(0.4, [0, 1], 1)
    want = np.tile([1.0, 0.0, 0.0], (4, 1)); want[0:2, 1] = 0.4
    assert_allclose(m.susceptibility, want)
    m.set_immunity_transition(0.02, None, 0)      # diagonal entries are never written
    assert_allclose(m.immunity_transition, [[0, 0, 0], [0.02, 0, 0], [0.02, 0, 0]])
    with pytest.raises(IndexError, match="There are no such susceptibility type!"):
        m.set_susceptibility_type(3, None)
    with pytest.raises(TypeError, match="Incorrect type of susceptibility type. Type should be int."):
        m.set_susceptibility_type(None, None)
    with pytest.raises(ValueError, match="Incorrect value of immunity transition rate. Value should be more or equal 0."):
        m.set_immunity_transition(-0.1, 0, 1)


# ---------------------------------------------------------------------------------------- demes
def test_population_parameters():
    m = sim(populations_number=3)
    m.set_population_size(250000, 1)
    assert list(m.population_size) == [1000000, 250000, 1000000] and list(m.susceptible[:, 0]) == [1000000, 250000, 1000000]
    m.set_contact_density(0.6, [0, 2])
    assert_allclose(m.contact_density, [0.6, 1.0, 0.6])
    m.set_npi([0.2, 0.05, 0.01], 2)
    after, start, end = m.npi
    assert_allclose(after, [0, 0, 0.2]); assert_allclose(start, [1, 1, 0.05]); assert_allclose(end, [1, 1, 0.01])
    m.set_sampling_multiplier(4.0, None)
    assert_allclose(m.sampling_multiplier, [4, 4, 4])
    with pytest.raises(ValueError, match="Incorrect length of npi parameters. Length should be equal 3."):
        m.set_npi([0.1, 0.2], None)
    with pytest.raises(ValueError, match="Incorrect value of second npi parameter. Value should be more or equal 0 and equal or less 1."):
        m.set_npi([0.1, 1.2, 0.5], None)
    with pytest.raises(IndexError, match="There are no such population!"):
        m.set_contact_density(1.0, 3)
    with pytest.raises(ValueError, match="Incorrect value of population size. Value should be more 0."):
        m.set_population_size(0, None)


def test_migration_matrix_diagonal_and_errors():
    m = sim(populations_number=3)
    m.set_migration_probability(0.1, 0, None)
    assert_allclose(m.migration_probability, [[0.8, 0.1, 0.1], [0, 1, 0], [0, 0, 1]])
    m.set_migration_probability(0.25, [1, 2], 0)
    assert_allclose(m.migration_probability, [[0.8, 0.1, 0.1], [0.25, 0.75, 0], [0.25, 0, 0.75]])
    m.set_total_migration_probability(0.3)
    assert_allclose(m.migration_probability, np.full((3, 3), 0.15) + np.eye(3) * 0.55)
    with pytest.raises(ValueError, match="Incorrect the sum of migration probabilities. The sum of migration probabilities from each population should be equal or less 1."):
        m.set_migration_probability(0.6, 0, None)
    m2 = sim(populations_number=2)
    with pytest.raises(ValueError, match="Incorrect value of migration probability. Value of migration probability from source population to target population should be more 0."):
        m2.set_migration_probability(1.0, 0, 1)
    with pytest.raises(ValueError, match="Incorrect value of migration probability. Value should be more or equal 0 and equal or less 1."):
        m2.set_migration_probability(1.5, 0, 1)


def test_output_spellings_exist():
    m = sim()
    for name in ("simulate", "genealogy", "export_newick", "export_mutations", "export_migrations",
                 "output_newick", "output_mutations", "output_migrations", "set_susceptibility_type",
                 "set_total_migration_probability"):
        assert callable(getattr(m, name))


def test_hot_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vgsim_b200._capi import VgsimError
    with pytest.raises(VgsimError, match="no CUDA device"):
        sim().simulate(10)
