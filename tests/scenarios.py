"""Scenario definitions used by the parity tests, bench.py and the golden-vector generator.

Each scenario is `(dims, setup)`: dims = (sites, demes, groups) and `setup(e)` applies parameter
setters to an ENGINE-level object `e` — either vgsim_b200._engine.BirthDeathModel or the reference's
own BirthDeathModel (oracle/_ref); both expose the same positional setter signatures
(reference src/_BirthDeath.pyx:1431-1702).  The models are the ones the reference's scenario script
defines (testing/check_simulator.py:34-151, table in SURVEY.md §4), its example
(testing/example.py:5-46), the Table-3 model (data/Table 3/Table 3.py:5-22) and the throughput
configurations of SURVEY.md §8(d).
"""


def s1(e):
    e.set_transmission_rate(4.0, None)
    e.set_recovery_rate(1.5, None)
    e.set_sampling_rate(0.3, None)


def s2(e):
    e.set_transmission_rate(4, 3)


def s3(e):
    e.set_susceptibility_type(1, None)


def s4(e):
    e.set_mutation_rate(0.01, None, None)
    e.set_susceptibility_type(1, 0)
    e.set_susceptibility_type(2, 1)
    e.set_susceptibility_type(2, 2)
    e.set_susceptibility_type(2, 3)
    e.set_immunity_transition(0.01, 0, 1)
    e.set_immunity_transition(0.01, 1, 2)
    e.set_immunity_transition(0.02, 2, 1)


def s5(e):
    e.set_population_size(2000000, None)
    e.set_contact_density(1.3, 0)
    e.set_contact_density(0.8, 1)
    e.set_migration_probability(0.01, 0, 1)
    e.set_migration_probability(0.005, 1, 0)


def s6(e):
    e.set_migration_probability(0.01, 0, 1)
    e.set_migration_probability(0.005, 2, 1)


def s7(e):
    s6(e)
    e.set_sampling_multiplier(2.5, 1)
    e.set_sampling_multiplier(2, 2)
    e.set_npi([0.5, 0.30, 0.15], 0)


def s8(e):
    e.set_mutation_rate(0.01, None, None)
    e.set_mutation_probabilities([1, 0, 0, 1], None, None)


def s9(e):
    e.set_transmission_rate(5.0, 12)
    e.set_recovery_rate(1.5, None)
    e.set_sampling_rate(0.3, None)
    e.set_mutation_rate(0.01, None, None)
    e.set_mutation_probabilities([1, 0, 0, 1], None, None)
    e.set_migration_probability(0.01, 0, 1)
    e.set_migration_probability(0.005, 2, 1)
    e.set_sampling_multiplier(2.5, 1)
    e.set_sampling_multiplier(2, 2)
    e.set_npi([0.5, 0.30, 0.15], 1)
    e.set_susceptibility_type(1, 0)
    e.set_susceptibility_type(2, 1)
    e.set_susceptibility_type(2, 2)
    e.set_susceptibility_type(2, 3)
    e.set_immunity_transition(0.000001, 0, 1)
    e.set_immunity_transition(0.000001, 1, 2)
    e.set_immunity_transition(0.000002, 2, 1)


def example(e):
    """testing/example.py:9-46 (engine-level calls)."""
    e.set_transmission_rate(0.25, None)
    e.set_transmission_rate(0.5, "GG")
    e.set_recovery_rate(0.099, None)
    e.set_sampling_rate(0.001, None)
    e.set_mutation_rate(0.00003, None, None)
    e.set_mutation_probabilities([1, 1, 1, 2], None, None)
    e.set_mutation_rate(3 * 0.00003, 'G*', 1)
    e.set_susceptibility_type(1, None)
    e.set_susceptibility_type(2, 'G*')
    e.set_susceptibility(0.1, None, 1)
    e.set_susceptibility(0.5, 'G*', 1)
    e.set_susceptibility(0.0, None, 2)
    e.set_immunity_transition(1 / 90, 1, 0)
    e.set_immunity_transition(1 / 180, 2, 0)
    e.set_population_size(10000000, 0)
    e.set_population_size(5000000, 1)
    e.set_population_size(1000000, 2)
    e.set_migration_probability(10 / 365 / 2, None, None)
    e.set_sampling_multiplier(3, 1)
    e.set_sampling_multiplier(0, 2)
    e.set_npi([0.1, 0.01, 0.002], None)


def example_phase2(e):
    """testing/example.py:51-58: parameter changes between the two simulate() calls."""
    e.set_immunity_transition(0.05, 0, 1)
    e.set_immunity_transition(0.05, 0, 2)
    e.set_contact_density(0.7, 0)
    e.set_contact_density(0.7, 1)
    e.set_migration_probability(2 / 365 / 2, 0, 2)
    e.set_migration_probability(2 / 365 / 2, 1, 2)


def t3(e):
    """3 sites (64 haplotypes) x 10 demes x 3 groups, 1e6 per deme."""
    e.set_transmission_rate(0.25, None)
    e.set_transmission_rate(0.5, 'GGG')
    e.set_recovery_rate(0.099, None)
    e.set_sampling_rate(0.001, None)
    e.set_mutation_rate(3e-4, None, None)
    e.set_mutation_probabilities([1, 1, 1, 2], None, None)
    e.set_susceptibility_type(1, None)
    e.set_susceptibility_type(2, 'G**')
    e.set_susceptibility(0.1, None, 1)
    e.set_susceptibility(0.5, 'G**', 1)
    e.set_susceptibility(0.0, None, 2)
    e.set_immunity_transition(1 / 90, 1, 0)
    e.set_immunity_transition(1 / 180, 2, 0)
    e.set_migration_probability(10 / 365 / 9, None, None)


def table3(K, total_migration=0.005, size=1000000):
    """data/Table 3/Table 3.py:5-22 with the removed API mapped as SURVEY §8(d) config 4 says."""
    def setup(e):
        e.set_transmission_rate(0.25, None)
        e.set_recovery_rate(0.099, None)
        e.set_sampling_rate(0.001, None)
        e.set_mutation_rate(1e-3, None, None)
        e.set_susceptibility_type(1, None)
        e.set_susceptibility(0.1, None, 1)
        e.set_susceptibility(0.0, None, 2)
        e.set_immunity_transition(1 / 90, 1, 0)
        e.set_immunity_transition(1 / 180, 2, 0)
        for p in range(K):
            e.set_population_size(size, p)
        e.set_total_migration_probability(total_migration)
        e.set_npi([0.1, 0.01, 0.002], None)
    return setup


SCENARIOS = {
    "s1": ((0, 1, 1), s1), "s2": ((1, 1, 1), s2), "s3": ((0, 1, 2), s3), "s4": ((1, 1, 3), s4),
    "s5": ((0, 2, 1), s5), "s6": ((0, 3, 1), s6), "s7": ((0, 3, 1), s7), "s8": ((2, 1, 1), s8),
    "s9": ((2, 3, 3), s9),
    "example": ((2, 3, 3), example),
    "t3": ((3, 10, 3), t3),
    "w": ((2, 100, 3), table3(100)),
    "table3_k10": ((2, 10, 3), table3(10)),
}


def _t3small(e):
    """T3 recipe on a small shape (2 sites, 4 demes, 3 groups) for quick parity runs."""
    e.set_transmission_rate(0.25, None)
    e.set_transmission_rate(0.5, 'GG')
    e.set_recovery_rate(0.099, None)
    e.set_sampling_rate(0.001, None)
    e.set_mutation_rate(3e-4, None, None)
    e.set_mutation_probabilities([1, 1, 1, 2], None, None)
    e.set_susceptibility_type(1, None)
    e.set_susceptibility_type(2, 'G*')
    e.set_susceptibility(0.1, None, 1)
    e.set_susceptibility(0.5, 'G*', 1)
    e.set_susceptibility(0.0, None, 2)
    e.set_immunity_transition(1 / 90, 1, 0)
    e.set_immunity_transition(1 / 180, 2, 0)
    e.set_migration_probability(10 / 365 / 3, None, None)


SCENARIOS["t3small"] = ((2, 4, 3), _t3small)


def _s8hi(e):
    """Scenario 8 with a mutation rate high enough that sampled lineages carry several mutations per branch (the
    writers' multiple-mutations-per-node path) and dense sampling."""
    e.set_mutation_rate(0.6, None, None)
    e.set_mutation_probabilities([1, 0, 0, 1], None, None)
    e.set_sampling_rate(0.2, None)


SCENARIOS["s8hi"] = ((2, 1, 1), _s8hi)
