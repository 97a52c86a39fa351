"""Command-line front end with the reference tool's flags (VGsim_cmd.py:11-150): reads the text parameter files,
configures a Simulator, runs simulate + genealogy and writes the requested outputs.

    python -m vgsim_b200.cli -rt model.rt -pm model.pp model.mg -su model.su -st model.st \\
        -it 100000 -s 1000 -seed 7 -nwk tree -tsv mutations --writeMigrations migrations

Additions: --method direct|tau, --replicates N (outputs are written for replicate 0), --device.
"""
import argparse
import math
import sys
from random import randrange

from . import io as _io
from ._interface import Simulator


def build_parser():
    ap = argparse.ArgumentParser(prog="vgsim_b200", description="VGsim-compatible epidemic + genealogy simulation on a B200")
    ap.add_argument("--iterations", "-it", type=int, default=1000, help="number of iterations (default is 1000)")
    ap.add_argument("--sampleSize", "-s", type=int, default=None, help="number of samples (default: iterations)")
    ap.add_argument("--time", "-t", type=float, default=None, help="epidemic time at which the simulation stops")
    ap.add_argument("--seed", "-seed", type=float, default=None, help="random seed")
    ap.add_argument("--rates", "-rt", default=None, help="file with rates for each haplotype")
    ap.add_argument("--populationModel", "-pm", nargs=2, default=None, metavar=("POPULATIONS", "MIGRATION"),
                    help="population file and migration-probability matrix file")
    ap.add_argument("--susceptibility", "-su", default=None, help="susceptibility file")
    ap.add_argument("--suscepTransition", "-st", default=None, help="susceptibility transition matrix file")
    ap.add_argument("--sampling_probability", action="store_true", help="the S column of the rates file is a probability")
    ap.add_argument("--createNewick", "-nwk", default=False, help="write the tree to <name>.nwk")
    ap.add_argument("--writeMutations", "-tsv", default=False, help="write the mutations to <name>.tsv")
    ap.add_argument("--writeMigrations", default=False, help="write the migrations to <name>.tsv")
    ap.add_argument("--output_chain_events", default=False, help="save the event chain to <name>.npy")
    ap.add_argument("--method", default="direct", choices=["direct", "tau"])
    ap.add_argument("--replicates", type=int, default=1)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("-citation", "-c", action="store_true", help="information for citation")
    return ap


def configure(args):
    """Simulator configured from the files named in `args` (the reference's defaults where a file is absent,
    VGsim_cmd.py:79-108)."""
    if args.rates is None:
        b, d, s, m = [2], [1], [0.1], [[]]
    else:
        b, d, s, m = _io.read_rates(args.rates)
    if args.populationModel is None:
        sizes, cd, after, start, end, mult = [1000000], [1], [1], [1], [1], [1]
        mig = [[0.0]]
    else:
        sizes, cd, after, start, end, mult = _io.read_populations(args.populationModel[0])
        mig = _io.read_matrix(args.populationModel[1])
    if args.susceptibility is None:
        sus, typ = [[1.0] for _ in b], [0 for _ in b]
    else:
        sus, typ = _io.read_susceptibility(args.susceptibility)
    trans = [[0.0]] if args.suscepTransition is None else _io.read_matrix(args.suscepTransition)
    seed = randrange(sys.maxsize) if args.seed is None else args.seed
    sim = Simulator(number_of_sites=int(round(math.log(len(b), 4))), populations_number=len(sizes),
                    number_of_susceptible_groups=len(sus[0]), seed=int(seed), sampling_probability=args.sampling_probability,
                    replicates=args.replicates, device=args.device)
    for h in range(len(b)):
        sim.set_transmission_rate(b[h], h)
        sim.set_recovery_rate(d[h], h)
        sim.set_sampling_rate(s[h], h)
        for u in range(len(m[0])):
            sim.set_mutation_rate(m[h][u][0], h, u)
            sim.set_mutation_probabilities(m[h][u][1:5], h, u)
    for p in range(len(sizes)):
        sim.set_population_size(sizes[p], p)
        sim.set_contact_density(cd[p], p)
        sim.set_npi([after[p], start[p], end[p]], p)
        sim.set_sampling_multiplier(mult[p], p)
        for q in range(len(sizes)):
            if p != q:
                sim.set_migration_probability(probability=mig[p][q], source=p, target=q)
    for h in range(len(sus)):
        for g in range(len(sus[h])):
            sim.set_susceptibility(float(sus[h][g]), h, g)
    for h in range(len(typ)):
        sim.set_susceptibility_type(typ[h], h)
    for i in range(len(trans)):
        for j in range(len(trans[i])):
            if i != j:
                sim.set_immunity_transition(trans[i][j], i, j)
    return sim, int(seed)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.citation:
        print("VGsim: scalable viral genealogy simulator for global pandemic")
        print("Vladimir Shchur, Vadim Spirin, Victor Pokrovskii, Evgeni Burovski, Nicola De Maio, Russell Corbett-Detig")
        print("medRxiv 2021.04.21.21255891; doi: https://doi.org/10.1101/2021.04.21.21255891")
        return 0
    sim, seed = configure(args)
    sample = args.iterations if args.sampleSize is None else args.sampleSize
    sim.simulate(args.iterations, sample, -1 if args.time is None else args.time, method=args.method)
    sim.genealogy(seed)
    if args.createNewick:
        sim.export_newick(args.createNewick)
    if args.writeMutations:
        sim.export_mutations(args.writeMutations)
    if args.writeMigrations:
        sim.export_migrations(args.writeMigrations)
    if args.output_chain_events:
        sim.export_chain_events(args.output_chain_events)
    return 0


if __name__ == "__main__":
    sys.exit(main())
