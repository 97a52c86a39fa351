#!/usr/bin/env python
"""Golden vectors for the Newick / TSV writers: the UNMODIFIED reference engine (oracle/_ref, built by
oracle/build_ref.py) simulates scenario 9, builds the genealogy, and the reference's own writers (src/IO.py:144-255,
byte-compiled into oracle/_ref/VGsim/IO.pyc.bin) and its export_migrations (src/_BirthDeath.pyx:1743-1754) write the four
files.  Stored: the writers' INPUT arrays (npz) and their output text.  tests/test_writers_parity.py feeds the arrays to
vgsim_b200.io and compares the text byte for byte.  Run here (needs oracle/_ref), commit the outputs."""
import os, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import oracle as O
from scenarios import SCENARIOS
from test_writers_parity import ref_io
RIO = ref_io()

for name, seed, n_iter, gseed in (("s9", 2020, 6000, 7), ("s5", 11, 4000, 3), ("s8", 5, 30000, 9), ("s8hi", 5, 3000, 9)):
    (U, K, S), setup = SCENARIOS[name]
    ref = O.make_reference(U, K, S, seed)
    setup(ref)
    with O.quiet():
        ref.SimulatePopulation(n_iter, n_iter, -1, 200)
        ref.GetGenealogy(gseed)
    tree, times, mut, populations = ref.output_tree_mutations()
    tree, times = np.asarray(tree).copy(), np.asarray(times).copy()
    pops = np.array([populations[t] for t in times], dtype=np.int64)
    with tempfile.TemporaryDirectory() as d:
        RIO.writeGenomeNewick(tree, times, populations, "g", d)
        RIO.writeMutations([list(x) for x in mut], len(tree), "g_mut", d)
        with O.quiet():
            ref.export_migrations("g_mig", d)
        text = {k: open(os.path.join(d, f)).read() for k, f in (("nwk", "g_tree.nwk"), ("pop", "g_sample_population.tsv"),
                                                                 ("mut", "g_mut.tsv"), ("mig", "g_mig.tsv"))}
    mig_rows = [l.split("\t") for l in text["mig"].splitlines()[1:]]
    np.savez_compressed(os.path.join(HERE, "writers_%s.npz" % name), tree=tree, times=times, pops=pops,
                        mut_node=np.array(mut[0], np.int64), mut_AS=np.array(mut[1], np.int64), mut_site=np.array(mut[2], np.int64),
                        mut_DS=np.array(mut[3], np.int64), mut_time=np.array(mut[4], np.float64),
                        mig_node=np.array([int(r[0]) for r in mig_rows], np.int64), mig_time=np.array([float(r[1]) for r in mig_rows]),
                        mig_old=np.array([int(r[2]) for r in mig_rows], np.int64), mig_new=np.array([int(r[3]) for r in mig_rows], np.int64),
                        **{"text_" + k: np.frombuffer(v.encode(), dtype=np.uint8) for k, v in text.items()})
    print(name, "nodes", len(tree), "mutations", len(mut[0]), "migrations", len(mig_rows), {k: len(v) for k, v in text.items()})
