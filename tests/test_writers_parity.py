"""Output-format parity: the Newick / sample-population / mutations / migrations text written by
vgsim_b200 equals, byte for byte, what the reference's own writers produce (src/IO.py:144-255, export_migrations
src/_BirthDeath.pyx:1743-1754).

Three layers:
  * golden (always runs, CPU and GPU box): tests/golden/writers_*.npz hold the writers' INPUT arrays of runs of the
    unmodified reference engine and the text the reference's writers made of them (tests/golden/make_writers_golden.py);
  * live (where oracle/_ref exists): the reference engine + reference writers on fresh seeds vs our writers on the
    same arrays;
  * device (-m gpu): trees, mutations and migrations from the CUDA path, exported through `Simulator.export_*`, vs the
    reference's writers (byte-compiled into oracle/_ref/VGsim/IO.pyc.bin by oracle/build_ref.py) fed with the same arrays.
"""
import glob
import os
import sys

import numpy as np
import pytest

from scenarios import SCENARIOS
from vgsim_b200 import io as vio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF_IO = os.path.exists(os.path.join(REF_DIR, "VGsim", "IO.pyc.bin"))


def ref_io():
    """The reference's src/IO.py, loaded from the sourceless bytecode oracle/build_ref.py leaves in oracle/_ref."""
    import importlib.machinery
    import importlib.util
    sys.setrecursionlimit(100000)   # the reference's Vertex builds itself recursively, one level per tree level
    if "vgsim_reference_IO" in sys.modules:
        return sys.modules["vgsim_reference_IO"]
    loader = importlib.machinery.SourcelessFileLoader("vgsim_reference_IO", os.path.join(REF_DIR, "VGsim", "IO.pyc.bin"))
    spec = importlib.util.spec_from_loader("vgsim_reference_IO", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules["vgsim_reference_IO"] = mod
    return mod


def ours(tmp, tree, times, pops, mut, mig):
    d = str(tmp)
    vio.writeGenomeNewick(tree, times, pops, "o", d)
    vio.writeMutations([list(x) for x in mut], len(tree), "o_mut", d)
    with open(os.path.join(d, "o_mig.tsv"), "w") as f:
        f.writelines(vio.migration_lines(*mig))
    return {k: open(os.path.join(d, f)).read() for k, f in (("nwk", "o_tree.nwk"), ("pop", "o_sample_population.tsv"),
                                                            ("mut", "o_mut.tsv"), ("mig", "o_mig.tsv"))}


@pytest.mark.parametrize("fn", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "writers_*.npz"))),
                         ids=lambda f: os.path.basename(f)[:-4])
def test_writers_match_reference_golden(fn, tmp_path):
    g = np.load(fn)
    mut = [g["mut_node"].tolist(), g["mut_AS"].tolist(), g["mut_site"].tolist(), g["mut_DS"].tolist(), g["mut_time"].tolist()]
    got = ours(tmp_path, g["tree"], g["times"], g["pops"], mut, (g["mig_node"], g["mig_time"], g["mig_old"], g["mig_new"]))
    for k in ("nwk", "pop", "mut", "mig"):
        want = bytes(g["text_" + k]).decode()
        assert got[k] == want, (k, got[k][:200], want[:200])
    # the {time: deme} dict the reference passes works too (node times are unique)
    vio.writeGenomeNewick(g["tree"], g["times"], {float(t): int(p) for t, p in zip(g["times"], g["pops"])}, "d", str(tmp_path))
    assert open(os.path.join(str(tmp_path), "d_sample_population.tsv")).read() == bytes(g["text_pop"]).decode()


@pytest.mark.skipif(not HAVE_REF_IO, reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("name,seed,n_iter", [("s9", 31, 5000), ("s8hi", 12, 2500), ("s6", 8, 20000), ("s1", 77, 3000)])
def test_writers_match_reference_live(name, seed, n_iter, tmp_path):
    from oracle import oracle as O
    RIO = ref_io()
    (U, K, S), setup = SCENARIOS[name]
    ref = O.make_reference(U, K, S, seed)
    setup(ref)
    with O.quiet():
        ref.SimulatePopulation(n_iter, n_iter, -1, 200)
        ref.GetGenealogy(seed + 1)
    tree, times, mut, populations = ref.output_tree_mutations()
    tree, times = np.asarray(tree).copy(), np.asarray(times).copy()
    d = str(tmp_path)
    RIO.writeGenomeNewick(tree, times, populations, "r", d)
    RIO.writeMutations([list(x) for x in mut], len(tree), "r_mut", d)
    with O.quiet():
        ref.export_migrations("r_mig", d)
    rows = [l.split("\t") for l in open(os.path.join(d, "r_mig.tsv")).read().splitlines()[1:]]
    mig = ([int(r[0]) for r in rows], [float(r[1]) for r in rows], [int(r[2]) for r in rows], [int(r[3]) for r in rows])
    got = ours(tmp_path, tree, times, populations, mut, mig)
    assert len(tree) > 10
    for k, f in (("nwk", "r_tree.nwk"), ("pop", "r_sample_population.tsv"), ("mut", "r_mut.tsv"), ("mig", "r_mig.tsv")):
        assert got[k] == open(os.path.join(d, f)).read(), k


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF_IO, reason="oracle/_ref (reference build) not present")
@pytest.mark.parametrize("name,n_iter", [("s9", 6000), ("s8hi", 2500), ("example", 60000)])
def test_device_exports_match_reference_writers(name, n_iter, tmp_path):
    """Device trees through Simulator.export_newick / export_mutations / export_migrations vs the reference's writers on
    the same arrays."""
    from vgsim_b200 import Simulator
    RIO = ref_io()
    (U, K, S), setup = SCENARIOS[name]
    sim = Simulator(U, K, S, seed=4242, verbose=False, replicates=3)
    setup(sim.simulation)
    sim.simulate(n_iter, method="direct")
    sim.genealogy(seed=99)
    d = str(tmp_path)
    for r in range(3):
        tree, times, mut, populations = sim.simulation.output_tree_mutations(r)
        if len(tree) < 3:
            continue
        sim.export_newick("dev%d" % r, d, replicate=r)
        sim.export_mutations("dev%d_mut" % r, d, replicate=r)
        sim.export_migrations("dev%d_mig" % r, d, replicate=r)
        RIO.writeGenomeNewick(np.asarray(tree), np.asarray(times), populations, "ref%d" % r, d)
        RIO.writeMutations([list(x) for x in mut], len(tree), "ref%d_mut" % r, d)
        for a, b in (("dev%d_tree.nwk", "ref%d_tree.nwk"), ("dev%d_sample_population.tsv", "ref%d_sample_population.tsv"),
                     ("dev%d_mut.tsv", "ref%d_mut.tsv")):
            assert open(os.path.join(d, a % r)).read() == open(os.path.join(d, b % r)).read(), (a % r)
        # migrations: the reference's writer is an engine method (src/_BirthDeath.pyx:1743-1754); its format, restated
        node, t, oldp, newp = sim.simulation._handle.get_migrations(r)
        want = "Node\tTime\tOld_population\tNew_population\n" + "".join(
            str(int(node[i])) + "\t" + str(float(t[i])) + "\t" + str(int(oldp[i])) + "\t" + str(int(newp[i])) + "\n"
            for i in range(len(node)))
        assert open(os.path.join(d, "dev%d_mig.tsv" % r)).read() == want
