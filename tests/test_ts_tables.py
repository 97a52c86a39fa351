"""export_ts_tables: the rows of the reference's export_ts (src/_BirthDeath.pyx:1909-1946) without tskit.

CPU test: the engine's table builder is fed a hand-made genealogy through a stand-in handle and compared with a literal
restatement of the reference's loops (row order before `tc.sort()`).  The GPU test does the same on a device genealogy."""
import numpy as np
import pytest

from vgsim_b200._engine import BirthDeathModel as Eng


def reference_rows(tree, tree_pop, times, mutations, migrations, genome_length, sites_position, pop_num, s_counter):
    """Literal restatement of export_ts: lists of the add_row argument tuples, table by table."""
    rows = {"migrations": [], "edges": [], "nodes": [], "sites": [], "mutations": []}
    for (node, t, old, new) in migrations:
        rows["migrations"].append((0.0, 1.0, node, old, new, times[0] - t))
    for i in range(2 * s_counter - 2):
        rows["edges"].append((0.0, genome_length, tree[i], i))
    child_or_parent = [1 for _ in range(2 * s_counter - 1)]
    for i in range(2 * s_counter - 2):
        child_or_parent[tree[i]] = 0
    for i in range(2 * s_counter - 1):
        rows["nodes"].append((child_or_parent[i], times[0] - times[i], tree_pop[i]))
    for p in sites_position:
        if p == 0:
            rows["sites"].append((p + 1, 'A'))
        elif p == genome_length:
            rows["sites"].append((p - 1, 'A'))
        else:
            rows["sites"].append((p, 'A'))
    allele = ['A', 'T', 'C', 'G']
    for (node, DS, AS, site, t) in mutations:          # get_mutation returns (nodeId, DS, AS, site, time)
        rows["mutations"].append((site, node, allele[DS], times[0] - t))
    return rows


def assert_tables_match(tb, rows, genome_length, pop_num):
    assert tb["sequence_length"] == genome_length and tb["populations"] == pop_num
    m = tb["migrations"]
    got = list(zip(m["left"], m["right"], m["node"], m["source"], m["dest"], m["time"]))
    assert got == [tuple(map(float, r[:2])) + tuple(r[2:5]) + (r[5],) for r in rows["migrations"]]
    e = tb["edges"]
    assert list(zip(e["left"], e["right"], e["parent"], e["child"])) == rows["edges"]
    n = tb["nodes"]
    assert list(zip(n["flags"], n["time"], n["population"])) == rows["nodes"]
    s = tb["sites"]
    assert list(zip(s["position"], s["ancestral_state"])) == rows["sites"]
    mu = tb["mutations"]
    assert list(zip(mu["site"], mu["node"], mu["derived_state"], mu["time"])) == rows["mutations"]


class _StandIn:
    """Handle stand-in: a 3-sample genealogy ((0,1)3,2)4 with one mutation and one migration."""
    tree = np.array([3, 3, 4, 4, -1], np.int64)
    pop = np.array([0, 1, 1, 1, 0], np.int64)
    times = np.array([9.0, 8.5, 7.0, 4.0, 1.5])

    def get_tree(self, replicate=0):
        return self.tree, self.pop, self.times

    def get_mutations(self, replicate=0):       # (node, AS, DS, site, time)
        return (np.array([1], np.int64), np.array([0], np.int64), np.array([2], np.int64), np.array([1], np.int64), np.array([6.0]))

    def get_migrations(self, replicate=0):      # (node, time, oldPop, newPop)
        return (np.array([0], np.int64), np.array([5.0]), np.array([1], np.int64), np.array([0], np.int64))


def test_tables_follow_the_reference_loops_cpu():
    e = Eng(2, 2, 1, 1, False, False, 1000, 0.0)
    e._handle = _StandIn()
    e._genealogy_done = True
    tb = e.export_ts_tables()
    h = e._handle
    rows = reference_rows(h.tree.tolist(), h.pop.tolist(), h.times.tolist(), [(1, 2, 0, 1, 6.0)], [(0, 5.0, 1, 0)],
                          e.genome_length, e.sitesPosition.tolist(), e.popNum, 3)
    assert_tables_match(tb, rows, e.genome_length, e.popNum)
    assert tb["sites"]["position"].tolist() == [1.0, 999.0]      # 0 and genome_length are moved inside (0, L)
    assert tb["nodes"]["flags"].tolist() == [1, 1, 1, 0, 0]
    try:
        import tskit
        tskit.TableCollection       # (the out-of-tree reference build runs with a stand-in module of that name)
    except (ImportError, AttributeError):
        with pytest.raises(ImportError, match="export_ts_tables"):
            e.export_ts()
    else:
        ts = e.export_ts()
        assert ts.num_samples == 3 and ts.num_mutations == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["s8", "s9"])
def test_tables_follow_the_reference_loops_device(name):
    from test_gpu_tau import make_engine
    e = make_engine(name, 77)
    e.SimulatePopulation(20000, 300, -1, 200)
    e.GetGenealogy(5)
    tree, times = e.get_tree()
    pop = e.get_tree_populations()
    node, AS, DS, site, mt = e.get_mutations()
    gnode, gt, gold, gnew = e.get_migrations()
    s_counter = (len(tree) + 1) // 2
    rows = reference_rows(tree.tolist(), pop.tolist(), times.tolist(),
                          list(zip(node.tolist(), DS.tolist(), AS.tolist(), site.tolist(), mt.tolist())),
                          list(zip(gnode.tolist(), gt.tolist(), gold.tolist(), gnew.tolist())),
                          e.genome_length, e.sitesPosition.tolist(), e.popNum, s_counter)
    assert len(rows["nodes"]) == len(tree) and len(rows["mutations"]) + len(rows["migrations"]) > 0
    assert_tables_match(e.export_ts_tables(), rows, e.genome_length, e.popNum)
