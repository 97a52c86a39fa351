"""Newick / TSV writers fed from the device trees (reference src/IO.py:144-255).

Same file names and byte-for-byte the same text as the reference writers, but iterative (the
reference recurses once per tree level and dies on deep trees) and linear-time (the reference's
writeMutations does an O(n^2) ``list.index`` scan).
"""
import numpy as np

_ALLELES = "ATCG"


def newick_string(parent, times):
    """'(left,right)node:branch' text of the tree given the parent array (root = -1).

    Children are ordered by node id like the reference's find_children (src/IO.py:212-222); every
    node is labelled with its id and the branch length is time(node) - time(parent)."""
    parent = np.asarray(parent, dtype=np.int64)
    times = [float(t) for t in np.asarray(times, dtype=np.float64)]
    n = len(parent)
    left = [-1] * n
    right = [-1] * n
    root = -1
    for i in range(n):
        p = int(parent[i])
        if p < 0:
            if root < 0:
                root = i
            continue
        if left[p] < 0:
            left[p] = i
        else:
            right[p] = i
    out = []
    # iterative traversal; stack entries: (node, stage)
    stack = [(root, 0)]
    while stack:
        node, stage = stack.pop()
        if left[node] < 0:
            base = times[int(parent[node])] if parent[node] >= 0 else times[node]
            out.append('{0}:{1}'.format(node, times[node] - base))
            continue
        if stage == 0:
            out.append('(')
            stack.append((node, 1))
            stack.append((left[node], 0))
        elif stage == 1:
            out.append(',')
            stack.append((node, 2))
            stack.append((right[node], 0))
        else:
            base = times[int(parent[node])] if parent[node] >= 0 else times[node]
            out.append('){0}:{1}'.format(node, times[node] - base))
    return ''.join(out), root, left, right


def population_lines(root, left, right, pops):
    """pre-order 'node<TAB>deme' lines (Vertex.write_population, src/IO.py:199-200,209)."""
    out = []
    stack = [root]
    while stack:
        node = stack.pop()
        out.append('{0}\t{1}\n'.format(node, int(pops[node])))
        if left[node] >= 0:
            stack.append(right[node])
            stack.append(left[node])
    return ''.join(out)


def writeGenomeNewick(pruferSeq, times, populations, name_file, file_path):
    """`populations` is the per-node deme array (tree_pop); the reference passes a {time: deme} dict
    built from the event log, which is wrong for tau-phase nodes."""
    text, root, left, right = newick_string(pruferSeq, times)
    if isinstance(populations, dict):
        populations = [populations[float(t)] for t in times]
    if file_path is not None:
        nwk, pop = file_path + '/' + name_file + '_tree.nwk', file_path + '/' + name_file + '_sample_population.tsv'
    elif name_file is not None:
        nwk, pop = name_file + '_tree.nwk', name_file + '_sample_population.tsv'
    else:
        nwk, pop = 'tree.nwk', 'sample_population.tsv'
    with open(nwk, 'w') as f:
        f.write(text)
        f.write(';')
    with open(pop, 'w') as f:
        f.write(population_lines(root, left, right, populations))


def mutation_lines(mut, len_prufer):
    """Lines of mutations.tsv.  mut = [nodeId, AS, site, DS, time] lists (output_tree_mutations).

    Reproduces the reference text exactly, including its quirk that a node carrying several
    mutations repeats its FIRST mutation once per occurrence (``mut[0].index(nodeId)``,
    src/IO.py:152-160)."""
    first = {}
    count = {}
    for j, node in enumerate(mut[0]):
        if node not in first:
            first[node] = j
            count[node] = 0
        count[node] += 1
    lines = []
    for i in sorted(first):
        if i >= len_prufer:
            continue
        j = first[i]
        one = _ALLELES[mut[1][j]] + str(mut[2][j]) + _ALLELES[mut[3][j]]
        lines.append(str(i) + '\t' + ','.join([one] * count[i]) + '\n')
    return lines


def writeMutations(mut, len_prufer, name_file, file_path):
    fn = (file_path + '/' if file_path is not None else '') + name_file + ".tsv"
    with open(fn, 'w') as f:
        f.writelines(mutation_lines(mut, len_prufer))
