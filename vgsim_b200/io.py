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


def migration_lines(node, time, old_pop, new_pop):
    """Lines of migrations.tsv (reference export_migrations, src/_BirthDeath.pyx:1743-1754): header, then one
    'node<TAB>time<TAB>old deme<TAB>new deme' row per Migrations record, the time printed as Python prints a float."""
    lines = ["Node\tTime\tOld_population\tNew_population\n"]
    for i in range(len(node)):
        lines.append(str(int(node[i])) + '\t' + str(float(time[i])) + '\t' + str(int(old_pop[i])) + '\t' + str(int(new_pop[i])) + "\n")
    return lines


def writeMutations(mut, len_prufer, name_file, file_path):
    fn = (file_path + '/' if file_path is not None else '') + name_file + ".tsv"
    with open(fn, 'w') as f:
        f.writelines(mutation_lines(mut, len_prufer))


# ---------------------------------------------------------------------------------------------------------
# Text parameter formats of the command-line tool (reference src/IO.py:4-142,
# testing/cmd_example/example.{rt,su,pp,mg,st}).  Same function names and return values as the reference's
# readers.  Deliberate differences: '#' comment lines and blank lines after the two header lines are skipped (the
# reference's `next` there is a no-op, so a comment line crashes it), and malformed input raises ValueError
# instead of calling sys.exit.

def _rows(fn, header_lines):
    with open(fn) as f:
        lines = f.read().splitlines()
    head = [l.rstrip().split(" ") for l in lines[:header_lines]]
    body = [l.rstrip().split(" ") for l in lines[header_lines:] if l.strip() and not l.startswith("#")]
    return head, body


def _site_allele(haplotype, site, sites):
    """Base-4 digit of `haplotype` at `site` (site 0 = most significant), reference calculate_allele (src/IO.py:64-68)."""
    return (haplotype // 4 ** (sites - site - 1)) % 4


def read_rates(fn):
    """*.rt -> (bRate[H], dRate[H], sRate[H], mRate[H][U][5]).  Columns `[H] B D S M0 M1 ...`; with `SP` in place of
    `S` the third column is a sampling PROBABILITY (d = D*(1-SP), s = D*SP, src/IO.py:32-37).  A mutation field is
    `rate` or `rate,p1,p2,p3`; the returned entry is [rate, w0..w3] with a 0 inserted at the haplotype's own allele
    (src/IO.py:53-62)."""
    head, body = _rows(fn, 2)
    cols = head[1]
    shift = 1 if cols[0] == "H" else 0
    if len(cols) - shift < 3:
        raise ValueError("At least three rates (B, D, S) are expected")
    sp = cols[2 + shift] == "SP"
    b, d, s, m = [], [], [], []
    for row in body:
        row = row[shift:]
        b.append(float(row[0]))
        if sp:
            d.append(float(row[1]) * (1 - float(row[2])))
            s.append(float(row[1]) * float(row[2]))
        else:
            d.append(float(row[1]))
            s.append(float(row[2]))
        muts = []
        for field in row[3:]:
            parts = field.split(",")
            if len(parts) == 1:
                muts.append([float(parts[0]), 1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0])
            elif len(parts) == 4:
                muts.append([float(x) for x in parts])
            else:
                raise ValueError("Error in mutations!!!")
        m.append(muts)
    H = len(m)
    sites = 0
    while 4 ** sites < H:
        sites += 1
    if 4 ** sites != H:
        raise ValueError("the number of haplotype rows must be a power of 4")
    for h in range(H):
        for u in range(len(m[0])):
            m[h][u].insert(_site_allele(h, u, len(m[0])) + 1, 0)
    return b, d, s, m


def read_susceptibility(fn):
    """*.su -> (susceptibility[H][S] as the file's strings, susceptibility type[H]); columns `[H] T S0 S1 ...`."""
    head, body = _rows(fn, 2)
    shift = 1 if head[1][0] == "H" else 0
    sus, typ = [], []
    for row in body:
        row = row[shift:]
        typ.append(int(row[0]))
        sus.append(row[1:])
    return sus, typ


def read_populations(fn):
    """*.pp -> (sizes, contactDensity, contactAfter, startLD, endLD, samplingMultiplier); columns
    `id size contactDensity [conDenAfterLD,startLD,endLD] [samplingMultiplier]`, the two optional fields in either
    order (src/IO.py:88-129)."""
    _, body = _rows(fn, 2)
    sizes, cd, after, start, end, mult = [], [], [], [], [], []
    for row in body:
        sizes.append(int(row[1]))
        cd.append(float(row[2]))
        extra = [f.split(",") for f in row[3:5]]
        npi = [e for e in extra if len(e) == 3]
        sm = [e for e in extra if len(e) == 1]
        if len(row) == 4:
            if sm:
                after.append(0); start.append(1.0); end.append(1.0); mult.append(float(sm[0][0]))
            elif npi:
                after.append(float(npi[0][0])); start.append(float(npi[0][1])); end.append(float(npi[0][2])); mult.append(1)
        elif len(row) == 5 and npi and sm:
            after.append(float(npi[0][0])); start.append(float(npi[0][1])); end.append(float(npi[0][2]))
            mult.append(float(sm[0][0]))
    return sizes, cd, after, start, end, mult


def read_matrix(fn):
    """*.mg / *.st -> list of rows of floats (one header line)."""
    _, body = _rows(fn, 1)
    return [[float(v) for v in row] for row in body]


# ---------------------------------------------------------------------------------------------------------
# output_epidemiology_timelines (reference src/_BirthDeath.pyx:1765-1847): compartment counts of every deme at the
# grid times i*currentTime/step_num, as a dict or as logs/PID<deme>.log files.  Reference behaviour kept as is:
# the start state is "everybody in group 0 of its deme, one case of haplotype 0 in deme 0" (:1768-1778) whatever
# the configured state was; MULTITYPE rows change nothing (`#TODO` at :1822); a grid point is written after the
# first row whose time is >= that point, at most one point per row (:1823).  Vectorised: the reference walks the
# log in a Python loop.

def epidemiology_timelines(chain, sizes, K, S, H, current_time, step_num):
    """(times[n_pts], sus[n_pts][K][S], inf[n_pts][K][H]) from a 6 x N chain in export_chain_events layout."""
    t = np.asarray(chain[0], dtype=np.float64)
    ty, hap, pop, nhap, npop = (np.asarray(chain[k]).astype(np.int64) for k in range(1, 6))
    n = t.shape[0]
    tp = [i * current_time / step_num for i in range(step_num + 1)]
    # row after which each grid point is written
    rows, j_prev = [], -1
    for k in range(step_num + 1):
        j = max(j_prev + 1, int(np.searchsorted(t, tp[k], side="left")))
        if j >= n:
            break
        rows.append(j)
        j_prev = j
    sus0 = np.zeros((K, S), np.int64)
    sus0[:, 0] = np.asarray(sizes, np.int64)
    inf0 = np.zeros((K, H), np.int64)
    inf0[0, 0] += 1
    sus0[0, 0] -= 1
    n_pts = len(rows)
    if n_pts == 0:
        return [], np.zeros((0, K, S), np.int64), np.zeros((0, K, H), np.int64)
    seg = np.searchsorted(np.asarray(rows), np.arange(n), side="left")     # grid segment each row falls into
    live = seg < n_pts                                                      # rows after the last written point
    d_inf = np.zeros((n_pts, K, H), np.int64)
    d_sus = np.zeros((n_pts, K, S), np.int64)

    def add(arr, mask, p, c, v):
        m = mask & live
        np.add.at(arr, (seg[m], p[m], c[m]), v)

    b, dth, mut, sch, mig = ty == 0, (ty == 1) | (ty == 2), ty == 3, ty == 4, ty == 5
    add(d_inf, b, pop, hap, 1);      add(d_sus, b, pop, nhap, -1)
    add(d_inf, dth, pop, hap, -1);   add(d_sus, dth, pop, nhap, 1)
    add(d_inf, mut, pop, hap, -1);   add(d_inf, mut, pop, nhap, 1)
    add(d_sus, sch, pop, hap, -1);   add(d_sus, sch, pop, nhap, 1)
    add(d_sus, mig, npop, nhap, -1); add(d_inf, mig, npop, hap, 1)
    return [tp[k] for k in range(n_pts)], sus0 + np.cumsum(d_sus, axis=0), inf0 + np.cumsum(d_inf, axis=0)


def timelines_as_dict(times, sus, inf):
    """The reference's return value: {"time": [...], "P<i>": {"S<j>": [...], "H<j>": [...]}}."""
    K, S, H = sus.shape[1], sus.shape[2], inf.shape[2]
    log = {"time": list(times)}
    for i in range(K):
        log["P" + str(i)] = {}
        for j in range(S):
            log["P" + str(i)]["S" + str(j)] = list(sus[:, i, j])
        for j in range(H):
            log["P" + str(i)]["H" + str(j)] = list(inf[:, i, j])
    return log


def write_timelines(times, sus, inf, directory="logs"):
    """logs/PID<i>.log in the reference's layout (:1779-1790, 1824-1832)."""
    import os
    if not os.path.isdir(directory):
        os.mkdir(directory)
    K, S, H = sus.shape[1], sus.shape[2], inf.shape[2]
    for i in range(K):
        with open(os.path.join(directory, "PID" + str(i) + ".log"), "w") as f:
            f.write("time" + "".join(" S" + str(j) for j in range(S)) + "".join(" H" + str(j) for j in range(H)) + "\n")
            for k in range(len(times)):
                f.write(str(times[k]) + " " + "".join(str(v) + " " for v in sus[k, i]) + "".join(str(v) + " " for v in inf[k, i]) + "\n")


def write_settings(base, hap_names, a):
    """<base>.rt/.pp/.mg/.su/.st from the engine's parameter arrays (`BirthDeathModel.param_arrays()`), in the layout of
    the reference's export_settings (src/_BirthDeath.pyx:1861-1905): numbers are written with str(), every mutation
    field is `rate,w1,w2,w3` (the weights of the three OTHER alleles), matrix rows end with a space."""
    H, U = a["mRate"].shape
    K, S = a["sizes"].shape[0], a["sigma"].shape[1]
    with open(base + ".rt", "w") as f:
        f.write("#Rates_format_version 0.0.1\nH B D S" + "".join(" M" + str(u) for u in range(U)) + "\n")
        for h in range(H):
            f.write(hap_names[h] + " " + str(a["b"][h]) + " " + str(a["d"][h]) + " " + str(a["s"][h]) + " ")
            for u in range(U):
                f.write(str(a["mRate"][h, u]) + "," + ",".join(str(a["hapMutType"][h, u, k]) for k in range(3)) + " ")
            f.write("\n")
    with open(base + ".pp", "w") as f:
        f.write("#Population_format_version 0.0.1\nid size contactDensity conDenAfterLD startLD endLD samplingMulriplier\n")
        for p in range(K):
            f.write(str(p) + " " + str(a["sizes"][p]) + " " + str(a["cd"][p]) + " " + str(a["cdAfter"][p]) + "," +
                    str(a["startLD"][p]) + "," + str(a["endLD"][p]) + " " + str(a["sm"][p]) + "\n")
    with open(base + ".mg", "w") as f:
        f.write("#Migration_format_version 0.0.1\n")
        for p in range(K):
            f.write("".join(str(a["m"][p, q]) + " " for q in range(K)) + "\n")
    with open(base + ".su", "w") as f:
        f.write("#Susceptibility_format_version 0.0.1\nH T" + "".join(" S" + str(g) for g in range(S)) + "\n")
        for h in range(H):
            f.write(hap_names[h] + " " + str(a["suscType"][h]) + "".join(" " + str(a["sigma"][h, g]) for g in range(S)) + "\n")
    with open(base + ".st", "w") as f:
        f.write("#Susceptibility_format_version 0.0.1\n")
        for g in range(S):
            f.write("".join(str(a["T"][g, g2]) + " " for g2 in range(S)) + "\n")
