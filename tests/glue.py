"""Test-side glue that turns hot-path results (per-pair conreci) into the reference's
containers and text outputs, so the oracle path and the CUDA path are compared through
exactly the same code.  Follows Arcs.cpp:1280-1285,1309-1319 (imap), 1378-1435
(pairContigs, via the oracle or the CUDA library), 1475-1526 + Arcs.h:185-229 (graph text)."""
import numpy as np

import oracle_lib as O


def contig_ends(contigs, min_size, end_length):
    """getContigKmers (Arcs.cpp:1046-1094): -> (ends [(seq, conreci)], names_by_contig, contig_of_conreci)
    conreci 2i+1 / 2i+2 = head / tail of the i-th contig with len >= min_size"""
    ends, names = [], []
    for name, seq in contigs:
        if len(seq) >= min_size:
            i = len(names)
            names.append(name)
            cut = O.lib().arks_oracle_end_cutoff(len(seq), end_length)
            ends.append((seq[:cut], 2 * i + 1))
            ends.append((seq[len(seq) - cut:], 2 * i + 2))
    return ends, names


def name_ids(names):
    """imap / pmap are keyed by contig NAME (Arcs.h:106-115): duplicate names merge.
    -> (id_of_contig int array, unique names list)"""
    first, ids, uniq = {}, [], []
    for n in names:
        if n not in first:
            first[n] = len(uniq)
            uniq.append(n)
        ids.append(first[n])
    return np.array(ids, dtype=np.uint32), uniq


def lex_rank(uniq_names):
    order = sorted(range(len(uniq_names)), key=lambda i: uniq_names[i].encode())
    rank = np.zeros(len(uniq_names), dtype=np.uint32)
    for r, i in enumerate(order):
        rank[i] = r
    return rank


def imap_rows(barcodes, conreci, names):
    """-> dict barcode -> {contig_id: [head, tail]} from stored pairs"""
    ids, uniq = name_ids(names)
    imap = {}
    for b, c in zip(barcodes, conreci):
        if c == 0:
            continue
        cid = int(ids[(c - 1) // 2])
        ht = imap.setdefault(b, {}).setdefault(cid, [0, 0])
        ht[0 if (c & 1) else 1] += 1
    return imap, uniq


def imap_to_arrays(imap, mult):
    """rows sorted by barcode id (ids assigned in sorted barcode order)"""
    bnames = sorted(imap.keys())
    bid = {b: i for i, b in enumerate(bnames)}
    rows = [(bid[b], c, ht[0], ht[1]) for b in bnames for c, ht in sorted(imap[b].items())]
    a = np.array(rows, dtype=np.uint32).reshape(-1, 4)
    m = np.array([mult.get(b, 0) for b in bnames], dtype=np.int32)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), a[:, 3].copy(), m, bnames


def gv_text(a, b, counts, uniq_names, rank, min_links, error_percent, edge_fn=O.edge):
    """createGraph + write_graphviz text; rows must be in (rank[a], rank[b]) order"""
    order = sorted(range(len(a)), key=lambda i: (int(rank[a[i]]), int(rank[b[i]])))
    vid, lines_v, lines_e = {}, [], []
    for i in order:
        ok, orient, weight = edge_fn(counts[i], min_links, error_percent)
        if not ok:
            continue
        for c in (int(a[i]), int(b[i])):
            if c not in vid:
                vid[c] = len(vid)
                lines_v.append("%d [id=%s];\n" % (vid[c], uniq_names[c]))
        lines_e.append("%d--%d [label=%d, weight=%d];\n" % (vid[int(a[i])], vid[int(b[i])], orient, weight))
    return "graph G {\n" + "".join(lines_v) + "".join(lines_e) + "}\n"
