"""How far do the stated deviations D1-D5 of the LASTZ-restatement (oracle/lastz_oracle.c, the spec the CUDA path reproduces bit
for bit) move mimeo's output? Runs the spec oracle and the sequential second statement (oracle/lastz_faithful.c: LASTZ's visiting
order, diagonal memory of failed extensions, floating-point entropy, row-wise y-drop, path-based cover test and barriers) on
the same genomes and compares what mimeo would write: alignment rows, hit counts, GFF3 text, base-level Jaccard of the annotated
bases. CPU only.   python scripts/spec_deviation.py > profiles/r2_spec_deviation.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import annot_oracle as ao       # noqa: E402
from oracle import lastz_oracle as lo       # noqa: E402
from tests.helpers import synth_c3, synth_genome   # noqa: E402


def annotated(gff):
    """{scaffold: sorted list of (start, end)} of a GFF3 text."""
    segs = {}
    for ln in gff.splitlines():
        if ln.startswith('#') or not ln.strip():
            continue
        f = ln.split('\t')
        segs.setdefault(f[0], []).append((int(f[3]), int(f[4])))
    return segs


def jaccard(a, b):
    inter = union = 0
    for name in set(a) | set(b):
        n = max([e for _, e in a.get(name, [])] + [e for _, e in b.get(name, [])] + [1]) + 1
        ma, mb = np.zeros(n, bool), np.zeros(n, bool)
        for s, e in a.get(name, []):
            ma[s:e] = True
        for s, e in b.get(name, []):
            mb[s:e] = True
        inter += int((ma & mb).sum()); union += int((ma | mb).sum())
    return inter / union if union else 1.0


def self_outputs(general, enc, minIdt, minLen, minCov, intraCov):
    tab, intra = ao.TAB_HEADER, ao.TAB_HEADER
    for a in sorted(enc):
        for b in sorted(enc):
            rows = ''.join(ao.filter_lastz_general(general[(a, b)], minLen, minIdt))
            if a == b:
                intra += rows
            else:
                tab += rows
    gff = ao.self_gff3(tab.splitlines(True), intra.splitlines(True), {n: len(c) for n, c in enc.items()}, minCov, intraCov, minLen, 'Self_Repeat', 'Self_Repeat')
    return tab, intra, gff


def map_outputs(general, ea, eb, minIdt, minLen):
    tab = ao.TAB_HEADER
    for a in sorted(ea):
        for b in sorted(eb):
            tab += ''.join(ao.filter_lastz_general(general[(a, b)], minLen, minIdt))
    hits = ao.import_align_rows(tab.splitlines(True), 'BHit', minLen, minIdt)
    return tab, ''.join(ao.write_gff_lines(hits, sorted((n, str(len(c))) for n, c in ea.items()), 'BHit'))


def compare(name, ea, eb, mode, workers):
    t0 = time.time()
    st_s, st_f = lo.Stats(), lo.Stats()
    spec = lo.general_all_pairs(ea, eb, 3000, workers, st_s)
    t1 = time.time()
    faith = lo.general_all_pairs(ea, eb, 3000, workers, st_f, faithful=True)
    t2 = time.time()
    an, bn = sorted(ea), sorted(eb if eb is not None else ea)
    rs, rf = lo.general_rows(spec, an, bn), lo.general_rows(faith, an, bn)
    out = {'config': name, 'mode': mode, 'pairs': len(spec), 'seconds': {'spec': round(t1 - t0, 1), 'sequential': round(t2 - t1, 1)},
           'alignment_rows': {'spec': len(rs), 'sequential': len(rf), 'identical': len(rs & rf), 'count_ratio': len(rs) / max(len(rf), 1)},
           'hsps_kept': {'spec': st_s.hsps_kept, 'sequential': st_f.hsps_kept},
           'gapped_cells': {'spec': st_s.gapped_cells, 'sequential': st_f.gapped_cells}, 'by_minIdt': {}}
    for minIdt in (80, 90):
        if mode == 'self':
            ts, is_, gs = self_outputs(spec, ea, minIdt, 100, 3, 4)
            tf, if_, gf = self_outputs(faith, ea, minIdt, 100, 3, 4)
            rows_s, rows_f = ts.count('\n') + is_.count('\n') - 2, tf.count('\n') + if_.count('\n') - 2
            tab_equal = ts == tf and is_ == if_
        else:
            ts, gs = map_outputs(spec, ea, eb, minIdt, 100)
            tf, gf = map_outputs(faith, ea, eb, minIdt, 100)
            rows_s, rows_f = ts.count('\n') - 1, tf.count('\n') - 1
            tab_equal = ts == tf
        out['by_minIdt'][str(minIdt)] = {
            'tab_rows': {'spec': rows_s, 'sequential': rows_f, 'ratio': rows_s / max(rows_f, 1)}, 'tab_identical': tab_equal,
            'gff3_identical': gs == gf, 'gff3_rows': {'spec': gs.count('\n'), 'sequential': gf.count('\n')},
            'annotated_bases_jaccard': jaccard(annotated(gs), annotated(gf))}
    return out


def main():
    workers = os.cpu_count() or 1
    res = {'what': 'spec (order-independent LASTZ-restatement, oracle/lastz_oracle.c = what the CUDA path computes) vs sequential second statement '
                   '(oracle/lastz_faithful.c); neither is LASTZ itself (absent from every box): PARITY UNPINNED', 'cores': workers, 'runs': []}
    g = synth_genome(1001, 10, 500_000, 20, copies=(5, 30), fam_len=(300, 3000), sub=0.106, indel=0.005)
    enc = {k: lo.encode(v) for k, v in g.items()}
    res['runs'].append(compare('C1: mimeo self, 5 Mbp (10 x 500 kbp), seed 1001, minLen 100 minCov 3 intraCov 4 strictSelf', enc, None, 'self', workers))
    a, b = synth_c3(1003, nscaf=8, scaf_len=500_000)
    ea, eb = {k: lo.encode(v) for k, v in a.items()}, {k: lo.encode(v) for k, v in b.items()}
    res['runs'].append(compare('C3 scaled to 4 Mbp x 4 Mbp (8 x 500 kbp; same generator, seed 1003): mimeo map, minLen 100', ea, eb, 'map', workers))
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
