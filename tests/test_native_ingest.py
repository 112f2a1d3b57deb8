"""CPU tests of the native text ingest: the C++ FASTA reader and the .tab BED projection behind the
C ABI (`mb2_fasta_read`, `mb2_tab_project`) against plain-Python statements of the same rules and the oracle's awk
projection, single- and multi-threaded, on ragged and empty inputs. No GPU needed: these entry points are host code."""
import os

import numpy as np
import pytest

from oracle import annot_oracle as ao
from tests.helpers import read_golden


def py_read_fasta(data: bytes):
    """Reference statement: a record starts at '>' in column one; id = first word; body minus \\n, \\r and blanks."""
    out = []
    starts = [i for i in range(len(data)) if data[i:i + 1] == b'>' and (i == 0 or data[i - 1:i] == b'\n')]
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(data)
        nl = data.find(b'\n', s, e)
        if nl < 0:
            nl = e
        header = data[s + 1:nl].decode().rstrip('\r')
        body = data[nl + 1:e]
        seq = bytes(c for c in body if c not in (10, 13, 32))
        out.append((header.split()[0] if header.split() else '', header, seq))
    return out


def rand_fasta(rng, nrec, maxlen, width, crlf=False):
    eol = b'\r\n' if crlf else b'\n'
    chunks = []
    for r in range(nrec):
        n = int(rng.integers(0, maxlen))
        seq = bytes(rng.choice(np.frombuffer(b'ACGTNacgtn', dtype=np.uint8), n).tolist())
        chunks.append(b'>rec%d some description %d' % (r, n) + eol)
        for i in range(0, n, width):
            chunks.append(seq[i:i + width] + eol)
        if rng.random() < 0.3:
            chunks.append(eol)
    return b''.join(chunks)


@pytest.mark.parametrize('threads', [1, 4])
def test_fasta_reader_matches_plain_statement(tmp_path, threads):
    from mimeo_b200 import fasta
    rng = np.random.default_rng(11)
    cases = [b'', b'>only_header', b'>a\nACGT', b'>a b c\r\nAC GT\r\n\r\n>b\nNN\n>\nAA\n', b'no header line\nACGT\n>x\nAC\n',
             rand_fasta(rng, 7, 5000, 60), rand_fasta(rng, 3, 3_000_000, 70, crlf=True), rand_fasta(rng, 40, 200, 13)]
    for k, data in enumerate(cases):
        p = tmp_path / f'c{k}.fa'
        p.write_bytes(data)
        got = [(i, h, bytes(s)) for i, h, s in fasta.read_fasta(str(p), nthreads=threads)]
        assert got == py_read_fasta(data), k


def test_fasta_reader_missing_file_raises(tmp_path):
    from mimeo_b200 import fasta, _lib
    with pytest.raises(_lib.Mb2Error):
        fasta.read_fasta(str(tmp_path / 'nope.fa'))


def py_project(lines):
    """awk '!/^#/ {print $1,$3,$4;}' for well-formed lines, blank lines dropped."""
    rows = []
    for ln in lines:
        if ln.startswith('#'):
            continue
        f = ln.split()
        if f:
            rows.append((f[0], int(f[2]), int(f[3])))
    return rows


@pytest.mark.parametrize('threads', [1, 3, 16])
def test_tab_projection_matches_awk_statement(tmp_path, threads):
    from mimeo_b200 import engine
    rng = np.random.default_rng(12)
    names = ['s10', 'S3', 's2', 'chr_long_name.1']
    lines = ['#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity']
    for _ in range(200_000):
        s = int(rng.integers(0, 10**6))
        lines.append('%s\t+\t%d\t%d\tq\t-\t1\t2\t3000\t%.1f' % (names[int(rng.integers(0, 4))], s, s + int(rng.integers(0, 5000)), rng.uniform(60, 100)))
        if rng.random() < 0.001:
            lines.append('' if rng.random() < 0.5 else '# lastz end-of-file')
    lines.append('  s2   +  7   9')                      # awk splits on runs of blanks
    p = tmp_path / 'big.tab'
    p.write_text('\n'.join(lines))                        # no trailing newline on purpose
    got_names, ids, start, end = engine.parse_tab_hits(str(p), nthreads=threads)
    want = py_project(lines)
    assert len(ids) == len(want)
    assert [got_names[i] for i in ids[:2000].tolist()] == [w[0] for w in want[:2000]]
    assert np.array_equal(np.asarray([got_names[i] for i in ids.tolist()]), np.asarray([w[0] for w in want]))
    assert start.tolist() == [w[1] for w in want] and end.tolist() == [w[2] for w in want]
    # names come out in order of first appearance
    seen = list(dict.fromkeys(w[0] for w in want))
    assert got_names == seen


@pytest.mark.parametrize('case', ['cov_dense', 'cov_order', 'cov_nointra'])
def test_tab_projection_of_goldens_equals_oracle_projection(tmp_path, case):
    """Same triples as the oracle's restatement of the reference's awk projection on the reference-generated goldens."""
    from mimeo_b200 import engine
    txt = read_golden(f'{case}.tab')
    p = tmp_path / 'g.tab'
    p.write_text(txt)
    names, ids, start, end = engine.parse_tab_hits(str(p))
    got = ['%s\t%d\t%d' % (names[i], s, e) for i, s, e in zip(ids.tolist(), start.tolist(), end.tolist())]
    want = [r.rstrip('\n') for r in ao.project_bed(txt.splitlines())]
    assert [g.split('\t') for g in got] == [w.split() for w in want]


def test_tab_projection_errors(tmp_path):
    from mimeo_b200 import engine
    for bad in ('a\t+\t5\n', 'a\t+\tx\t9\n', 'a\t+\t5\t9.5\n'):
        p = tmp_path / 'bad.tab'
        p.write_text('#h\n' + bad)
        with pytest.raises(RuntimeError, match='line 2'):
            engine.parse_tab_hits(str(p))
    p = tmp_path / 'empty.tab'
    p.write_text('#only a header\n')
    names, ids, start, end = engine.parse_tab_hits(str(p))
    assert names == [] and len(ids) == 0


def test_native_tab_formatter_matches_plain_statement():
    """mb2_format_tab against the plain-Python statement of the awk/sed/sort filter: ties, rounding of the printed identity at
    the threshold, zero columns, duplicated rows, names that are prefixes of each other."""
    from mimeo_b200 import align as A
    from tests.helpers import py_tab_blocks
    rng = np.random.default_rng(3)
    names = ['s10', 'S3', 's2', 's', 'zz_long.name']
    for trial in range(25):
        n = int(rng.integers(0, 3000))
        nc = rng.integers(0, 5000, n)
        nm = (nc * rng.uniform(0.5, 1, n)).astype(np.int64)
        s1 = rng.integers(1, 2000, n)
        hits = {'t_id': rng.integers(0, 5, n).astype(np.int32), 'q_id': rng.integers(0, 5, n).astype(np.int32),
                'strand': rng.integers(0, 2, n).astype(np.int32), 'start1': s1.astype(np.int32),
                'end1': (s1 + rng.integers(0, 300, n)).astype(np.int32), 'start2': rng.integers(1, 9999, n).astype(np.int32),
                'end2': rng.integers(1, 99999, n).astype(np.int32), 'score': rng.integers(3000, 10**6, n).astype(np.int32),
                'nmatch': nm.astype(np.int32), 'ncols': nc.astype(np.int32)}
        if n > 20:
            for k in range(5):                               # exact duplicates and near-duplicates
                for f in hits:
                    hits[f][k + 5] = hits[f][k]
            hits['score'][7] = hits['score'][2] + 1
            hits['nmatch'][11], hits['ncols'][11] = 7995, 10000      # 79.95 -> '80.0' (binary rounding decides)
            hits['nmatch'][12], hits['ncols'][12] = 1599, 2000       # 79.95 again with another denominator
            hits['end1'][11] = hits['start1'][11] + 99               # length1 = 100 exactly
            hits['end1'][12] = hits['start1'][12] + 98               # length1 = 99
        got = A.tab_blocks(hits, names, names, 100, 80)
        assert got == py_tab_blocks(hits, names, names, 100, 80), trial


def test_native_tab_formatter_rounds_the_identity_like_printf():
    """Every ratio nm/nc with nc < 200: the printed '%.1f' (integer fast path or C library at ties) must equal Python's."""
    from mimeo_b200 import align as A
    from tests.helpers import py_tab_blocks
    for nc in range(1, 200):
        nm = np.arange(0, nc + 1)
        n = len(nm)
        pos = np.arange(1, n + 1).astype(np.int32)
        hits = {'t_id': np.zeros(n, np.int32), 'q_id': np.zeros(n, np.int32), 'strand': np.zeros(n, np.int32), 'start1': pos,
                'end1': pos + 200, 'start2': np.ones(n, np.int32), 'end2': np.ones(n, np.int32), 'score': np.full(n, 3000, np.int32),
                'nmatch': nm.astype(np.int32), 'ncols': np.full(n, nc, np.int32)}
        for min_idt in (0, 80):
            assert A.tab_blocks(hits, ['a'], ['a'], 100, min_idt) == py_tab_blocks(hits, ['a'], ['a'], 100, min_idt), (nc, min_idt)


def _pandas_map_gff(tab_path, prefix, minLen, minIdt, ftype, chrlens):
    from mimeo_b200 import wrappers as W
    return ''.join(W.writeGFFlines(W.import_Align(infile=tab_path, prefix=prefix, minLen=minLen, minIdt=minIdt), chrlens, ftype))


def test_native_map_gff_equals_pandas_path_and_goldens(tmp_path):
    """mb2_map_gff (one native pass) against import_Align + writeGFFlines (the pandas pair that mirrors the reference) and the
    reference-generated goldens: string-order sort of coordinates, stability on full ties, UID width, filter edge (end-start)."""
    from mimeo_b200 import wrappers as W
    rng = np.random.default_rng(21)
    names = ['chr1', 'chr10', 'chr2', 'Scaf_9', 's']
    lines = ['#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity']
    for _ in range(5000):
        s = int(rng.integers(1, 20000))
        lines.append('%s\t+\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%.1f' % (names[int(rng.integers(0, 5))], s, s + int(rng.integers(90, 400)),
                                                                  names[int(rng.integers(0, 5))], '+-'[int(rng.integers(0, 2))],
                                                                  int(rng.integers(1, 9999)), int(rng.integers(1, 9999)),
                                                                  int(rng.integers(3000, 99999)), rng.uniform(85, 100)))
    lines += [lines[5], lines[5], lines[17]]                    # full ties: stability decides
    lines.append('  # indented comment lines are skipped too')
    tab = tmp_path / 'm.tab'
    tab.write_text('\n'.join(lines) + '\n')
    chrlens = [('chr1', '20000'), ('chr2', '30000')]
    for prefix, minLen, minIdt in (('BHit', 100, 90), (None, 100, 95), ('X', 0, 0), ('p', 250, 99)):
        n, body = W.map_gff_text(str(tab), prefix, minLen, minIdt, 'BHit')
        want = _pandas_map_gff(str(tab), prefix, minLen, minIdt, 'BHit', None)
        header = '##gff-version 3\n##seqid\tsource\ttype\tstart\tend\tscore\tstrand\tphase\tattributes\n'
        assert header + body == want and n == body.count('\n') > 0
    out = tmp_path / 'o.gff3'
    W.write_map_gff(str(tab), str(out), chrlens, 'BHit', 100, 90, 'BHit')
    assert out.read_text() == _pandas_map_gff(str(tab), 'BHit', 100, 90, 'BHit', chrlens)
    # reference-generated goldens (made by running the reference's own import_Align + writeGFFlines)
    import json
    m = json.loads(read_golden('manifest.json'))['map']
    g = tmp_path / 'gold_map.tab'
    g.write_text(read_golden('map.tab'))
    o = tmp_path / 'gold_map.gff3'
    W.write_map_gff(str(g), str(o), [tuple(x) for x in m['chrlens']], m['prefix'], m['minLen'], m['minIdt'], m['ftype'])
    assert o.read_text() == read_golden('map.gff3')
    g7 = tmp_path / 'gold_kat7.tab'
    g7.write_text(read_golden('map_kat7.tab'))
    W.write_map_gff(str(g7), str(o), None, None, 100, 95, 'HGT')
    assert o.read_text() == read_golden('map_kat7.gff3')


def test_native_map_gff_empty_and_malformed(tmp_path):
    from mimeo_b200 import wrappers as W
    p = tmp_path / 'e.tab'
    p.write_text('#h\nchr1\t+\t5\t50\tq\t+\t1\t2\t3000\t99.0\n')
    assert W.map_gff_text(str(p), 'BHit', 100, 90)[0] == 0
    with pytest.raises(SystemExit):
        W.write_map_gff(str(p), None, None, 'BHit', 100, 90)
    p.write_text('#h\nchr1\t+\t5\n')
    with pytest.raises(RuntimeError, match='line 2'):
        W.map_gff_text(str(p), 'BHit', 100, 90)


def test_native_split_fasta_bytes_equal_the_plain_writer(tmp_path):
    """mb2_fasta_split against the plain numpy statement of SeqIO.write's layout (fasta.write_fasta_record), record by
    record: awkward lengths (0, 1, 59, 60, 61, multiples of 60), headers with descriptions, CRLF and blank-padded input."""
    import ctypes as C
    from mimeo_b200 import _lib, fasta, utils
    rng = np.random.default_rng(5)
    lens = [0, 1, 59, 60, 61, 120, 121, 4097, 250_003]
    recs = []
    text = []
    for k, n in enumerate(lens):
        seq = rng.choice(np.frombuffer(b'ACGTNacgt', dtype=np.uint8), n)
        hdr = f'rec{k} some description {k}' if k % 2 else f'rec{k}'
        recs.append((f'rec{k}', hdr, seq))
        w = [70, 60, 13][k % 3]
        body = '\n'.join(seq[i:i + w].tobytes().decode() for i in range(0, n, w))
        eol = '\r\n' if k == 3 else '\n'
        text.append('>' + hdr + eol + body.replace('\n', eol) + (eol if n else ''))
    fa = tmp_path / 'g.fa'
    fa.write_bytes(''.join(text).encode())
    out, ref = tmp_path / 'native', tmp_path / 'plain'
    out.mkdir(); ref.mkdir()
    utils.splitFasta(str(fa), str(out))
    for rid, hdr, seq in recs:
        fasta.write_fasta_record(str(ref / (rid + '.fa')), hdr, seq)
    assert sorted(os.listdir(out)) == sorted(os.listdir(ref))
    for fn in os.listdir(ref):
        assert (out / fn).read_bytes() == (ref / fn).read_bytes(), fn
    # not unique + unique=False: the last record of an id wins; unique=True: the files before the repeat exist, then exit 1
    dup = tmp_path / 'dup.fa'
    dup.write_text('>a first\nAC\n>b\nGG\n>a second\nTTT\n>c\nA\n')
    d1, d2 = tmp_path / 'd1', tmp_path / 'd2'
    d1.mkdir(); d2.mkdir()
    utils.splitFasta(str(dup), str(d1), unique=False)
    assert sorted(os.listdir(d1)) == ['a.fa', 'b.fa', 'c.fa'] and (d1 / 'a.fa').read_text() == '>a second\nTTT\n'
    with pytest.raises(SystemExit) as e:
        utils.splitFasta(str(dup), str(d2))
    assert e.value.code == 1 and sorted(os.listdir(d2)) == ['a.fa', 'b.fa'] and (d2 / 'a.fa').read_text() == '>a first\nAC\n'
    n = C.c_uint64(0)
    assert _lib.lib().mb2_fasta_split(os.fsencode(str(dup)), os.fsencode(str(tmp_path / 'missing_dir')), 0, 60, 2, C.byref(n)) != 0


def test_native_gff_rows_equal_the_plain_statement():
    """mb2_format_gff against the f-string statement of the reference's closing awk (wrappers.py:1166-1173): names with
    odd bytes, ids beyond 99999 (the %05d field widens), prefix None as the reference would print it, thread seams."""
    from mimeo_b200 import engine
    rng = np.random.default_rng(11)
    names = ['S3', 's10', 's2', 'chr_with.dots|and|bars', 'x' * 70]
    for n, first in [(0, 1), (1, 1), (7, 1), (20_011, 1), (5000, 99_990)]:
        c = np.sort(rng.integers(0, len(names), n)).astype(np.int32)
        s = rng.integers(0, 2_000_000_000, n).astype(np.int32)
        e = (s.astype(np.int64) + rng.integers(0, 100_000, n)).clip(0, 2**31 - 1).astype(np.int32)
        for prefix in ('Self_Repeat', None):
            want = ''.join(f'{names[int(c[k])]}\tmimeo-self\tSelf_Repeat_intra\t{int(s[k])}\t{int(e[k])}\t.\t+\t.\tID={prefix}_{first + k:05d}\n'
                           for k in range(n))
            got = engine.segment_gff_text(c, s, e, names, 'mimeo-self', 'Self_Repeat_intra', prefix, first)
            assert got == want
    with pytest.raises(Exception):
        engine.segment_gff_text(np.array([5], np.int32), np.array([1], np.int32), np.array([2], np.int32), names, 'a', 'b', 'c')
