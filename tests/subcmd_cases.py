"""Synthetic command lines for the files-in -> files-out tests: ABIF / SCF trace files, FASTA and wildtype-trace references,
option variants, and the cases the reference answers with a non-zero exit code. Shared by the CPU and GPU suites."""
import os

import numpy as np

from tracy_b200 import synth

ALIGN_SUFFIXES = (".abif", ".align.fa", ".txt", ".json")
CONS_SUFFIXES = ("_1st.abif", "_2nd.abif", ".align.fa", ".fa", ".fq", ".txt")
ASM_SUFFIXES = (".align.fa", ".json", ".vertical", ".cons.fa", ".cons.fq")
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def sanger_file(rng, seq, het=0.03, scf=False):
    """A trace file whose peaks spell `seq` (a second, lower peak at a fraction `het` of the positions)."""
    nbc = len(seq)
    ns = 12 * nbc + 40
    tr = rng.integers(0, 25, size=(4, ns)).astype(np.int64)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int64)
    shape = np.array([0.15, 0.55, 1.0, 0.55, 0.15])
    for j in range(nbc):
        x = b"ACGT".index(seq[j])
        h = int(rng.integers(600, 1400))
        tr[x, pos[j] - 2: pos[j] + 3] += (h * shape).astype(np.int64)
        if rng.random() < het:
            y = int(rng.integers(0, 4))
            tr[y, pos[j] - 2: pos[j] + 3] += (h * rng.uniform(0.4, 0.8) * shape).astype(np.int64)
    if scf:
        return synth.scf_bytes([tr[k] for k in range(4)], pos)
    order = b"GATC"
    return synth.abif_bytes([tr[b"ACGT".index(c)] for c in order], order, pos, bytes(seq), rng.integers(5, 60, nbc))


def _write(path, data):
    with open(path, "wb") as fh:
        fh.write(data)
    return path


def make_align_jobs(root, n=24, seed=1):
    """n + 5 `tracy align` command lines: FASTA references (both strands, lower case, CRLF, IUPAC), wildtype-trace references, SCF
    traces, and the failures (missing file, reference that is no FASTA, > 50 kbp, multi-record FASTA, trims larger than the trace)."""
    rng = np.random.default_rng(seed)
    jobs, argv, kws = [], [], []
    for i in range(n):
        g = synth.random_seq(rng, int(rng.integers(1500, 3000)))
        st, L = int(rng.integers(100, 600)), int(rng.integers(450, 800))
        s = synth.mutate_seq(rng, g[st: st + L], 0.01, 0.004)
        if i % 3 == 1:
            s = s.translate(COMP)[::-1]
        t = _write(os.path.join(root, f"al{i}.{'scf' if i % 5 == 4 else 'ab1'}"), sanger_file(rng, s, scf=(i % 5 == 4)))
        if i % 4 == 3:
            gp = _write(os.path.join(root, f"al{i}_wt.ab1"), sanger_file(rng, g[max(st - 80, 0): st + L + 90], het=0.0))
        else:
            body = g.decode()
            if i % 4 == 1:
                body = body.lower()
            lines = [body[k: k + 70] for k in range(0, len(body), 70)]
            if i % 4 == 2:
                lines[3] = lines[3][:10] + "R" + lines[3][11:]
            eol = "\r\n" if i % 6 == 5 else "\n"
            gp = _write(os.path.join(root, f"al{i}.fa"), (">ref:%d (x)%s" % (i, eol) + eol.join(lines) + eol).encode())
        kw, av = {}, []
        if i % 3 == 2:
            kw, av = dict(trim_stringency=float(2 + i % 5)), ["-t", str(2 + i % 5)]
        elif i % 3 == 0:
            kw, av = dict(trim_left=20 + i, trim_right=35, linelimit=80), ["-q", str(20 + i), "-u", "35", "-l", "80"]
        jobs.append((t, gp, os.path.join(root, f"al{i}.out")))
        argv.append(av); kws.append(kw)
    t0, g0 = jobs[0][0], jobs[0][1]
    big = _write(os.path.join(root, "big.fa"), b">big\n" + synth.random_seq(rng, 50001) + b"\n")
    multi = _write(os.path.join(root, "multi.fa"), b">a\nACGTACGT\n>b\nACGT\n")
    junk = _write(os.path.join(root, "junk.txt"), b"hello world, not a reference\n")
    for k, (t, g, kw, av) in enumerate(((os.path.join(root, "missing.ab1"), g0, {}, []), (t0, big, {}, []), (t0, multi, {}, []), (t0, junk, {}, []),
                                        (t0, g0, dict(trim_left=400, trim_right=400), ["-q", "400", "-u", "400"]))):
        jobs.append((t, g, os.path.join(root, f"alx{k}.out")))
        argv.append(av); kws.append(kw)
    return jobs, dict(argv=argv, groups=_group(kws))


def make_consensus_jobs(root, n=20, seed=2):
    """n + 2 `tracy consensus` command lines: overlapping trace pairs (second trace on either strand), option variants
    (-i intersect, -a IUPAC, -t trimming, trims per trace, label), a pair without overlap (exit code 1) and a missing file."""
    rng = np.random.default_rng(seed)
    jobs, argv, kws = [], [], []
    for i in range(n):
        g = synth.random_seq(rng, 1600)
        s1 = g[100: 100 + int(rng.integers(500, 750))]
        st2 = int(rng.integers(300, 500))
        s2 = synth.mutate_seq(rng, g[st2: st2 + int(rng.integers(500, 800))], 0.01, 0.003)
        if i % 2:
            s2 = s2.translate(COMP)[::-1]
        if i == n - 1:
            s2 = synth.random_seq(rng, 600)                          # no overlap
        a = _write(os.path.join(root, f"c{i}_a.ab1"), sanger_file(rng, s1, het=0.06))
        b = _write(os.path.join(root, f"c{i}_b.{'scf' if i % 7 == 3 else 'ab1'}"), sanger_file(rng, s2, het=0.06, scf=(i % 7 == 3)))
        kw, av = {}, []
        if i % 4 == 1:
            kw, av = dict(intersect=True, iupac=True), ["-i", "-a"]
        elif i % 4 == 2:
            kw, av = dict(trim_stringency=3.0, label="S%d" % (i % 3), iupac=True), ["-t", "3", "-b", "S%d" % (i % 3), "-a"]
        elif i % 4 == 3:
            kw, av = dict(trim_left1=30, trim_right1=60, trim_left2=10, trim_right2=45, linelimit=50, min_overlap=40, match_fraction=0.6), \
                     ["-q", "30", "-u", "60", "-r", "10", "-s", "45", "-l", "50", "-c", "40", "-f", "0.6"]
        jobs.append((a, b, os.path.join(root, f"c{i}.out")))
        argv.append(av); kws.append(kw)
    jobs.append((jobs[0][0], os.path.join(root, "nothing.ab1"), os.path.join(root, "cx0.out")))
    argv.append([]); kws.append({})
    jobs.append((jobs[0][0], jobs[1][1], os.path.join(root, "cx1.out")))      # unrelated traces
    argv.append([]); kws.append({})
    return jobs, dict(argv=argv, groups=_group(kws))


def _group(kws):
    groups = {}
    for i, kw in enumerate(kws):
        groups.setdefault(tuple(sorted(kw.items())), []).append(i)
    return [(dict(k), idx) for k, idx in groups.items()]


def compare_dirs(prefixes, suffixes):
    """Every file the reference wrote under <prefix>.ref<suffix> must exist under <prefix><suffix> with the same bytes, and
    nothing may be written that the reference did not write."""
    checked = 0
    for p in prefixes:
        for sfx in suffixes:
            a, b = p + ".ref" + sfx, p + sfx
            assert os.path.exists(a) == os.path.exists(b), (b, "exists:", os.path.exists(b), "reference wrote it:", os.path.exists(a))
            if os.path.exists(a):
                with open(a, "rb") as fa, open(b, "rb") as fb:
                    x, y = fa.read(), fb.read()
                if x != y:
                    k = next((q for q in range(min(len(x), len(y))) if x[q] != y[q]), min(len(x), len(y)))
                    raise AssertionError((b, "differs at byte", k, x[max(k - 60, 0): k + 60], y[max(k - 60, 0): k + 60]))
                checked += 1
    return checked


def make_assemble_jobs(root, n=6, seed=3):
    """n + 3 `tracy assemble` command lines: trace sets tiling a contig (half of the traces on the reverse strand), de novo and
    reference-guided, option variants (-d, -a fastq, -i, -j, -t 0), a set with an unrelated trace (excluded), a set whose traces
    do not overlap (exit code -1 de novo), a missing file, a reference nothing matches."""
    rng = np.random.default_rng(seed)
    jobs, argv, kws = [], [], []
    for i in range(n):
        nt = int(rng.integers(3, 7))
        L, stepsz = 520, 230
        contig = synth.random_seq(rng, stepsz * nt + L + 60)
        traces = []
        for k in range(nt):
            s = synth.mutate_seq(rng, contig[30 + stepsz * k: 30 + stepsz * k + L + int(rng.integers(-40, 40))], 0.008, 0.003)
            if (k + i) % 2:
                s = s.translate(COMP)[::-1]
            traces.append(_write(os.path.join(root, f"as{i}_{k}.{'scf' if (i + k) % 6 == 5 else 'ab1'}"), sanger_file(rng, s, het=0.02, scf=((i + k) % 6 == 5))))
        if i == 2:
            traces.insert(1, _write(os.path.join(root, f"as{i}_alien.ab1"), sanger_file(rng, synth.random_seq(rng, 500))))
        refp = None
        if i % 2:
            refp = _write(os.path.join(root, f"as{i}.fa"), b">contig%d\n" % i + contig + b"\n")
        kw, av = {}, []
        if i % 3 == 1:
            kw, av = dict(fraction_called=0.5, fmt="fastq", inc_cons=True), ["-d", "0.5", "-a", "fastq", "-i"]
        elif i % 3 == 2:
            kw, av = dict(trim_stringency=0.0, inc_ref=bool(refp), match_fraction=0.4), ["-t", "0", "-f", "0.4"] + (["-j"] if refp else [])
        jobs.append((traces, refp, os.path.join(root, f"as{i}.out")))
        argv.append(av); kws.append(kw)
    far = [_write(os.path.join(root, f"asx_{k}.ab1"), sanger_file(rng, synth.random_seq(rng, 450))) for k in range(3)]
    jobs.append((far, None, os.path.join(root, "asx0.out")))
    argv.append([]); kws.append({})
    jobs.append(([jobs[0][0][0], os.path.join(root, "gone.ab1")], None, os.path.join(root, "asx1.out")))
    argv.append([]); kws.append({})
    other = _write(os.path.join(root, "asx_other.fa"), b">other\n" + synth.random_seq(rng, 1200) + b"\n")       # no trace matches: outputs of the tail only
    jobs.append((far, other, os.path.join(root, "asx2.out")))
    argv.append([]); kws.append({})
    return jobs, dict(argv=argv, groups=_group(kws))


DEC_SUFFIXES = (".abif", ".decomp", ".align1", ".align2", ".align3", ".json")


def het_file(rng, a1, a2, frac=0.6, scf=False):
    """A trace file with two alleles mixed frac : 1 - frac in signal space (the second allele shifted against the first behind an indel)."""
    nbc = min(len(a1), len(a2))
    ns = 12 * nbc + 40
    tr = rng.integers(0, 20, size=(4, ns)).astype(np.int64)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int64)
    shape = np.array([0.15, 0.55, 1.0, 0.55, 0.15])
    for j in range(nbc):
        h = int(rng.integers(700, 1300))
        tr[b"ACGT".index(a1[j]), pos[j] - 2: pos[j] + 3] += (h * frac * shape).astype(np.int64)
        tr[b"ACGT".index(a2[j]), pos[j] - 2: pos[j] + 3] += (h * (1 - frac) * shape).astype(np.int64)
    if scf:
        return synth.scf_bytes([tr[k] for k in range(4)], pos)
    order = b"GATC"
    return synth.abif_bytes([tr[b"ACGT".index(c)] for c in order], order, pos, bytes(a1[:nbc]), rng.integers(5, 60, nbc))


def make_decompose_jobs(root, n=20, seed=4):
    """n + 3 `tracy decompose` command lines: heterozygous insertions and deletions of 1-25 bp on either strand, homozygous traces (no
    shift: the homozygous breakpoint search and the nearest-SNP viewport), SNVs, option variants (-i, -t, -q/-u, -l, -c), SCF files,
    an unrelated trace (exit code -1 with P.abif left behind), a reference that is no FASTA, a missing file."""
    rng = np.random.default_rng(seed)
    jobs, argv, kws = [], [], []
    for i in range(n):
        g = synth.random_seq(rng, int(rng.integers(1400, 2400)))
        st, L, bp, ln = int(rng.integers(100, 400)), int(rng.integers(560, 760)), int(rng.integers(180, 380)), int(rng.integers(1, 26))
        a1 = bytearray(g[st: st + L])
        if i % 4 == 3:
            a2 = bytearray(a1)                                                 # homozygous
        elif i % 2:
            a2 = bytearray((g[st: st + bp] + g[st + bp + ln:])[:L])             # deletion on the second allele
        else:
            a2 = bytearray((g[st: st + bp] + synth.random_seq(rng, ln) + g[st + bp:])[:L])
        for _ in range(int(rng.integers(0, 4))):
            a2[int(rng.integers(0, len(a2)))] = b"ACGT"[int(rng.integers(0, 4))]
        a1, a2 = bytes(a1), bytes(a2)
        if i % 3 == 1:
            a1, a2 = a1.translate(COMP)[::-1], a2.translate(COMP)[::-1]
        if i == n - 1:
            a1 = a2 = synth.random_seq(rng, 600)                               # matches nothing: alignment below the threshold
        t = _write(os.path.join(root, f"d{i}.{'scf' if i % 6 == 4 else 'ab1'}"), het_file(rng, a1, a2, [0.6, 0.5, 0.7][i % 3], scf=(i % 6 == 4)))
        gp = _write(os.path.join(root, f"d{i}.fa"), b">locus%d\n" % i + g + b"\n")
        kw, av = dict(maxindel=30), ["-i", "30"]
        if i % 5 == 1:
            kw, av = dict(maxindel=40, trim_stringency=4.0, linelimit=70), ["-i", "40", "-t", "4", "-l", "70"]
        elif i % 5 == 2:
            kw, av = dict(maxindel=30, trim_left=30, trim_right=70, madc=4), ["-i", "30", "-q", "30", "-u", "70", "-c", "4"]
        jobs.append((t, gp, os.path.join(root, f"d{i}.out")))
        argv.append(av); kws.append(kw)
    junk = _write(os.path.join(root, "djunk.txt"), b"not a reference\n")
    for k, (t, g_, kw, av) in enumerate(((jobs[0][0], junk, dict(maxindel=30), ["-i", "30"]), (os.path.join(root, "dgone.ab1"), jobs[0][1], dict(maxindel=30), ["-i", "30"]),
                                         (jobs[0][0], jobs[0][1], dict(maxindel=30, trim_left=350, trim_right=350), ["-i", "30", "-q", "350", "-u", "350"]))):
        jobs.append((t, g_, os.path.join(root, f"dx{k}.out")))
        argv.append(av); kws.append(kw)
    return jobs, dict(argv=argv, groups=_group(kws))
