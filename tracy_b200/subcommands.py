"""Files in -> files out: whole tracy subcommands for MANY invocations at once, as a host pipeline around the batched GPU calls.

  align      `tracy align -r ref trace`                 reference src/sage.h:58-357       one job per trace file
  consensus  `tracy consensus trace1 trace2`            reference src/consensus.h:332-590 one job per trace pair
  assemble   `tracy assemble [-r ref] traces...`        reference src/assemble.h:57-600   one job per trace set

A call takes a list of jobs (what N command lines of the reference would name) and returns one exit code per job -- the value
the reference's entry point returns for that command line -- having written the same files under each job's output prefix.
No option parsing: the options are keyword arguments named like the reference's config fields.

Pipeline: jobs go through in chunks. Reading the input files of chunk k+1 (reader threads) and formatting / writing the output
files of chunk k-1 (writer threads) overlap with the GPU stages of chunk k, which all run on the calling thread (a Context
serves one host thread at a time): trace decode (tb_trace_unpack), basecall, createProfile, the score-only orientation batch,
the alignment batches. The GPU calls release the GIL, so the writers' string work runs underneath them.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import consensus as cons_mod
from . import drivers, msa, trim, writers
from .api import AlignConfig, DnaScore

PP, PS = "pp", "ps"
MAX_SINGLE_FASTA_SIZE = 50000                        # src/fasta.h:10-12
_FIX_NAME = "\\,'\"()[]{}<>:\t\r#"                   # _fixReferenceName, src/fasta.h:16-35
_DEGENERATE_TO_N = bytes.maketrans(b"WSMKRYBDHV", b"N" * 10)        # _replaceDegenerateBases, src/fasta.h:38-53


def _stem(path):
    f = os.path.basename(path)
    k = f.rfind(".")
    return f if k <= 0 else f[:k]


def _read(path):
    try:
        with open(path, "rb") as fh:
            return fh.read()
    except OSError:
        return None


def load_single_fasta(data):
    """loadSingleFasta (src/fasta.h:55-95) on the bytes of a file. Returns (name, sequence) or None where the reference returns
    false (a second '>' record, characters outside the IUPAC nucleotides)."""
    name, seq = "", []
    for line in data.split(b"\n"):
        if not line:
            continue
        if line[:1] == b">":
            if name:
                return None
            name = (line[1:-1] if line.endswith(b"\r") else line[1:]).decode("latin-1")
        else:
            seq.append((line[:-1] if line.endswith(b"\r") else line).upper())
    s = b"".join(seq).translate(_DEGENERATE_TO_N)
    if s.translate(None, b"ACGTN"):                                  # anything left is neither a nucleotide nor an IUPAC code
        return None
    return "".join(ch for ch in name if ch not in _FIX_NAME), s


def genome_type(data):
    """genomeType (src/fmindex.h:58-71) on the first bytes of the reference file: 0 gzipped FASTA (indexed genome), 2 trace file,
    1 single FASTA, -1 unknown."""
    if data[:2] == b"\x1f\x8b":
        return 0
    if data[:4] == b"ABIF" or data[:4] == b".scf":
        return 2
    if data[:1] == b">":
        return 1
    return -1


class _Writers:
    """Output side of the pipeline: {path: text} sets are encoded and written by worker threads while the caller goes on."""

    def __init__(self, workers):
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self.pending = []

    def submit(self, fn, *args):
        self.pending.append(self.pool.submit(fn, *args))

    @staticmethod
    def write(prefix, files):
        for suffix, text in files.items():
            with open(prefix + suffix, "wb") as fh:
                fh.write(text.encode("latin-1") if isinstance(text, str) else bytes(text))

    def drain(self):
        for f in self.pending:
            f.result()
        self.pending = []
        self.pool.shutdown()


def _chunks(n, size):
    return [range(s, min(s + size, n)) for s in range(0, n, size)]


def _load_traces(ctx, blobs, pratio):
    """readab / readscf + basecall for a list of file images (None = unreadable). Returns per file None (the reference prints an
    error and returns -1) or dict(acgt, ploc, bcpos, primary, secondary, consensus, qual (estimated)) -- one decode call and one
    basecall call for the whole list."""
    live = [i for i, b in enumerate(blobs) if b]
    out = [None] * len(blobs)
    if not live:
        return out
    tr = ctx.read_traces([blobs[i] for i in live])
    good = [(i, t) for i, t in zip(live, tr) if t["format"] >= 0 and t["ok"] and t["traceACGT"] is not None and len(t["basecallpos"])]
    if not good:
        return out
    bc = ctx.basecall([t["traceACGT"] for _, t in good], [t["basecallpos"] for _, t in good], pratio)
    for (i, t), b in zip(good, bc):
        out[i] = dict(acgt=t["traceACGT"], ploc=t["basecallpos"], bcpos=b["bcPos"], primary=b["primary"], secondary=b["secondary"],
                      consensus=b["consensus"], qual=trim.trace_quality(b["bcPos"], b["secondary"])[0])
    return out


def _trims(t, stringency, left, right):
    if stringency >= 1:
        return trim.trace_quality(t["bcpos"], t["secondary"], stringency)[1]
    return left, right


# ---- tracy align ---------------------------------------------------------------------------------------------------------------
def align(ctx, jobs, pratio=0.33, trim_stringency=0.0, trim_left=50, trim_right=50, linelimit=60, sc=DnaScore(3, -5, -10, -4), chunk=256, workers=4):
    """`tracy align -r <genome> -o <outprefix> <trace>` for every job = (trace path, genome path, outprefix). The genome is a single
    FASTA file (<= 50 kbp) or a wildtype trace file, as src/sage.h:196-305 tells them apart; indexed genomes go through
    drivers.align_genome_batch with an index built from the text (tb_index_build) and are not a file format of this function.
    Writes outprefix.abif / .align.fa / .txt / .json; returns the reference's exit codes (0, 1 missing file, -1 otherwise)."""
    semiglobal = AlignConfig(True, False)
    rc = [0] * len(jobs)
    wr = _Writers(workers)
    trim_stringency = min(float(trim_stringency), 9.0)                       # src/sage.h:122
    with ThreadPoolExecutor(max_workers=max(1, workers)) as readers, ThreadPoolExecutor(max_workers=1) as lookahead:
        parts = _chunks(len(jobs), chunk)
        ahead = None

        def fetch(part):
            paths = [jobs[i][0] for i in part] + [jobs[i][1] for i in part]
            return list(readers.map(_read, paths))

        for pi, part in enumerate(parts):
            blobs = ahead.result() if ahead is not None else fetch(part)
            ahead = lookahead.submit(fetch, parts[pi + 1]) if pi + 1 < len(parts) else None
            k = len(part)
            tblob, gblob = blobs[:k], blobs[k:]
            for j, i in enumerate(part):                                     # reference first, then the trace: src/sage.h:125-134
                if not gblob[j] or not tblob[j]:
                    rc[i] = 1
            traces = _load_traces(ctx, [b if rc[i] == 0 else None for b, i in zip(tblob, part)], pratio)
            fasta, wild, meta = [], [], {}
            for j, i in enumerate(part):
                if rc[i]:
                    continue
                t = traces[j]
                if t is None:
                    rc[i] = -1
                    continue
                tl, trr = _trims(t, trim_stringency, trim_left, trim_right)
                if tl + trr >= len(t["bcpos"]):
                    rc[i] = -1
                    continue
                t["tl"], t["tr"] = tl, trr
                gt = genome_type(gblob[j])
                if gt == 1:
                    fa = load_single_fasta(gblob[j])
                    if fa is None or len(fa[1]) > MAX_SINGLE_FASTA_SIZE:
                        rc[i] = -1
                    else:
                        meta[j] = fa
                        fasta.append(j)
                elif gt == 2:
                    wild.append(j)
                else:
                    rc[i] = -1                                               # unknown format; gzipped genomes: see the docstring
                if rc[i] == -1:
                    wr.submit(_abif_only, jobs[i][2], t)     # traceTxtOut ran before the reference was looked at (src/sage.h:188)
            # wildtype traces: read + basecall as one more batch
            wt = _load_traces(ctx, [gblob[j] for j in wild], pratio)
            for j, g in zip(list(wild), wt):
                if g is None:
                    rc[part[j]] = -1
                    wild.remove(j)
                    wr.submit(_abif_only, jobs[part[j]][2], traces[j])
                else:
                    meta[j] = g
            live = fasta + wild
            if not live:
                continue
            lt = [traces[j] for j in live]
            full = ctx.create_profile([t["acgt"] for t in lt], [t["bcpos"] for t in lt], [t["primary"] for t in lt], [t["secondary"] for t in lt])
            trimmed = ctx.create_profile([t["acgt"] for t in lt], [t["bcpos"] for t in lt], [t["primary"] for t in lt], [t["secondary"] for t in lt],
                                         [t["tl"] for t in lt], [t["tr"] for t in lt])
            res = {}
            if fasta:
                nf = len(fasta)
                r = drivers.align_batch(ctx, trimmed[:nf], full[:nf], [meta[j][1] for j in fasta], sc, [traces[j]["tl"] for j in fasta],
                                        [traces[j]["tr"] for j in fasta])
                for j, x in zip(fasta, r):
                    x["chr"] = meta[j][0]
                    x["refslice_len"] = len(x["refslice"])
                    res[j] = x
            if wild:
                nf = len(fasta)
                gl = [meta[j] for j in wild]
                gfull = ctx.create_profile([g["acgt"] for g in gl], [g["bcpos"] for g in gl], [g["primary"] for g in gl], [g["secondary"] for g in gl])
                grev = ctx.revcomp_profile(gfull)
                tw, fw = trimmed[nf:], full[nf:]
                s = ctx.gotoh(PP, list(tw) + list(tw), list(gfull) + list(grev), sc, semiglobal, traceback=False)[0]      # src/sage.h:289-290
                nw = len(wild)
                fwd = [bool(s[q] > s[nw + q]) for q in range(nw)]
                refp = [gfull[q] if fwd[q] else grev[q] for q in range(nw)]
                score, ops, ol = ctx.gotoh(PP, fw, refp, sc, semiglobal)                                                   # src/sage.h:311
                for q, j in enumerate(wild):
                    row0, row1 = drivers.rows_from_ops(PP, fw[q], refp[q], bytes(ops[q, : ol[q]]))
                    pri = bytes(gl[q]["primary"])
                    res[j] = dict(forward=fwd[q], chr="wildtype", pos=0, score=int(score[q]), row0=row0, row1=row1,
                                  refslice_len=len(pri if fwd[q] else drivers.reverse_complement_seq(pri)))
            for j in live:
                wr.submit(_align_out, jobs[part[j]], traces[j], res[j], linelimit)
    wr.drain()
    return rc


def _abif_only(prefix, t):
    writers.write_trace_txt(prefix + ".abif", t["acgt"], t["bcpos"], t["qual"], t["primary"], t["secondary"], t["consensus"], t["tl"], t["tr"])


def _align_out(job, t, r, linelimit):
    # native writers (csrc/writers.cu, same bytes as writers.align_files): ~390 KB of text per trace, formatted outside the interpreter
    writers.write_align_files(job[2], _stem(job[0]), t["acgt"], t["bcpos"], t["qual"], t["primary"], t["secondary"], t["consensus"], t["tl"], t["tr"],
                              r["row0"], r["row1"], r["chr"], r["pos"], r["refslice_len"], r["forward"], r["score"], linelimit)


# ---- tracy consensus -----------------------------------------------------------------------------------------------------------
def consensus(ctx, jobs, label="Consensus", pratio=0.33, match_fraction=0.5, min_overlap=25, trim_stringency=0.0, trim_left1=50, trim_right1=50,
              trim_left2=50, trim_right2=50, linelimit=60, intersect=False, iupac=False, sc=DnaScore(3, -5, -10, -4), chunk=256, workers=4):
    """`tracy consensus -o <outprefix> <trace1> <trace2>` for every job = (trace1 path, trace2 path, outprefix). Writes
    outprefix_1st.abif, _2nd.abif, .align.fa, .fa, .fq and .txt; returns the reference's exit codes (0; 1 for a missing file or
    "No sufficient trace overlap!"; -1 for unreadable traces or trims that leave nothing)."""
    rc = [0] * len(jobs)
    wr = _Writers(workers)
    with ThreadPoolExecutor(max_workers=max(1, workers)) as readers, ThreadPoolExecutor(max_workers=1) as lookahead:
        parts = _chunks(len(jobs), chunk)
        ahead = None

        def fetch(part):
            return list(readers.map(_read, [jobs[i][0] for i in part] + [jobs[i][1] for i in part]))

        for pi, part in enumerate(parts):
            blobs = ahead.result() if ahead is not None else fetch(part)
            ahead = lookahead.submit(fetch, parts[pi + 1]) if pi + 1 < len(parts) else None
            k = len(part)
            for j, i in enumerate(part):
                if not blobs[j] or not blobs[k + j]:
                    rc[i] = 1
            tr = _load_traces(ctx, [b if rc[part[j % k]] == 0 else None for j, b in enumerate(blobs)], pratio)
            live = []
            for j, i in enumerate(part):
                if rc[i]:
                    continue
                a, b = tr[j], tr[k + j]
                if a is None or b is None:
                    rc[i] = -1
                    continue
                a["tl"], a["tr"] = _trims(a, trim_stringency, trim_left1, trim_right1)
                b["tl"], b["tr"] = _trims(b, trim_stringency, trim_left2, trim_right2)
                if a["tl"] + a["tr"] >= len(a["bcpos"]) or b["tl"] + b["tr"] >= len(b["bcpos"]):
                    rc[i] = -1
                    continue
                wr.submit(_cons_abif, jobs[i][2], a, b)                                    # src/consensus.h:489-490
                live.append(j)
            if not live:
                continue
            both = [tr[j] for j in live] + [tr[k + j] for j in live]
            prof = ctx.create_profile([t["acgt"] for t in both], [t["bcpos"] for t in both], [t["primary"] for t in both], [t["secondary"] for t in both],
                                      [t["tl"] for t in both], [t["tr"] for t in both])
            nl = len(live)
            res = drivers.consensus_batch(ctx, prof[:nl], prof[nl:], sc, min_overlap, match_fraction)
            rev = None
            for q, j in enumerate(live):
                i = part[j]
                if not res[q]["ok"]:
                    rc[i] = 1                                                              # src/consensus.h:546-549
                    continue
                p2 = prof[nl + q] if res[q]["forward"] else msa._revcomp(prof[nl + q])
                wr.submit(_cons_out, jobs[i], res[q], prof[q], p2, label, not intersect, iupac, linelimit)
    wr.drain()
    return rc


def _cons_abif(prefix, a, b):
    for t, tag in ((a, "_1st.abif"), (b, "_2nd.abif")):
        writers.write_trace_txt(prefix + tag, t["acgt"], t["bcpos"], t["qual"], t["primary"], t["secondary"], t["consensus"], t["tl"], t["tr"])


def _cons_out(job, r, p1, p2, label, union, iupac, linelimit):
    s1, s2 = _stem(job[0]), _stem(job[1])
    cs, qual = cons_mod.pairwise_consensus_native(r["row0"], r["row1"], p1, p2, union, iupac)
    _Writers.write(job[2], {".align.fa": cons_mod.consensus_align_fasta(s1, s2, r["row0"], r["row1"], r["forward"]),
                            ".fa": cons_mod.consensus_fasta(label, cs), ".fq": cons_mod.consensus_fastq(label, cs, qual),
                            ".txt": cons_mod.plot_clustal_pairwise(s1, s2, r["row0"], r["row1"], r["forward"], r["score"], linelimit)})


# ---- tracy assemble ------------------------------------------------------------------------------------------------------------
def assemble(ctx, jobs, pratio=0.33, trim_stringency=4.0, match_fraction=0.5, fraction_called=0.1, fmt="fasta", inc_cons=False, inc_ref=False,
             sc=DnaScore(3, -5, -10, -4), workers=4):
    """`tracy assemble [-r <reference.fa>] -o <outprefix> <trace>...` for every job = (list of trace paths, reference path or None,
    outprefix). Writes outprefix.align.fa / .json / .vertical / .cons.fa|.cons.fq; returns the reference's exit codes. The trace
    files of ALL jobs are decoded and basecalled in one GPU batch; every job then runs its own DP sequence (all-pairs orientation
    table, exclusion, progressive alignment -- drivers.assemble_denovo / assemble_reference)."""
    rc = [0] * len(jobs)
    wr = _Writers(workers)
    if trim_stringency != 0:
        trim_stringency = min(max(float(trim_stringency), 1.0), 9.0)                       # src/assemble.h:131-134
    match_fraction = min(max(float(match_fraction), 0.0), 1.0)
    flat = [(ji, p) for ji, job in enumerate(jobs) for p in job[0]]
    with ThreadPoolExecutor(max_workers=max(1, workers)) as readers:
        blobs = list(readers.map(_read, [p for _, p in flat]))
        refs = list(readers.map(lambda job: _read(job[1]) if job[1] else b"", jobs))
    for (ji, _), b in zip(flat, blobs):
        if not b:
            rc[ji] = 1                                                                     # "Trace file is missing", src/assemble.h:111-116
    traces = _load_traces(ctx, [b if rc[ji] == 0 else None for (ji, _), b in zip(flat, blobs)], pratio)
    per_job = {}
    for (ji, _), t in zip(flat, traces):
        per_job.setdefault(ji, []).append(t)
    # one createProfile batch over every usable trace of every job
    plan = []
    for ji, job in enumerate(jobs):
        if rc[ji]:
            continue
        ref = None
        if job[1]:
            ref = load_single_fasta(refs[ji]) if refs[ji] else ("", b"")                  # an unreadable file loads as an empty sequence
            if ref is None or len(ref[1]) > MAX_SINGLE_FASTA_SIZE:
                rc[ji] = -1
                continue
        ts = per_job.get(ji, [])
        for t in ts:
            if t is None:
                rc[ji] = -1
                break
            t["tl"], t["tr"] = (0, 0)
            if trim_stringency:
                t["tl"], t["tr"] = trim.trace_quality(t["bcpos"], t["secondary"], trim_stringency)[1]
                if t["tl"] + t["tr"] >= len(t["bcpos"]):
                    rc[ji] = -1                                                            # "Too stringent trimming parameters!"
                    break
        if rc[ji] == 0:
            plan.append((ji, ref, ts))
    allt = [t for _, _, ts in plan for t in ts]
    prof = ctx.create_profile([t["acgt"] for t in allt], [t["bcpos"] for t in allt], [t["primary"] for t in allt], [t["secondary"] for t in allt],
                              [t["tl"] for t in allt], [t["tr"] for t in allt]) if allt else []
    at = 0
    for ji, ref, ts in plan:
        p = prof[at: at + len(ts)]
        at += len(ts)
        names = [_stem(x) for x in jobs[ji][0]]
        if ref is not None:
            r = drivers.assemble_reference(ctx, p, ref[1], sc, match_fraction, fraction_called, inc_ref)
            if not r["idx"]:
                wr.submit(_assemble_empty, jobs[ji][2], fmt)
                continue
            order, fwd, rows = r["idx"], r["forward"], r["rows"]
            row_of = [len(order) - 1 - q for q in range(len(order))]
        else:
            r = drivers.assemble_denovo(ctx, p, sc, match_fraction, fraction_called)
            if r["rows"] is None:
                rc[ji] = -1                                                                # "At least 2 traces are required", src/assemble.h:460-463
                continue
            order = [r["kept"][q] for q in r["seqidx"]]
            fwd = [bool(r["forward"][r["kept"][q]]) for q in r["seqidx"]] if len(r["forward"]) == len(p) else [bool(r["forward"][q]) for q in r["seqidx"]]
            rows = r["rows"]
            row_of = list(range(len(order)))
        wr.submit(_assemble_out, jobs[ji][2], [names[q] for q in order], fwd, rows, row_of, r, [ts[q] for q in order], inc_cons, fmt, ref is not None)
    wr.drain()
    return rc


def _assemble_empty(prefix, fmt):
    """No trace matched the reference (src/assemble.h:236: the output block is skipped): the tail of assemble() still writes an empty
    .vertical and the consensus record."""
    _Writers.write(prefix, {".vertical": "", ".cons.fa" if fmt == "fasta" else ".cons.fq": (">Consensus\n\n" if fmt == "fasta" else "@Consensus\n\n+\n\n")}
                   if fmt in ("fasta", "fastq") else {".vertical": ""})


def _assemble_out(prefix, names, fwd, rows, row_of, r, ts, inc_cons, fmt, reference_last):
    # native (csrc/writers.cu): hard trim, reverse complement of flipped traces, padding along the alignment row and the four files -- the
    # same bytes as writers.assemble_files over trim.trim_basecalls / reverse_complement_trace / writers.alignment_trace_padding
    writers.write_assemble_files(prefix, names, fwd, rows, row_of, r["gapped"], r["consensus"], r["quality"], ts, inc_cons, fmt, reference_last)


# ---- tracy decompose -----------------------------------------------------------------------------------------------------------
def decompose(ctx, jobs, pratio=0.33, trim_stringency=0.0, trim_left=50, trim_right=50, maxindel=1000, madc=5, qual_cut=45, linelimit=60,
              sc=DnaScore(3, -5, -10, -4), chunk=256, workers=4):
    """`tracy decompose -r <reference.fa> -o <outprefix> <trace>` (reference src/indigo.h:42-455, without -v / -a) for every job =
    (trace path, single-FASTA reference path, outprefix). Writes outprefix.abif, .decomp, .align1, .align2, .align3 and .json; returns the
    reference's exit codes (0; 1 missing file; -1 unreadable trace, trims larger than the trace, reference that is no single FASTA of at
    most 50 kbp, alignment below the score threshold, no usable breakpoint). Every DP / sweep / fit stage of a chunk of jobs is one
    batched GPU call (drivers.decompose_batch); P.bcf (-v) needs htslib and is not written. Wildtype-trace and indexed-genome
    references are not file formats of this function (drivers.align_genome_batch anchors traces in a genome)."""
    rc = [0] * len(jobs)
    wr = _Writers(workers)
    trim_stringency = min(float(trim_stringency), 9.0)                       # src/indigo.h:106
    maxindel = max(int(maxindel), 1)                                         # :103
    with ThreadPoolExecutor(max_workers=max(1, workers)) as readers, ThreadPoolExecutor(max_workers=1) as lookahead:
        parts = _chunks(len(jobs), chunk)
        ahead = None

        def fetch(part):
            return list(readers.map(_read, [jobs[i][0] for i in part] + [jobs[i][1] for i in part]))

        for pi, part in enumerate(parts):
            blobs = ahead.result() if ahead is not None else fetch(part)
            ahead = lookahead.submit(fetch, parts[pi + 1]) if pi + 1 < len(parts) else None
            k = len(part)
            tblob, gblob = blobs[:k], blobs[k:]
            for j, i in enumerate(part):                                     # trace first, then the reference: src/indigo.h:124-134
                if not tblob[j] or not gblob[j]:
                    rc[i] = 1
            traces = _load_traces(ctx, [b if rc[i] == 0 else None for b, i in zip(tblob, part)], pratio)
            groups = {}
            for j, i in enumerate(part):
                if rc[i]:
                    continue
                t = traces[j]
                if t is None:
                    rc[i] = -1
                    continue
                t["tl"], t["tr"] = _trims(t, trim_stringency, trim_left, trim_right)
                if t["tl"] + t["tr"] >= len(t["bcpos"]):
                    rc[i] = -1
                    continue
                wr.submit(_abif_only, jobs[i][2], t)                         # traceTxtOut with the ORIGINAL basecalls, src/indigo.h:187
                fa = load_single_fasta(gblob[j]) if genome_type(gblob[j]) == 1 else None
                if fa is None or len(fa[1]) > MAX_SINGLE_FASTA_SIZE:
                    rc[i] = -1
                    continue
                t["fa"] = fa
                groups.setdefault((t["tl"], t["tr"]), []).append(j)
            for (tl, trr), idx in groups.items():
                ts = [traces[j] for j in idx]
                res = drivers.decompose_batch(ctx, [t["acgt"] for t in ts], [t["bcpos"] for t in ts], [t["primary"] for t in ts], [t["secondary"] for t in ts],
                                              [t["fa"][1] for t in ts], sc, tl, trr, maxindel, madc)
                for j, t, r in zip(idx, ts, res):
                    i = part[j]
                    if r is None:
                        rc[i] = -1                                           # "Alignment of trace to reference failed!" / no breakpoint
                        continue
                    wr.submit(_decompose_out, jobs[i], t, r, dict(trim_left=tl, trim_right=trr, qual_cut=qual_cut, pratio=pratio, input=jobs[i][0], genome=jobs[i][1]),
                              linelimit)
    wr.drain()
    return rc


def _decompose_out(job, t, r, cfg, linelimit):
    tl, trr = cfg["trim_left"], cfg["trim_right"]
    chr_name = t["fa"][0].encode("latin-1")
    bp = r["breakpoint"]
    breakpoint = bp["breakpoint"]
    if not bp["indelshift"]:                                                 # centre the viewport on the first SNP, src/indigo.h:391-395
        _, _, best = trim.trace_quality(t["bcpos"], r["secondary"], None, best=True)
        breakpoint = trim.nearest_snp(r["primary"], r["secondary"], tl, trr, best)
    a1, a2, a3 = r["align1"], r["align2"], r["align3"]
    allele1 = (a1["row0"], a1["row1"], chr_name, a1["pos"], r["forward"], a1["score"])
    allele2 = (a2["row0"], a2["row1"], chr_name, a2["pos"], r["forward"], a2["score"])
    # P.abif was written with the basecalls as they were before the decomposition; the rest through the native writers (same bytes as
    # writers.decompose_files with an empty variant list)
    writers.write_decompose_files(job[2], cfg, t["acgt"], t["bcpos"], t["qual"], r["primary"], r["secondary"], t["consensus"], r["decomp"], allele1, allele2,
                                  (a3["row0"], a3["row1"], a3["score"]), (len(a1["refslice"]), len(a2["refslice"])), bp["indelshift"], breakpoint,
                                  r["allele_fractions"], linelimit)
