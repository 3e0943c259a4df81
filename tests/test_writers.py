"""Text writers (tracy_b200/writers.py) against outputs of the reference's own plotAlignment / writeDecomposition
(tests/golden/make_golden_writers.py), byte for byte."""
import os

import numpy as np

from conftest import ROOT
from tracy_b200 import writers


def test_plot_alignment_and_decomposition_match_reference():
    G = np.load(os.path.join(ROOT, "tests", "golden", "writers_golden.npz"))
    for i in range(int(G["n"])):
        pos, rl, fw, score, key, ll = (int(x) for x in G[f"cfg{i}"])
        got = writers.plot_alignment(bytes(G[f"r0_{i}"]), bytes(G[f"r1_{i}"]), bytes(G[f"chr{i}"]), pos, rl, bool(fw), score, key, tuple(G[f"a1a2_{i}"]), ll)
        assert got.encode("latin-1") == bytes(G[f"txt{i}"]), i
    assert writers.write_decomposition(G["decomp"]).encode() == bytes(G["decomp_txt"])


def test_align_fasta_layout():
    # reference src/sage.h:326-339: ">" stem, row 0, ">" chr " (forward)" | " (reverse)", row 1
    assert writers.align_fasta("trace1", b"AC-GT", b"ACTGT", b"chr2", True) == ">trace1\nAC-GT\n>chr2 (forward)\nACTGT\n"
    assert writers.align_fasta("t", b"A", b"A", "x", False) == ">t\nA\n>x (reverse)\nA\n"


def _trace_writer_cases(seed, n, max_samples):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_trace_writers", os.path.join(ROOT, "tests", "golden", "make_golden_trace_writers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.cases(seed, n, max_samples)


def _mine(c):
    txt = writers.trace_txt(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], c["tl"], c["tr"]).encode("latin-1")
    js = writers.trace_json(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"]).encode("latin-1")
    aj = b""
    if c["wellformed"]:
        aj = writers.trace_align_json(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"], c["row0"], c["row1"], b"chr7", c["pos"], c["fwd"]).encode("latin-1")
    return txt, js, aj


def test_trace_writers_match_reference_goldens():
    """P.abif (traceTxtOut), the basecall JSON (traceJsonOut) and P.json of `tracy align` (alignmentTracePadding +
    traceAlignJsonOut) byte for byte against outputs of the reference (tests/golden/make_golden_trace_writers.py)."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "trace_writers_golden.npz"))
    _, cs = _trace_writer_cases(77, int(G["n"]), 160)
    for i, c in enumerate(cs):
        txt, js, aj = _mine(c)
        assert txt == bytes(G[f"txt{i}"]), i
        assert js == bytes(G[f"json{i}"]), i
        assert aj == bytes(G[f"ajson{i}"]), i


def test_trace_writers_differential(oracle_ref):
    """More seeded cases against the reference build itself, where it exists."""
    import pytest
    if oracle_ref is None:
        pytest.skip("reference build not present")
    mod, cs = _trace_writer_cases(5, 150, 500)
    for i, c in enumerate(cs):
        assert _mine(c) == mod.reference_outputs(oracle_ref, c), i


def _decompose_json_cases(seed, n):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_decompose_json", os.path.join(ROOT, "tests", "golden", "make_golden_decompose_json.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.cases(seed, n)


def _my_decompose_json(c):
    from tracy_b200 import variants
    var = []
    for r0, r1, ch, p in c["als"]:
        variants.call_variants(r0, r1, ch, p, var)
    var = writers.sort_variants(var)
    keys = [(v["chr"], v["pos"], v["basenum"]) for v in var]
    unique = len(set(keys)) == len(keys)          # std::sort leaves the order of records with equal keys open
    doc = writers.decompose_json(c["cfg"], c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], var, c["allele1"], c["allele2"], c["align3"], c["decomp"],
                                 c["indelshift"], c["breakpoint"], c["a1a2"]).encode("latin-1")
    return doc, unique


def test_decompose_json_matches_reference_goldens():
    """P.json of `tracy decompose` (traceAlleleAlignJsonOut over callVariants + sort) byte for byte against documents written by
    the reference (tests/golden/make_golden_decompose_json.py); every document also parses as JSON."""
    import json
    G = np.load(os.path.join(ROOT, "tests", "golden", "decompose_json_golden.npz"))
    _, cs = _decompose_json_cases(21, int(G["n"]))
    for i, c in enumerate(cs):
        doc, unique = _my_decompose_json(c)
        if unique:
            assert doc == bytes(G[f"json{i}"]), i
        parsed = json.loads(doc)
        assert parsed["meta"]["program"] == "tracy" and len(parsed["variants"]["rows"]) == len(parsed["variants"]["xranges"])


def test_decompose_json_differential(oracle_ref):
    import pytest
    if oracle_ref is None:
        pytest.skip("reference build not present")
    mod, cs = _decompose_json_cases(8, 80)
    compared = 0
    for i, c in enumerate(cs):
        doc, unique = _my_decompose_json(c)
        if unique:
            assert doc == mod.reference_json(oracle_ref, c), i
            compared += 1
    assert compared > 40


def _assemble_writer_cases(seed, n):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_assemble_writers", os.path.join(ROOT, "tests", "golden", "make_golden_assemble_writers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.cases(seed, n)


def _my_assemble_parts(c):
    from tracy_b200 import trim
    g = trim.reverse_complement_trace(c["acgt"], c["bcpos"], c["qual"], c["pri"], c["sec"], c["sec"])
    return dict(acgt_sum=[int(x) for x in (g["acgt"].astype(np.int64) * np.arange(1, g["acgt"].shape[1] + 1)).sum(axis=1)], bcpos=[int(x) for x in g["bcpos"]],
                qual=[int(x) for x in g["qual"]], primary=g["primary"], secondary=g["secondary"],
                byrow=writers.aligned_trace_by_row(c["rows"], c["row"], c["name"], c["fwd"], c["isref"]))


def test_assemble_output_functions_match_reference_goldens():
    """alignedTraceByRow and reverseComplementTrace -- the two reference functions inside assemble()'s output section."""
    import json
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "assemble_writers_golden.json")))
    _, cs = _assemble_writer_cases(41, len(want))
    for i, c in enumerate(cs):
        assert _my_assemble_parts(c) == want[i], i


def test_assemble_output_functions_differential(oracle_ref):
    import pytest
    if oracle_ref is None:
        pytest.skip("reference build not present")
    mod, cs = _assemble_writer_cases(3, 200)
    for i, c in enumerate(cs):
        assert _my_assemble_parts(c) == mod.reference_outputs(oracle_ref, c), i


def test_assemble_files_layout():
    """The files of `tracy assemble` (reference src/assemble.h:473-579) put together from the pinned parts: FASTA records per row with
    the orientation, a JSON document that parses and holds one msa entry and one gapped trace per row, one `.vertical` line per column,
    the consensus as FASTA or FASTQ."""
    import json
    from tracy_b200 import trim
    rng = np.random.default_rng(12)
    rows = np.frombuffer(b"--ACGT-AC" b"TTAC-TGAC" b"-TACGT---", np.uint8).reshape(3, 9)
    names, fwd = ["t1", "t2", "t3"], [True, False, True]
    padded = []
    for i in range(3):
        seq = bytes(rows[i]).replace(b"-", b"")
        n = len(seq)
        acgt = rng.integers(0, 900, (4, n * 10 + 8)).astype(np.int32)
        bcpos = (np.arange(n) * 10 + 4).astype(np.int32)
        tr = dict(acgt=acgt, bcpos=bcpos, qual=np.full(n, 30, np.uint8), primary=seq.decode(), secondary=seq.decode(), consensus=seq.decode())
        if not fwd[i]:
            tr = trim.reverse_complement_trace(acgt, bcpos, tr["qual"], seq, seq, seq)
            tr["primary"] = tr["secondary"] = tr["consensus"] = seq.decode()           # keep the row's characters for the padding walk
        padded.append(writers.alignment_trace_padding(bytes(rows[i]), tr["acgt"], tr["bcpos"], tr["qual"], tr["primary"], tr["secondary"], tr["consensus"]))
    out = writers.assemble_files(names, fwd, rows, b"TTACGTGAC", b"TTACGTGAC", b"IIIIIIIII", padded, include_consensus=True, fmt="fastq")
    assert out[".align.fa"] == ">t1 (forward)\n--ACGT-AC\n>t2 (reverse)\nTTAC-TGAC\n>t3 (forward)\n-TACGT---\n>Consensus\nTTACGTGAC\n"
    assert out[".vertical"].splitlines()[2] == "AAA|A" and len(out[".vertical"].splitlines()) == 9
    assert out[".cons.fq"] == "@Consensus\nTTACGTGAC\n+\nIIIIIIIII\n"
    doc = json.loads(out[".json"])
    assert [m["traceFileName"] for m in doc["msa"]] == names and doc["msa"][0]["leadingGaps"] == "2" and doc["msa"][2]["align"] == "TACGT"
    assert len(doc["gappedTraces"]) == 3 and doc["gappedTraces"][0]["leadingGaps"] == 2 and doc["gappedTraces"][1]["basecalls"]
    assert ".cons.fa" in writers.assemble_files(names, fwd, rows, b"TTACGTGAC", b"TTACGTGAC", b"IIIIIIIII", padded)
    # reference-guided layout (src/assemble.h:284-376): traces in rank order = rows bottom-up, the reference in the last row
    rg = writers.assemble_files(names[:2], fwd[:2], rows, b"TTACGTGAC", b"TTACGTGAC", b"IIIIIIIII", [padded[1], padded[0]], reference_last=True)
    assert rg[".align.fa"] == ">t1 (forward)\nTTAC-TGAC\n>t2 (reverse)\n--ACGT-AC\n>Reference\n-TACGT---\n"
    doc = json.loads(rg[".json"])
    assert [m["reference"] for m in doc["msa"]] == [False, False, True] and doc["msa"][2]["traceFileName"] == "" and len(doc["gappedTraces"]) == 2


def test_per_subcommand_file_sets():
    """`align -o P` writes P.abif, P.align.fa, P.txt, P.json; `decompose -o P` writes P.abif, P.decomp, P.align1-3,
    P.json (P.bcf needs htslib). The composed sets hold exactly the texts of the single writers."""
    import json
    from tracy_b200 import variants
    rng = np.random.default_rng(3)
    n = 30
    acgt = rng.integers(0, 2000, (4, n * 12 + 30)).astype(np.int32)
    bcpos = (np.arange(n) * 12 + 9).astype(np.int32)
    qual = rng.integers(0, 61, n).astype(np.uint8)
    pri = bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8))
    row0, row1 = pri[:10] + b"--" + pri[10:], b"AC" + pri[2:10] + b"GT" + pri[10:]
    files = writers.align_files("t1", acgt, bcpos, qual, pri, pri, pri, 2, 3, row0, row1, b"chr5", 1000, 40, True, 77)
    assert sorted(files) == [".abif", ".align.fa", ".json", ".txt"]
    assert files[".json"] == writers.trace_align_json(acgt, bcpos, qual, pri, pri, pri, row0, row1, b"chr5", 1000, True)
    assert files[".abif"].count("\n") == acgt.shape[1] + 1 and json.loads(files[".json"])["refpos"] == 1001
    var = variants.call_variants(row0[2:], row1[2:], b"chr5", 1000, [])
    a1 = (row0[2:], row1[2:], b"chr5", 1000, True, 50)
    cfg = dict(trim_left=2, trim_right=3, qual_cut=45, pratio=0.33, input="x/t1.ab1", genome="g.fa")
    dfiles = writers.decompose_files(cfg, acgt, bcpos, qual, pri, pri, pri, [(-1, 5), (0, 9), (1, 4)], var, a1, a1, (row0[2:], row1[2:], 12), (40, 40), False, 4, (0.6, 0.4))
    assert sorted(dfiles) == [".abif", ".align1", ".align2", ".align3", ".decomp", ".json"]
    assert dfiles[".align3"].startswith(">Alt1 (Estimated allelic Fraction: 0.6)") and ">Alt2 (Estimated allelic Fraction: 0.4)" in dfiles[".align3"]
    assert json.loads(dfiles[".json"])["decomposition"]["x"] == [-1, 0, 1]


def _fastx_cases(seed, n_cases):
    rng = np.random.default_rng(seed)
    for it in range(n_cases):
        n = int(rng.integers(1, 80))
        ns = n * 10 + int(rng.integers(2, 30))
        bcpos = np.sort(rng.choice(ns, n, replace=False)).astype(np.int32)
        if it % 9 == 0 and n > 2:
            bcpos[n // 2] = bcpos[n // 2 - 1]
        qual = rng.integers(0, 61, n).astype(np.uint8)
        pri, sec, con = (bytes(rng.choice(list(a), n).astype(np.uint8)) for a in (b"ACGT", b"ACGTRY", b"ACGTN"))
        tl = int(rng.integers(0, n // 2 + 1))
        tr = int(rng.integers(0, n - tl)) if n - tl > 0 else 0
        yield ["primary", "secondary", "consensus", "other"][it % 4], tl, tr, ns, bcpos, qual, pri, sec, con


def test_trace_fasta_fastq(oracle_ref):
    """The fasta / fastq formats of the basecall subcommand (traceFastaOut / traceFastqOut, src/fasta.h:98-158): fixed cases written by
    the reference, and a differential run where its build exists."""
    q = np.array([0, 10, 40, 60, 20], np.uint8)
    pos = np.array([3, 9, 15, 21, 27], np.int32)
    assert writers.trace_fasta("primary", 1, 1, b"ACGTA", b"ACRTA", b"ACNTA") == ">primary\nCGT\n"
    assert writers.trace_fasta("consensus", 0, 0, b"ACGTA", b"ACRTA", b"ACNTA") == ">consensus\nACNTA\n"
    assert writers.trace_fasta("nothing", 0, 0, b"ACGTA", b"ACRTA", b"ACNTA") == ""
    assert writers.trace_fastq("secondary", 1, 1, 30, pos, q, b"ACGTA", b"ACRTA", b"ACNTA") == "@secondary\nCRT\n+\n+I]\n"
    assert writers.trace_fastq("nothing", 0, 2, 30, pos, q, b"ACGTA", b"ACRTA", b"ACNTA") == "+\n!+I\n"
    if oracle_ref is not None:
        assert oracle_ref.trace_fastx(1, "secondary", 1, 1, 30, pos, q, b"ACGTA", b"ACRTA", b"ACNTA") == b"@secondary\nCRT\n+\n+I]\n"
        assert oracle_ref.trace_fastx(1, "nothing", 0, 2, 30, pos, q, b"ACGTA", b"ACRTA", b"ACNTA") == b"+\n!+I\n"
        for i, (ot, tl, tr, ns, bcpos, qual, pri, sec, con) in enumerate(_fastx_cases(8, 300)):
            assert writers.trace_fasta(ot, tl, tr, pri, sec, con).encode("latin-1") == oracle_ref.trace_fastx(0, ot, tl, tr, ns, bcpos, qual, pri, sec, con), i
            assert writers.trace_fastq(ot, tl, tr, ns, bcpos, qual, pri, sec, con).encode("latin-1") == oracle_ref.trace_fastx(1, ot, tl, tr, ns, bcpos, qual, pri, sec, con), i


def test_native_writers_equal_python_writers(tmp_path):
    """csrc/writers.cu (tb_write_*) against the Python writers, which are pinned on the reference: random traces and alignments incl.
    leading / trailing / inner gap runs, a single basecall, basecall positions that never come up, all four plotAlignment keys."""
    import ctypes as C

    from tracy_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(77)
    for it in range(40):
        nbc = 1 if it == 0 else int(rng.integers(2, 120))
        ns = 12 * nbc + int(rng.integers(5, 40))
        acgt = rng.integers(-5, 3000, size=(4, ns)).astype(np.int32)
        bcpos = np.sort(rng.choice(np.arange(2, ns - 1), size=nbc, replace=False)).astype(np.int32)
        if it % 7 == 3 and nbc > 4:
            bcpos[2] = bcpos[1]                                    # a position that is never met by the forward walk
        qual = rng.integers(0, 61, nbc).astype(np.uint8)
        pri = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))
        sec = bytes(rng.choice(list(b"ACGTRYKMSWN"), nbc).astype(np.uint8))
        con = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))
        # an alignment of the nbc trace bases against a reference: row0 carries exactly nbc non-gap characters
        row0, row1, k = bytearray(), bytearray(), 0
        for _ in range(int(rng.integers(0, 4))):
            row0 += b"-"; row1 += bytes([b"ACGT"[int(rng.integers(0, 4))]])
        while k < nbc:
            u = rng.random()
            if u < 0.08:
                g = int(rng.integers(1, 5)); row0 += b"-" * g; row1 += bytes(rng.choice(list(b"ACGT"), g).astype(np.uint8))
            elif u < 0.14:
                row0 += bytes([pri[k]]); row1 += b"-"; k += 1
            else:
                row0 += bytes([pri[k]]); row1 += bytes([pri[k] if rng.random() < 0.9 else b"ACGT"[int(rng.integers(0, 4))]]); k += 1
        for _ in range(int(rng.integers(0, 4))):
            row0 += b"-"; row1 += b"A"
        row0, row1 = bytes(row0), bytes(row1)
        tl, trr = int(rng.integers(0, nbc + 3)), int(rng.integers(0, nbc + 3))
        pos, fwd, score, ll = int(rng.integers(0, 10 ** 6)), bool(it % 2), int(rng.integers(-500, 3000)), [60, 80, 7][it % 3]
        reflen = len(row1) - row1.count(b"-")
        prefix = str(tmp_path / f"n{it}")
        writers.write_align_files(prefix, "trace%d" % it, acgt, bcpos, qual, pri, sec, con, tl, trr, row0, row1, "chr%d" % it, pos, reflen, fwd, score, ll)
        want = writers.align_files("trace%d" % it, acgt, bcpos, qual, pri, sec, con, tl, trr, row0, row1, b"chr%d" % it, pos, reflen, fwd, score, ll)
        for sfx, text in want.items():
            with open(prefix + sfx, "rb") as fh:
                assert fh.read() == text.encode("latin-1"), (it, sfx)
        for key in (1, 2, 3):
            p = str(tmp_path / f"n{it}.k{key}")
            assert L.tb_write_plot_alignment(p.encode(), row0, row1, len(row0), b"chrX", pos, reflen, int(fwd), score, key, 0.37, 0.6300000001, ll) == 0
            with open(p, "rb") as fh:
                assert fh.read() == writers.plot_alignment(row0, row1, b"chrX", pos, reflen, fwd, score, key, (0.37, 0.6300000001), ll).encode("latin-1"), (it, key)
    # invalid arguments are refused, nothing is written
    assert L.tb_write_trace_txt(str(tmp_path / "x").encode(), None, 0, 0) == 1


def test_native_decompose_writers_equal_python_writers(tmp_path):
    """tb_write_decompose_json / tb_write_decomposition / tb_write_plot_alignment (keys 1-3) against writers.decompose_files (pinned on the
    reference's traceAlleleAlignJsonOut, writeDecomposition, plotAlignment) with an empty variant list."""
    rng = np.random.default_rng(78)
    for it in range(25):
        nbc = int(rng.integers(2, 150))
        ns = 12 * nbc + int(rng.integers(5, 40))
        acgt = rng.integers(0, 3000, size=(4, ns)).astype(np.int32)
        bcpos = np.sort(rng.choice(np.arange(2, ns - 1), size=nbc, replace=False)).astype(np.int32)
        qual = rng.integers(0, 61, nbc).astype(np.uint8)
        pri = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))
        sec = bytes(rng.choice(list(b"ACGTRYKMSWNB"), nbc).astype(np.uint8))
        con = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))

        def rows(n):
            a = bytes(rng.choice(list(b"ACGT-"), n, p=[.23, .23, .23, .23, .08]).astype(np.uint8))
            b = bytes(rng.choice(list(b"ACGT-"), n, p=[.23, .23, .23, .23, .08]).astype(np.uint8))
            return a, b
        r10, r11 = rows(int(rng.integers(1, 400)))
        r20, r21 = rows(int(rng.integers(1, 400)))
        r30, r31 = rows(int(rng.integers(1, 300)))
        tl, trr = int(rng.integers(0, max(nbc // 3, 1))), int(rng.integers(0, max(nbc // 3, 1)))
        bp = int(rng.integers(0, max(nbc - tl, 1)))
        decomp = [(int(-i), int(rng.integers(0, 300))) for i in range(int(rng.integers(0, 30)), -1, -1)] + [(i, int(rng.integers(0, 300))) for i in range(1, int(rng.integers(1, 30)))]
        cfg = dict(trim_left=tl, trim_right=trr, qual_cut=45, pratio=[0.33, 0.5, 0.1][it % 3], input="/some/dir/trace%d.ab1" % it, genome="ref%d.fa" % it)
        a1 = (r10, r11, b"chr%d" % it, int(rng.integers(0, 10 ** 6)), bool(it % 2), int(rng.integers(-100, 3000)))
        a2 = (r20, r21, b"chr%d" % it, int(rng.integers(0, 10 ** 6)), bool(it % 2), int(rng.integers(-100, 3000)))
        a3 = (r30, r31, int(rng.integers(-100, 3000)))
        fr = (float(rng.integers(0, 101)) / 100, float(rng.integers(0, 101)) / 100 + 1e-9 * it)
        lens = (len(r11) - r11.count(b"-"), len(r21) - r21.count(b"-"))
        prefix = str(tmp_path / f"d{it}")
        writers.write_decompose_files(prefix, cfg, acgt, bcpos, qual, pri, sec, con, decomp, a1, a2, a3, lens, bool(it % 3), bp, fr, [60, 75][it % 2])
        want = writers.decompose_files(cfg, acgt, bcpos, qual, pri, sec, con, decomp, [], a1, a2, a3, lens, bool(it % 3), bp, fr, [60, 75][it % 2])
        for sfx in (".decomp", ".align1", ".align2", ".align3", ".json"):
            with open(prefix + sfx, "rb") as fh:
                assert fh.read() == want[sfx].encode("latin-1"), (it, sfx)


def test_native_assemble_writer_equals_python_writers(tmp_path):
    """tb_write_assemble_files against writers.assemble_files over trim.trim_basecalls / trim.reverse_complement_trace /
    writers.alignment_trace_padding (each pinned on the reference): random alignments with leading / trailing / inner gaps, trims, flipped
    traces with IUPAC calls, reference-guided layout, both consensus formats and none."""
    from tracy_b200 import trim
    rng = np.random.default_rng(79)
    for it in range(12):
        nt = int(rng.integers(1, 7))
        ncol = int(rng.integers(30, 260))
        ref_last = it % 3 == 2
        nrow = nt + (1 if ref_last else 0)
        rows = np.full((nrow, ncol), ord("-"), np.uint8)
        traces, fwd, names = [], [], []
        for i in range(nt):
            nbc_kept = int(rng.integers(1, ncol - 5))
            start = int(rng.integers(0, ncol - nbc_kept))
            cols = np.sort(rng.choice(np.arange(start, ncol), size=nbc_kept, replace=False)) if rng.random() < 0.5 else np.arange(start, start + nbc_kept)
            tl, trr = int(rng.integers(0, 6)), int(rng.integers(0, 6))
            nbc = nbc_kept + tl + trr
            ns = 12 * nbc + int(rng.integers(5, 30))
            acgt = rng.integers(0, 2500, size=(4, ns)).astype(np.int32)
            bcpos = np.sort(rng.choice(np.arange(2, ns - 1), size=nbc, replace=False)).astype(np.int32)
            qual = rng.integers(0, 61, nbc).astype(np.uint8)
            pri = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))
            sec = bytes(rng.choice(list(b"ACGTRYKMSWHVDBN"), nbc).astype(np.uint8))
            con = bytes(rng.choice(list(b"ACGTN"), nbc).astype(np.uint8))
            f = bool(rng.random() < 0.5)
            kept = pri[tl: nbc - trr]
            shown = kept if f else bytes(trim._COMPLEMENT.get(chr(c), chr(c)).encode()[0] for c in kept[::-1])
            rows[i, cols] = np.frombuffer(shown, np.uint8)
            traces.append(dict(acgt=acgt, bcpos=bcpos, qual=qual, primary=pri, secondary=sec, consensus=con, tl=tl, tr=trr))
            fwd.append(f); names.append("trace_%d" % i)
        if ref_last:
            rows[nt] = rng.choice(list(b"ACGT-"), ncol).astype(np.uint8)
        row_of = [nt - 1 - i for i in range(nt)] if ref_last else list(range(nt))
        if ref_last:                                              # trace i sits in row nt-1-i: put its characters there
            rows[:nt] = rows[:nt][::-1].copy()
        gapped = bytes(rng.choice(list(b"ACGT-"), ncol).astype(np.uint8))
        cs = gapped.replace(b"-", b"")
        qs = bytes(rng.integers(33, 100, len(cs)).astype(np.uint8))
        fmt = ["fasta", "fastq", "other"][it % 3]
        padded = []
        for i, t in enumerate(traces):
            nb = trim.trim_basecalls(t["acgt"].shape[1], t["bcpos"], t["qual"], t["primary"], t["secondary"], t["consensus"], t["tl"], t["tr"])
            a = t["acgt"]
            if not fwd[i]:
                rv = trim.reverse_complement_trace(a, nb["bcpos"], nb["qual"], nb["primary"], nb["secondary"], nb["consensus"])
                a, nb = rv["acgt"], rv
            padded.append(writers.alignment_trace_padding(bytes(rows[row_of[i]]), a, nb["bcpos"], nb["qual"], nb["primary"], nb["secondary"], nb["consensus"]))
        want = writers.assemble_files(names, fwd, rows, gapped, cs, qs, padded, include_consensus=bool(it % 2), fmt=fmt, reference_last=ref_last)
        prefix = str(tmp_path / f"a{it}")
        writers.write_assemble_files(prefix, names, fwd, rows, row_of, gapped, cs, qs, traces, include_consensus=bool(it % 2), fmt=fmt, reference_last=ref_last)
        for sfx in (".align.fa", ".json", ".vertical", ".cons.fa", ".cons.fq"):
            assert os.path.exists(prefix + sfx) == (sfx in want), (it, sfx)
            if sfx in want:
                with open(prefix + sfx, "rb") as fh:
                    got = fh.read()
                assert got == want[sfx].encode("latin-1"), (it, sfx)
