"""Files in -> files out: tracy_b200.subcommands against the reference's OWN subcommand entry points (`int sage(argc, argv)`,
`int consensus(argc, argv)`, `int assemble(argc, argv)`, `int indigo(argc, argv)`, src/sage.h:58 / src/consensus.h:332 /
src/assemble.h:57 / src/indigo.h:42, run unmodified behind oracle/_ref on the same input files),
byte for byte over every output file plus the exit codes. CPU suite: the pipeline's DP calls are served by the reference's
functions (tests/refctx.py), so what is checked here is the file plumbing, option handling, exit codes, pairwiseConsensus and
the writers; tests/test_gpu_subcommands.py runs the same comparison on the CUDA kernels."""
import os

import numpy as np
import pytest

from tracy_b200 import subcommands, synth

from subcmd_cases import (ALIGN_SUFFIXES, ASM_SUFFIXES, CONS_SUFFIXES, DEC_SUFFIXES, compare_dirs, make_align_jobs, make_assemble_jobs, make_consensus_jobs,
                          make_decompose_jobs)


@pytest.fixture(scope="module")
def refctx(oracle_ref):
    if oracle_ref is None:
        pytest.skip("oracle/_ref not built")
    from refctx import RefContext
    return RefContext(oracle_ref)


def test_align_files_vs_reference_main(refctx, oracle_ref, tmp_path):
    jobs, opts = make_align_jobs(str(tmp_path), n=6, seed=11)
    want = [oracle_ref.subcommand("align", ["-r", g, "-o", o + ".ref"] + extra + [t]) for (t, g, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.align(refctx, [jobs[i] for i in idx], chunk=2, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert 0 in want and -1 in want and 1 in want
    compare_dirs([o for _, _, o in jobs], ALIGN_SUFFIXES)


def test_consensus_files_vs_reference_main(refctx, oracle_ref, tmp_path):
    jobs, opts = make_consensus_jobs(str(tmp_path), n=6, seed=12)
    want = [oracle_ref.subcommand("consensus", ["-o", o + ".ref"] + extra + [a, b]) for (a, b, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.consensus(refctx, [jobs[i] for i in idx], chunk=2, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert 0 in want and 1 in want
    compare_dirs([o for _, _, o in jobs], CONS_SUFFIXES)


def test_assemble_files_vs_reference_main(refctx, oracle_ref, tmp_path):
    jobs, opts = make_assemble_jobs(str(tmp_path), n=3, seed=13)
    want = [oracle_ref.subcommand("assemble", (["-r", r] if r else []) + ["-o", o + ".ref"] + extra + list(t)) for (t, r, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.assemble(refctx, [jobs[i] for i in idx], **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert 0 in want and -1 in want and 1 in want
    assert compare_dirs([o for _, _, o in jobs], ASM_SUFFIXES) >= 12


def test_decompose_files_vs_reference_main(refctx, oracle_ref, tmp_path):
    jobs, opts = make_decompose_jobs(str(tmp_path), n=8, seed=14)
    want = [oracle_ref.subcommand("decompose", ["-r", g, "-o", o + ".ref"] + extra + [t]) for (t, g, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.decompose(refctx, [jobs[i] for i in idx], chunk=3, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert want.count(0) >= 6 and -1 in want and 1 in want
    assert compare_dirs([o for _, _, o in jobs], DEC_SUFFIXES) >= 6 * 6


def test_load_single_fasta_rules():
    assert subcommands.load_single_fasta(b">chr(1):x\r\nacgtn\r\nRYKM\n") == ("chr1x", b"ACGTNNNNN")
    assert subcommands.load_single_fasta(b">a\nACGT\n>b\nAC\n") is None
    assert subcommands.load_single_fasta(b">a\nAC-GT\n") is None
    assert subcommands.genome_type(b"\x1f\x8b\x08") == 0 and subcommands.genome_type(b"ABIF....") == 2 and subcommands.genome_type(b">x") == 1
