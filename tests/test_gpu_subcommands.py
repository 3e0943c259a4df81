"""Files in -> files out on the GPU: tracy_b200.subcommands (decode, basecall, createProfile and every DP stage as batched CUDA
calls; readers / writers on host threads) against the reference's OWN subcommand entry points run on the same files behind
oracle/_ref -- `tracy align` (src/sage.h:58), `tracy consensus` (src/consensus.h:332), `tracy assemble` (src/assemble.h:57) --
`tracy decompose` (src/indigo.h:42) -- byte for byte over every output file, plus the exit codes. 29 + 22 + 11 + 27 command lines."""
import pytest

from tracy_b200 import subcommands

from subcmd_cases import (ALIGN_SUFFIXES, ASM_SUFFIXES, CONS_SUFFIXES, DEC_SUFFIXES, compare_dirs, make_align_jobs, make_assemble_jobs, make_consensus_jobs,
                          make_decompose_jobs)

pytestmark = pytest.mark.gpu


def _need_ref(oracle_ref):
    if oracle_ref is None:
        pytest.skip("oracle/_ref not built (it is built in the container that holds /root/reference and travels with the snapshot)")


def test_align_files_gpu(ctx, oracle_ref, tmp_path):
    _need_ref(oracle_ref)
    jobs, opts = make_align_jobs(str(tmp_path), n=24, seed=101)
    want = [oracle_ref.subcommand("align", ["-r", g, "-o", o + ".ref"] + extra + [t]) for (t, g, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.align(ctx, [jobs[i] for i in idx], chunk=3, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert want.count(0) >= 24
    assert compare_dirs([o for _, _, o in jobs], ALIGN_SUFFIXES) >= 4 * 24


def test_consensus_files_gpu(ctx, oracle_ref, tmp_path):
    _need_ref(oracle_ref)
    jobs, opts = make_consensus_jobs(str(tmp_path), n=20, seed=102)
    want = [oracle_ref.subcommand("consensus", ["-o", o + ".ref"] + extra + [a, b]) for (a, b, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.consensus(ctx, [jobs[i] for i in idx], chunk=4, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert want.count(0) >= 18 and 1 in want
    assert compare_dirs([o for _, _, o in jobs], CONS_SUFFIXES) >= 6 * 18


def test_assemble_files_gpu(ctx, oracle_ref, tmp_path):
    _need_ref(oracle_ref)
    jobs, opts = make_assemble_jobs(str(tmp_path), n=8, seed=103)
    want = [oracle_ref.subcommand("assemble", (["-r", r] if r else []) + ["-o", o + ".ref"] + extra + list(t)) for (t, r, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.assemble(ctx, [jobs[i] for i in idx], **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert want.count(0) >= 8 and -1 in want and 1 in want
    assert compare_dirs([o for _, _, o in jobs], ASM_SUFFIXES) >= 4 * 8


def test_decompose_files_gpu(ctx, oracle_ref, tmp_path):
    _need_ref(oracle_ref)
    jobs, opts = make_decompose_jobs(str(tmp_path), n=24, seed=104)
    want = [oracle_ref.subcommand("decompose", ["-r", g, "-o", o + ".ref"] + extra + [t]) for (t, g, o), extra in zip(jobs, opts["argv"])]
    for kw, idx in opts["groups"]:
        got = subcommands.decompose(ctx, [jobs[i] for i in idx], chunk=5, **kw)
        assert got == [want[i] for i in idx], (kw, got, [want[i] for i in idx])
    assert want.count(0) >= 20 and -1 in want and 1 in want
    assert compare_dirs([o for _, _, o in jobs], DEC_SUFFIXES) >= 6 * 20
