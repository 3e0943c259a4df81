"""Generates tests/golden/glue_golden.npz from the REFERENCE ITSELF (oracle/_ref/libtracy_ref.so = unmodified tracy
headers behind oracle/ref_bridge.cpp) for the rows either side of the DP (a3, a5, a12, a15, a16, a17):
createProfile, reverseComplementProfile, trimReferenceSlice, findBreakpoint, _createProfile(char MSA),
revSeqBasedOnDist, msa + consensus. Run in the build container (needs /root/reference):

    python tests/golden/make_golden_glue.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tracy_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SC = (3, -5, -10, -4)
CODES = b"ACGTRYSWKMN"


def trace_case(rng, nbc, style):
    """A synthetic Trace/BaseCalls: int32[4][ns] samples with peaks every ~12 samples, basecall positions on the peaks."""
    ns = 12 * nbc + 40
    tr = rng.integers(0, 25, size=(4, ns)).astype(np.int32)
    pos = (12 * np.arange(nbc) + 10 + rng.integers(-2, 3, nbc)).astype(np.int32)
    pri, sec = bytearray(), bytearray()
    for j in range(nbc):
        b = int(rng.integers(0, 4))
        tr[b, pos[j]] += int(rng.integers(300, 1500))
        p, s = b"ACGT"[b], b"ACGT"[b]
        u = rng.random()
        if u < 0.2:                                  # heterozygous position: second peak, IUPAC secondary
            b2 = (b + int(rng.integers(1, 4))) % 4
            tr[b2, pos[j]] += int(rng.integers(200, 900))
            s = CODES[int(rng.integers(0, len(CODES)))] if style == 1 else b"ACGT"[b2]
        elif u < 0.25:
            p = ord("N")
        elif u < 0.3 and style == 2:
            tr[:, pos[j]] = 0                         # totalsig == 0 -> flat 0.25 column
        elif u < 0.35 and style == 2:
            tr[:, pos[j]] = -tr[:, pos[j]]            # negative samples (int16 traces can carry them)
        pri.append(p); sec.append(s)
    return tr, pos, bytes(pri), bytes(sec)


def assembly_case(rng, ntr, contig_len, tlen, noise):
    contig = synth.random_seq(rng, contig_len)
    step = max((contig_len - tlen) // max(ntr - 1, 1), 1)
    profs, truth = [], []
    for t in range(ntr):
        st = min(t * step + int(rng.integers(0, 5)), contig_len - tlen)
        s = synth.mutate_seq(rng, contig[st: st + tlen], 0.01, 0.004)
        p = synth.profile_from_seq(rng, s, noise)
        rc = bool(rng.integers(0, 2)) and t > 0
        if rc:
            p = np.ascontiguousarray(p[[3, 2, 1, 0, 4, 5], ::-1])
        profs.append(p); truth.append(not rc)
    return profs, truth


def main():
    ref = loader.ref()
    assert ref is not None, "oracle/_ref/libtracy_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(20261018)
    d = {}
    # createProfile + reverseComplementProfile + findBreakpoint
    ncp = 18
    d["ncp"] = np.int64(ncp)
    for i in range(ncp):
        nbc = [1, 2, 7, 40, 133, 300][i % 6]
        tr, pos, pri, sec = trace_case(rng, nbc, i % 3)
        tl, trr = [(0, 0), (3, 5), (50, 50), (nbc, 0)][i % 4]
        p = ref.create_profile(tr, pos, pri, sec, tl, trr)
        d[f"cp_tr{i}"], d[f"cp_pos{i}"] = tr, pos
        d[f"cp_pri{i}"], d[f"cp_sec{i}"] = np.frombuffer(pri, np.uint8), np.frombuffer(sec, np.uint8)
        d[f"cp_trim{i}"] = np.array([tl, trr], np.int32)
        d[f"cp_out{i}"] = p
        d[f"cp_rc{i}"] = ref.revcomp_profile(p)
        d[f"cp_bp{i}"] = np.array(ref.find_breakpoint(p), np.float64)
    nbp = 10
    d["nbp"] = np.int64(nbp)
    for i in range(nbp):                              # profiles with a real signal-quality break
        m = int(rng.integers(60, 400))
        p = synth.random_profile(rng, m, "trace")
        cut = int(rng.integers(26, m - 26))
        q = synth.random_profile(rng, m - cut, "ties") * np.float32(0.5 + 0.05 * i)
        if i % 2: p[:, cut:] = q
        else: p[:, : m - cut] = q
        d[f"bp_p{i}"] = p
        d[f"bp_out{i}"] = np.array(ref.find_breakpoint(p), np.float64)
    # trimReferenceSlice on real alignments
    ntr = 16
    d["ntrim"] = np.int64(ntr)
    for i in range(ntr):
        nref = int(rng.integers(200, 500)); m = int(rng.integers(40, 180))
        refseq = synth.random_seq(rng, nref)
        st = int(rng.integers(0, nref - m))
        tseq = synth.mutate_seq(rng, refseq[st: st + m], 0.03, 0.03)
        score, r0, r1 = ref.gotoh(tseq, refseq, 1, 0, SC)
        fw, pos, tl, trr = bool(i % 2), int(rng.integers(0, 10 ** 6)), [0, 10, 50, 300][i % 4], [0, 7, 50, 300][(i // 2) % 4]
        out, npos = ref.trim_reference_slice(r0, r1, refseq, fw, pos, tl, trr)
        d[f"tr_r0{i}"], d[f"tr_r1{i}"], d[f"tr_ref{i}"] = (np.frombuffer(x, np.uint8) for x in (r0, r1, refseq))
        d[f"tr_cfg{i}"] = np.array([fw, pos, tl, trr, npos], np.int64)
        d[f"tr_out{i}"] = np.frombuffer(out, np.uint8)
    # _createProfile(char MSA)
    nal = 8
    d["nal"] = np.int64(nal)
    for i in range(nal):
        nrow, ncol = int(rng.integers(1, 9)), int(rng.integers(1, 80))
        rows = rng.choice(np.frombuffer(b"ACGTNacgtn-----RX", np.uint8), size=(nrow, ncol))
        if i % 3 == 0: rows[0, :] = 0x2D
        d[f"al_rows{i}"] = rows
        d[f"al_out{i}"] = ref.profile_from_alignment(rows)
    # revSeqBasedOnDist + msa + consensus
    nms = 5
    d["nmsa"] = np.int64(nms)
    for i in range(nms):
        ntr_, clen, tlen, noise = [(4, 260, 120, 0.3), (7, 420, 150, 0.35), (10, 500, 140, 0.25), (6, 300, 100, 0.45), (12, 640, 160, 0.3)][i]
        profs, truth = assembly_case(rng, ntr_, clen, tlen, noise)
        fwd = ref.rev_seq_based_on_dist(profs, [True] * ntr_, SC)
        oriented = [p if f else ref.revcomp_profile(p) for p, f in zip(profs, fwd)]
        r = ref.msa(oriented, SC, 0.5)
        d[f"ms_n{i}"] = np.int64(ntr_)
        for k, p in enumerate(profs):
            d[f"ms_p{i}_{k}"] = p
        d[f"ms_fwd{i}"] = np.array(fwd, np.uint8)
        d[f"ms_rows{i}"], d[f"ms_idx{i}"], d[f"ms_dist{i}"] = r["rows"], r["seqidx"].astype(np.int64), r["dist"].astype(np.int64)
        d[f"ms_gapped{i}"], d[f"ms_cons{i}"], d[f"ms_qual{i}"] = (np.frombuffer(r[k], np.uint8) for k in ("gapped", "cons", "qual"))
        print(f"msa case {i}: n={ntr_} fwd={fwd} (planted {truth}) ncol={r['rows'].shape[1]} cons={len(r['cons'])}")
    np.savez_compressed(os.path.join(OUT, "glue_golden.npz"), **d)
    print("wrote glue_golden.npz", os.path.getsize(os.path.join(OUT, "glue_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
