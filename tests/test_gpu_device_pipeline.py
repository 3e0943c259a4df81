"""File bytes -> samples -> basecalls -> profiles -> anchoring -> DP with every intermediate in HBM (TB_MEM_DEVICE on every
entry point): trace unpack, basecall, createProfile, anchoring, and tb_gotoh_ps reading its reference windows straight out of
the anchoring index' device copy of the genome. torch is only the allocator. Checked against the host-buffer calls of the
same entry points (which are parity-tested against the reference elsewhere) and against the C oracle's DP."""
import ctypes as C

import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, capi, synth

pytestmark = pytest.mark.gpu


def _traces(rng, genome, n):
    files, truth = [], []
    for i in range(n):
        L = int(rng.integers(300, 520))
        p = int(rng.integers(0, len(genome) - L))
        seq = genome[p:p + L]
        ns = 12 * L + 40
        ch = rng.integers(0, 20, size=(4, ns))
        pos = 12 * np.arange(L) + 10
        for j, c in enumerate(seq):
            ch[b"ACGT".index(c), pos[j] - 1: pos[j] + 2] += np.array([350, 1000, 350])
        files.append(synth.abif_bytes([ch[2], ch[0], ch[3], ch[1]], b"GATC", pos, seq, np.full(L, 40)))
        truth.append((p, L))
    return files, truth


def test_device_resident_pipeline(ctx, oracle_port):
    import torch
    dev = torch.device("cuda", 0)
    lib, h = ctx._lib, ctx._h
    rng = np.random.default_rng(17)
    genome = synth.random_seq(rng, 60000)
    text = genome + b"\n"
    N = 24
    files, truth = _traces(rng, genome, N)
    host = ctx.read_traces(files)                                      # host-buffer reference for the unpack step

    def P(t):
        return C.c_void_p(t.data_ptr())

    def NP(a):
        return a.ctypes.data_as(C.c_void_p)

    # ---- unpack: file bytes (host) -> samples / positions in HBM
    flen = np.array([len(f) for f in files], np.int64)
    foff = np.concatenate([[0], np.cumsum(flen)[:-1]]).astype(np.int64)
    blob = np.frombuffer(b"".join(files), np.uint8)
    info = (capi.TraceInfo * N)()
    assert lib.tb_trace_scan(NP(blob), NP(foff), NP(flen), N, C.cast(info, C.c_void_p)) == 0
    ns = np.array([info[i].nsamples for i in range(N)], np.int64)
    nb = np.array([info[i].nbasecalls for i in range(N)], np.int64)
    soff = np.concatenate([[0], np.cumsum(4 * ns)[:-1]]).astype(np.int64)
    boff = np.concatenate([[0], np.cumsum(nb)[:-1]]).astype(np.int64)
    d_samples = torch.zeros(int(4 * ns.sum()), dtype=torch.int32, device=dev)
    tot = int(nb.sum())
    d_ploc = torch.zeros(tot, dtype=torch.int32, device=dev)
    d_qual, d_b1, d_b2 = (torch.zeros(tot, dtype=torch.uint8, device=dev) for _ in range(3))
    ctx._check(lib.tb_trace_unpack(h, NP(blob), NP(foff), NP(flen), N, capi.TB_MEM_DEVICE, P(d_samples), NP(soff), P(d_ploc), P(d_qual), P(d_b1), P(d_b2), NP(boff)))
    smp = d_samples.cpu().numpy()
    for i in range(N):
        assert np.array_equal(smp[soff[i]: soff[i] + 4 * ns[i]].reshape(4, -1), host[i]["traceACGT"])

    # ---- basecall on the device arenas
    d_soff, d_slen = torch.from_numpy(soff).to(dev), torch.from_numpy(ns.astype(np.int32)).to(dev)
    d_boff, d_blen = torch.from_numpy(boff).to(dev), torch.from_numpy(nb.astype(np.int32)).to(dev)
    d_pos = torch.zeros(tot, dtype=torch.int32, device=dev)
    d_pri, d_sec, d_con = (torch.zeros(tot, dtype=torch.uint8, device=dev) for _ in range(3))
    d_olen = torch.zeros(N, dtype=torch.int32, device=dev)
    bb = capi.BasecallBatch(capi.Arena(P(d_samples), P(d_soff), P(d_slen)), capi.Arena(P(d_ploc), P(d_boff), P(d_blen)), N, capi.TB_MEM_DEVICE)
    ctx._check(lib.tb_basecall(h, C.byref(bb), C.c_float(0.33), P(d_pos), P(d_pri), P(d_sec), P(d_con), P(d_boff), P(d_olen)))
    want_bc = ctx.basecall([x["traceACGT"] for x in host], [x["basecallpos"] for x in host], 0.33)
    olen = d_olen.cpu().numpy()
    con = d_con.cpu().numpy()
    for i in range(N):
        assert olen[i] == len(want_bc[i]["primary"])
        assert con[boff[i]: boff[i] + olen[i]].tobytes() == want_bc[i]["consensus"]

    # ---- createProfile (trimmed) on the device arenas
    tl, tr = 30, 30
    poff = (6 * boff).astype(np.int64)
    d_poff = torch.from_numpy(poff).to(dev)
    d_prof = torch.zeros(int(6 * tot), dtype=torch.float32, device=dev)
    d_plen = torch.zeros(N, dtype=torch.int32, device=dev)
    d_tl = torch.full((N,), tl, dtype=torch.int32, device=dev)
    d_tr = torch.full((N,), tr, dtype=torch.int32, device=dev)
    pb = capi.ProfileBatch(capi.Arena(P(d_samples), P(d_soff), P(d_slen)), capi.Arena(P(d_pos), P(d_boff), P(d_olen)), P(d_pri), P(d_sec), P(d_tl), P(d_tr), N,
                           capi.TB_MEM_DEVICE)
    ctx._check(lib.tb_create_profile(h, C.byref(pb), P(d_prof), P(d_poff), P(d_plen)))
    plen = d_plen.cpu().numpy()
    want_prof = ctx.create_profile([x["traceACGT"] for x in host], [b["bcPos"] for b in want_bc], [b["primary"] for b in want_bc],
                                   [b["secondary"] for b in want_bc], tl, tr)
    prof_host = d_prof.cpu().numpy()
    for i in range(N):
        assert np.array_equal(prof_host[poff[i]: poff[i] + 6 * plen[i]].reshape(6, -1), want_prof[i])

    # ---- anchoring on the device consensus arena
    idx = ctx.build_index(text)
    d_anch, d_fwd = (torch.zeros(N, dtype=torch.uint8, device=dev) for _ in range(2))
    d_sup = torch.zeros(N, dtype=torch.int32, device=dev)
    d_best = torch.zeros(N, dtype=torch.int64, device=dev)
    arena = capi.Arena(P(d_con), P(d_boff), P(d_olen))
    res = capi.AnchorResult(P(d_anch), P(d_fwd), P(d_sup), P(d_best), None)
    ctx._check(lib.tb_anchor(h, idx._h, C.byref(arena), N, capi.TB_MEM_DEVICE, capi.AnchorConfig(tl, tr, 15, 3), C.byref(res)))
    want_a = ctx.anchor(idx, [b["consensus"] for b in want_bc], tl, tr, 15, 3)
    assert np.array_equal(d_anch.cpu().numpy().astype(bool), want_a["anchored"]) and want_a["anchored"].all()
    best = d_best.cpu().numpy()
    assert np.array_equal(best, want_a["bestpos"]) and np.array_equal(best, [t[0] for t in truth])

    # ---- DP: the windows are slices of the index' own device copy of the genome (no second upload of reference bytes)
    s0 = np.maximum(best - 200, 0)
    wlen = (np.minimum(best + nb + 200, len(genome)) - s0).astype(np.int32)
    d_woff, d_wlen = torch.from_numpy(s0.astype(np.int64)).to(dev), torch.from_numpy(wlen).to(dev)
    d_scores = torch.zeros(N, dtype=torch.int32, device=dev)
    stride = int((plen.max() + wlen.max() + 15) // 16 * 16)
    d_ops = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
    d_ol = torch.zeros(N, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.gotoh_device("ps", d_prof.data_ptr(), d_poff.data_ptr(), d_plen.data_ptr(), idx.device_text, d_woff.data_ptr(), d_wlen.data_ptr(), N,
                     d_scores.data_ptr(), d_ops.data_ptr(), stride, d_ol.data_ptr(), DnaScore(3, -5, -10, -4), AlignConfig(True, False))
    sc, ops, ol = d_scores.cpu().numpy(), d_ops.cpu().numpy(), d_ol.cpu().numpy()
    for i in range(N):
        ws, wops = oracle_port.gotoh_ps(want_prof[i], genome[s0[i]: s0[i] + wlen[i]], 1, 0, (3, -5, -10, -4))
        assert int(sc[i]) == ws and bytes(ops[i, : ol[i]]) == wops, i
    idx.close()


def test_allelic_fraction_device_mode(ctx):
    import os
    import torch
    from conftest import ROOT
    G = np.load(os.path.join(ROOT, "tests", "golden", "fraction_golden.npz"))
    dev = torch.device("cuda", 0)
    idx = [i for i in range(int(G["n"])) if tuple(int(x) for x in G[f"cfg{i}"]) == (50, 50)]
    tr = [G[f"tr{i}"] for i in idx]
    tlen = np.array([t.shape[1] for t in tr], np.int32)
    toff = np.concatenate([[0], np.cumsum(4 * tlen.astype(np.int64))[:-1]]).astype(np.int64)
    blen = np.array([len(G[f"pos{i}"]) for i in idx], np.int32)
    boff = np.concatenate([[0], np.cumsum(blen.astype(np.int64))[:-1]]).astype(np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_tr, d_pos = t(np.concatenate([x.reshape(-1) for x in tr])), t(np.concatenate([G[f"pos{i}"] for i in idx]))
    d_pri, d_sec = t(np.concatenate([G[f"pri{i}"] for i in idx])), t(np.concatenate([G[f"sec{i}"] for i in idx]))
    d_toff, d_tlen, d_boff, d_blen = t(toff), t(tlen), t(boff), t(blen)
    d_a1, d_a2 = (torch.zeros(len(idx), dtype=torch.float64, device=dev) for _ in range(2))
    P = lambda x: C.c_void_p(x.data_ptr())
    b = capi.FractionBatch(capi.Arena(P(d_tr), P(d_toff), P(d_tlen)), capi.Arena(P(d_pos), P(d_boff), P(d_blen)), P(d_pri), P(d_sec), 50, 50, len(idx), capi.TB_MEM_DEVICE)
    ctx._check(ctx._lib.tb_allelic_fraction(ctx._h, C.byref(b), P(d_a1), P(d_a2)))
    got = np.stack([d_a1.cpu().numpy(), d_a2.cpu().numpy()], 1)
    for row, i in zip(got, idx):
        assert row.tobytes() == G[f"out{i}"].tobytes(), i
