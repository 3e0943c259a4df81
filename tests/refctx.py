"""Context stand-in for CPU tests of the host pipelines (tracy_b200.subcommands): the call shapes of tracy_b200.Context, every
call served by the reference's own functions behind oracle/_ref. TEST INFRASTRUCTURE -- it exists so that the file plumbing,
exit codes and writers are checked in the CPU suite; the same pipelines run on the CUDA kernels in the -m gpu tests."""
import os
import tempfile

import numpy as np

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore
from oracle import loader


class RefContext:
    def __init__(self, ref):
        self.ref = ref

    def read_traces(self, files):
        out = []
        for f in files:
            with tempfile.NamedTemporaryFile(suffix=".trace", delete=False) as fh:
                fh.write(bytes(f))
            try:
                t = self.ref.read_trace(fh.name)
            finally:
                os.unlink(fh.name)
            ragged = len({len(c) for c in t["samples"]}) != 1
            out.append(dict(format=t["format"], ok=t["ok"], status=3 if ragged else 0, traceACGT=None if ragged or t["format"] < 0 else np.stack(t["samples"]),
                            basecallpos=t["basecallpos"], qual=t["qual"], basecalls1=t["basecalls1"], basecalls2=t["basecalls2"]))
        return out

    def basecall(self, traces, ploc, sigratio=0.33):
        return [self.ref.basecall(t, p, sigratio) for t, p in zip(traces, ploc)]

    def create_profile(self, traces, bcpos, primary, secondary, trim_left=None, trim_right=None):
        n = len(traces)
        tl = np.broadcast_to(0 if trim_left is None else trim_left, (n,))
        tr = np.broadcast_to(0 if trim_right is None else trim_right, (n,))
        return [self.ref.create_profile(traces[i], bcpos[i], primary[i], secondary[i], int(tl[i]), int(tr[i])) for i in range(n)]

    def revcomp_profile(self, profiles):
        return [self.ref.revcomp_profile(p) for p in profiles]

    def gotoh(self, kind, a1, a2, sc=DnaScore(3, -5, -10, -4), ac=AlignConfig(True, False), traceback=True, out=None, rows=None, packed=False):
        s4 = (sc.match, sc.mismatch, sc.go, sc.ge)
        h, v = int(ac.horizontal), int(ac.vertical)

        def items(a):                                              # an Arena of offsets into one packed array, or a plain list
            if isinstance(a, tracy_b200.Arena):
                if a.base.dtype == np.float32:
                    return [a.base[o: o + 6 * n_].reshape(6, n_) for o, n_ in zip(a.off, a.len)]
                return [a.base[o: o + n_].tobytes() for o, n_ in zip(a.off, a.len)]
            return a
        a1, a2 = items(a1), items(a2)
        n = len(a1)
        scores = np.zeros(n, np.int32)
        if not traceback:
            for i in range(n):
                scores[i] = self.ref.gotoh_score(a1[i], a2[i], h, v, s4)
            return scores, None, None
        res = [self.ref.gotoh(a1[i], a2[i], h, v, s4) for i in range(n)]
        stride = max(max((len(r[1]) for r in res), default=1), 1)
        ops, ol = np.zeros((n, stride), np.uint8), np.zeros(n, np.int32)
        for i, (s, r0, r1) in enumerate(res):
            scores[i] = s
            o = loader.ops_from_rows(r0, r1)
            ops[i, : len(o)] = np.frombuffer(o, np.uint8)
            ol[i] = len(o)
        return scores, ops, ol

    def decompose_sweep(self, refrows, primaries, secondaries, vi_end, align_index, var_index, ndel, nins, grid=False):
        """The sweeps live inside decomposeAlleles in the reference; the C restatement (oracle/gotoh_oracle.c, pinned on
        decomposeAlleles' goldens) serves them one trace at a time."""
        port = loader.port()
        n = len(refrows)
        S = int(max(1, max(ndel, default=1), max(nins, default=1)))
        fref, fins = np.zeros((n, S), np.int32), np.zeros((n, S), np.int32)
        g = np.zeros((n, S, S), np.int32) if grid else None
        for t in range(n):
            a, b, c = port.decompose_sweep(bytes(refrows[t]), bytes(primaries[t]), bytes(secondaries[t]), int(vi_end[t]), int(align_index[t]), int(var_index[t]),
                                           int(ndel[t]), int(nins[t]), grid)
            fref[t, : len(a)] = a
            fins[t, : len(b)] = b
            if grid and c is not None:
                c = np.asarray(c)
                g[t, : c.shape[0], : c.shape[1]] = c
        return fref, fins, g

    def allelic_fraction(self, traces, bcpos, primary, secdecompose, trim_left=50, trim_right=50):
        return np.array([self.ref.allelic_fraction(traces[i], bcpos[i], bytes(primary[i]), bytes(secdecompose[i]), trim_left, trim_right)
                         for i in range(len(traces))], np.float64).reshape(-1, 2)
