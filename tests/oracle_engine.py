"""TEST INFRASTRUCTURE: an Engine look-alike that answers from the CPU oracle.

It lets the *host* logic of the drop-in layer (window arithmetic, read chopping, driver control flow, the
coroutine scheduler, output writers) be checked against the reference's golden tables on a machine without a
GPU.  It lives under tests/ and is never importable from the product package; the GPU tests run the same
checks through the real ``vapor_b200.engine.Engine``."""
import numpy as np

from oracle import batch_oracle as BO
from oracle import vapor_oracle as O
from vapor_b200.engine import Results


class OracleEngine:
    def close(self):
        pass

    def score(self, batch):
        e = BO.score_batch(batch, with_hits=True)
        return Results(e["task_score"], e["task_status"], e["task_stat"], e["task_hits"], e["task_hitsum"],
                       e["sv_qs"], e["sv_gs"], e["sv_gq"], e["sv_gt"], e["sv_nscore"])

    def dotdata(self, k, seq1, seq2):
        s1 = seq1 if isinstance(seq1, str) else bytes(seq1).decode("latin-1")
        s2 = seq2 if isinstance(seq2, str) else bytes(seq2).decode("latin-1")
        return O.dotdata(int(k), s1, s2).astype(np.int32)

    def selfplot_qc(self, seqs, ks):
        out = np.zeros((len(seqs), 8), dtype=np.int64)
        for i, (s, k) in enumerate(zip(seqs, ks)):
            s = s if isinstance(s, str) else bytes(s).decode("latin-1")
            try:
                d = O.dotdata(int(k), s, s)
            except KeyError:
                out[i, 7] = 2
                continue
            low = d[d[:, 0] > d[:, 1]]
            out[i, :3] = [len(d), int((d[:, 0] == d[:, 1]).sum()), len(low)]
            out[i, 3:7] = [low[:, 0].min(), low[:, 0].max(), low[:, 1].min(), low[:, 1].max()] if len(low) else [0xFFFFFFFF, 0, 0xFFFFFFFF, 0]
            out[i, 7] = 1
        return out

    def summarize(self, score_lists):
        n = len(score_lists)
        qs = np.zeros(n); gs = np.zeros(n); gq = np.zeros(n); gt = np.full(n, 255, np.uint8); ns = np.zeros(n, np.int32)
        for i, s in enumerate(score_lists):
            ns[i] = len(s)
            r = O.summarize_sv([float(v) for v in s])
            if r is not None:
                qs[i], gs[i], gq[i], gt[i] = r["QS"], r["GS"], r["GQ"], r["GT"]
        return qs, gs, gq, gt, ns
