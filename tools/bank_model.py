#!/usr/bin/env python3
"""Python model of Engine's shared-memory index maps (csrc/ntt_engine.cuh): enumerates the bank conflicts of every
scatter / gather of an exchange for a given (LOGN, LOGR, word bytes).  usage: bank_model.py LOGN LOGR WORDBYTES"""
import sys

def model(LOGN, LOGR, WB, verbose=True, FF=False):
    N, R = 1 << LOGN, 1 << LOGR
    T = N // R
    P = (LOGN + LOGR - 1) // LOGR
    RX = LOGN - (P - 1) * LOGR
    kFF = FF and P >= 2 and RX < LOGR
    R1 = LOGR if kFF else RX
    LOGROW = 5 if WB == 4 else 4
    PH = 1 << LOGROW
    s0 = lambda q: 0 if q == 0 else ((LOGN - LOGR if q == P - 1 else q * LOGR) if kFF else R1 + (q - 1) * LOGR)
    blk_words = lambda q: N if q == 0 else N >> s0(q)
    stride = lambda q: T if q == 0 else blk_words(q) >> LOGR
    wants_perm = lambda q: q >= 1 and stride(q) < PH and blk_words(q) >= PH and R < PH
    D = PH // R if R < PH else 1
    perm_feasible = (T // PH) >= D and ((T // PH) % D) == 0
    kxor = any(wants_perm(q) for q in range(P)) and not perm_feasible
    def decomp(q, tid):
        S = stride(q)
        if q == 0: return 0, tid
        if wants_perm(q) and not kxor:
            w, l = tid // PH, tid % PH
            return (w // D) * (D * (PH // S)) + (w % D) + D * (l // S), l % S
        return tid // S, tid % S
    def elem(q, tid, k):
        b, o = decomp(q, tid)
        return b * blk_words(q) + o + k * stride(q)
    swz = lambda i: i ^ ((i >> LOGR) & (PH - 1))
    sidx = (lambda i: swz(i)) if kxor else (lambda i: i + (i >> LOGROW))
    worst = 0
    for q in range(P):
        # every pass's layout is used for one side of an exchange (except both ends use only one)
        tot = 0; mx = 0
        for w0 in range(0, T, 32):
            lanes = range(w0, min(w0 + 32, T))
            for k in range(R):
                # a warp-wide access of WB bytes per lane: 64-bit accesses are served in two halves of 16 lanes
                groups = [list(lanes)] if WB == 4 else [list(lanes)[:16], list(lanes)[16:]]
                for g in groups:
                    banks = {}
                    for t in g:
                        a = sidx(elem(q, t, k)) * WB // 4
                        for h in range(WB // 4):
                            banks.setdefault((a + h) % 32, set()).add(a + h)
                    m = max(len(v) for v in banks.values()) if banks else 1
                    mx = max(mx, m); tot += m - 1
        if verbose: print("LOGN=%d LOGR=%d WB=%d pass %d: T=%d stride=%d perm=%s xor=%s  max way=%d extra wavefronts=%d" % (LOGN, LOGR, WB, q, T, stride(q), wants_perm(q), kxor, mx, tot))
        worst = max(worst, mx)
    return worst

if __name__ == "__main__":
    model(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), FF=len(sys.argv) > 4)
