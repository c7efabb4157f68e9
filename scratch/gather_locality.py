"""CPU experiment: L1 wavefronts (distinct 128-byte lines per warp load) of the gather kernel's scattered
record reads, for the round-1 packed slot layout vs a row-major full record layout."""
import sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import bench
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pr = bench.build_problem(L, 1.0)
D, R = pr['D'], pr['R'].tocsr()
n = D[0].shape[0]; B = 7; E = n // B
Eu = (D[1] @ R).tocsr(); Es = (D[3] @ R).tocsr()
m = R.shape[1]
def elem_cols(A):
    out = np.full((E, B), -1, dtype=np.int64)
    for e in range(E):
        c = np.unique(A.indices[A.indptr[e*B]:A.indptr[(e+1)*B]])
        out[e, :len(c)] = c
    return out
cu, cs = elem_cols(Eu), elem_cols(Es)
# own s col per point
own_s = np.full((E, B), -1)
for e in range(E):
    for l in range(B):
        i = e*B+l
        if Es.indptr[i+1] > Es.indptr[i]:
            own_s[e, l] = np.flatnonzero(cs[e] == Es.indices[Es.indptr[i]])[0]
LPE = 8
def tri(q, q2): return q*B - q*(q-1)//2 + (q2-q)
def packed(off, pk, npad): K = npad//LPE; return off + (pk % K)*LPE + pk//K
NT = 32
offA = dict(uu=0, us=32, ss=32+56); NSA = 96
def slotA(v1, q1, v2, q2):
    if (v1, q1) > (v2, q2): v1, q1, v2, q2 = v2, q2, v1, q1
    if v1 == 0 and v2 == 0: return packed(0, tri(q1, q2), NT)
    if v1 == 0 and v2 == 1: return offA['us'] + q1*LPE + q2
    return offA['ss'] + q1 if q1 == q2 else -1
NSB = 154
def slotB(v1, q1, v2, q2):   # row (v1,q1), col (v2,q2)
    if v1 == 0: return q1*14 + (q2 if v2 == 0 else 7 + q2)
    if v2 == 0: return 98 + q1*8 + q2
    return 98 + q1*8 + 7 if q1 == q2 else -1
rows, cols, srcA, srcB = [], [], [], []
for e in range(E):
    loc = [(0, q, cu[e, q]) for q in range(B) if cu[e, q] >= 0] + [(1, q, cs[e, q]) for q in range(B) if cs[e, q] >= 0]
    for (v1, q1, g1) in loc:
        for (v2, q2, g2) in loc:
            a = slotA(v1, q1, v2, q2)
            if a < 0: continue
            rows.append(g1); cols.append(g2); srcA.append(e*NSA + a); srcB.append(e*NSB + slotB(v1, q1, v2, q2))
rows = np.array(rows); cols = np.array(cols); srcA = np.array(srcA); srcB = np.array(srcB)
key = rows * m + cols
order = np.argsort(key, kind='stable')
key = key[order]; srcA = srcA[order]; srcB = srcB[order]
uniq, first, cnt = np.unique(key, return_index=True, return_counts=True)
nnz = len(uniq)
print('L', L, 'E', E, 'm', m, 'nnzH', nnz, 'contribs', len(key), 'cnt hist', np.bincount(cnt)[:10])
def wavefronts(src, name, line=16):
    tot = 0
    for j in range(2):   # first / second contribution of the two-wide ELL
        has = cnt > j
        idx = np.where(has & (cnt <= 2), first + j, -1)
        ln = np.where(idx >= 0, src[np.maximum(idx, 0)] // line, -1)
        pad = (-len(ln)) % 32
        ln = np.concatenate([ln, np.full(pad, -1)]).reshape(-1, 32)
        ln.sort(axis=1)
        d = (np.diff(ln, axis=1) != 0).sum(axis=1) + 1 - (ln[:, 0] == -1)   # distinct non-negative values
        tot += d.sum()
    print(name, 'wavefronts', tot, 'per entry', tot / nnz)
    return tot
wavefronts(srcA, 'round-1 packed layout (96/elem)')
wavefronts(srcB, 'row-major full layout (154/elem)')
wavefronts(srcA, 'packed, 32B sectors', line=4)
wavefronts(srcB, 'row-major, 32B sectors', line=4)
def variant(RU, RS, NSX, name):
    def slot(v1, q1, v2, q2):
        if v1 == 0: return q1*RU + (q2 if v2 == 0 else 7 + q2)
        if v2 == 0: return 7*RU + q1*RS + q2
        return 7*RU + q1*RS + 7 if q1 == q2 else -1
    src = []
    for e in range(E):
        loc = [(0, q, cu[e, q]) for q in range(B) if cu[e, q] >= 0] + [(1, q, cs[e, q]) for q in range(B) if cs[e, q] >= 0]
        for (v1, q1, g1) in loc:
            for (v2, q2, g2) in loc:
                if slotA(v1, q1, v2, q2) < 0: continue
                src.append(e*NSX + slot(v1, q1, v2, q2))
    src = np.array(src)[order]
    wavefronts(src, name)
variant(16, 8, 168, 'u-run 16 s-run 8 (168)')
variant(16, 8, 176, 'u-run 16 s-run 8, record 176 (line aligned)')
