"""Crude timing model of the dual kernel's issue order (cycles): compares the lockstep order with every
interleaving of a half-tile phase shift between the two sub-tiles.  Constants read off
profiles/r02_dual_timeline_recapture.txt; predicts 53 k cycles per tile for the current order against 63.6 k
measured (17 % optimistic).  Used once, for the decision recorded in DESIGN.md 5b."""
import itertools
HID, OUT = 850, 650
DRAIN_H, WRITE_H = 1400, 2400
EPI_O, TAIL, STAGE = 3700, 1100, 3600
# per sub-tile chunk list for cfg2: (kind, layer, ch, kbs)
def chunks():
    c = []
    c.append(("H", 0, 1, [0])); c.append(("H", 0, 0, [0]))
    for l in (1, 2):
        c.append(("H", l, 1, [3, 2, 1, 0])); c.append(("H", l, 0, [1, 0]))
    for ch, n in ((3, 4), (2, 3), (1, 2), (0, 1)):
        c.append(("O", 3, ch, list(range(n - 1, -1, -1))))
    return c
CH = chunks()
NH = 6
def simulate(order, n_iter=6):
    """order: list of (u, chunk_index, lag) for one cycle."""
    t_pipe = 0.0
    # per sub-tile state
    st = [dict(acc_free=0.0, a_ready={}, stage=STAGE if u == 0 else STAGE, tile_end=[]) for u in range(2)]
    for u in range(2):
        st[u]["a_ready"] = {(0, 0): STAGE}
    ends = [[], []]
    seq = []
    for it in range(n_iter + 1):
        for (u, ci, lag) in order:
            if lag and it == 0: continue
            if (not lag) and it == n_iter: continue
            seq.append((u, ci))
    marks = []
    for (u, ci) in seq:
        kind, l, ch, kbs = CH[ci]
        s = st[u]
        t = max(t_pipe, s["acc_free"])
        for kb in kbs:
            if kind == "H" or True:
                need = s["a_ready"].get((l, kb), 0.0)
            t = max(t, need)
            t += HID if kind == "H" else OUT
        t_pipe = t
        d_full = t + 200
        if kind == "H":
            s["acc_free"] = d_full + DRAIN_H
            w = d_full + WRITE_H
            s["a_ready"][(l + 1, 2 * ch)] = w
            s["a_ready"][(l + 1, 2 * ch + 1)] = w
        else:
            s["acc_free"] = d_full + EPI_O
            if ch == 0:
                end = d_full + EPI_O + TAIL
                ends[u].append(end)
                s["a_ready"] = {(0, 0): end + STAGE}
                s["acc_free"] = end
    per = [(e[-1] - e[1]) / (len(e) - 2) for e in ends]
    return per, ends
lock = []
for ci in range(10):
    for u in range(2):
        lock.append((u, ci, 0))
print("lockstep", simulate(lock)[0])
def merge(a, b, pattern):
    out, ia, ib = [], 0, 0
    for p in pattern:
        if p == "a": out.append(a[ia]); ia += 1
        else: out.append(b[ib]); ib += 1
    assert ia == len(a) and ib == len(b)
    return out
H0 = [(0, i, 0) for i in range(6)]; O0 = [(0, i, 0) for i in range(6, 10)]
H1 = [(1, i, 0) for i in range(6)]; O1 = [(1, i, 1) for i in range(6, 10)]
best = []
for pat in set(itertools.permutations("aaaaaabbbb")):
    pat = "".join(pat)
    order = merge(H0, O1, pat) + merge(H1, O0, pat)
    per, _ = simulate(order)
    best.append((max(per), pat))
best.sort()
for b in best[:10]: print(b)
print(best[-1])
