"""Host logic of the wide fused layer kernel (csrc/fused_wide.cu), on the CPU: the issue schedule
built for hidden width 384 / 512 conditioners covers every non-zero weight tile, carries the right
barrier obligations, and survives an independent, randomly interleaved replay of the kernel's
barrier protocol (no deadlock, no mbarrier running two phases ahead of its waiter)."""

import ctypes
import random

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from zuko_b200 import _engine as E

FIRST, LAST, AWAIT, AFREE, OUT = 8, 16, 32, 64, 128

CASES = {
    "cfg2_nsf": lambda: zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3),  # hidden width 256: 2 chunks, 4 K blocks
    "cfg3_maf": lambda: zuko.flows.MAF(32, 0, transforms=2, hidden_features=[512] * 4),
    "cfg5_nsf": lambda: zuko.flows.NSF(64, 16, transforms=2, bins=16, hidden_features=[512] * 3),
    "nsf24_k8_h512": lambda: zuko.flows.NSF(24, 0, transforms=2, bins=8, hidden_features=[512, 512]),  # classes not aligned to chunks
    "nsf10c3_h384": lambda: zuko.flows.NSF(10, 3, transforms=2, bins=8, hidden_features=[384, 384]),
    "maf100c28_h512": lambda: zuko.flows.MAF(100, 28, transforms=2, hidden_features=[512] * 2),     # K0 = 128: 2 input K blocks, 2 affine chunks
}


def _conditioner(layer):
    masks = [m.mask.numpy().astype(np.uint8) for m in layer.hyper if hasattr(m, "mask")]
    dims = [masks[0].shape[1]] + [m.shape[0] for m in masks]
    return dims, masks


def _schedule(dims, masks, uni, bins, D, C):
    L = len(masks)
    keep = [np.ascontiguousarray(m) for m in masks]
    ptrs = (ctypes.c_void_p * L)(*[m.ctypes.data for m in keep])
    cdims = (ctypes.c_int * (L + 1))(*dims)
    out = np.zeros(2 * 4096, np.uint32)
    rd = np.zeros(8, np.uint32)
    perm = np.zeros(sum(dims[1:-1]), np.int32)
    n = E.lib().zk_debug_wide_schedule(L, cdims, ptrs, uni, bins, D, C, out.ctypes.data, 4096, rd.ctypes.data, perm.ctypes.data)
    return n, out[: 2 * max(n, 0)].reshape(-1, 2), rd, perm


def _layer_args(layer):
    uni = E.ZK_UNI_RQS if layer.total > 2 else E.ZK_UNI_AFFINE
    bins = (layer.total + 1) // 3 if uni == E.ZK_UNI_RQS else 0
    return uni, bins


@pytest.mark.parametrize("name", list(CASES))
def test_schedule_covers_every_nonzero_tile(name):
    torch.manual_seed(0)
    flow = CASES[name]()
    for layer in flow.transform.transforms:  # the second layer has the reversed order
        dims, masks = _conditioner(layer)
        D, C = layer.features, layer.context
        uni, bins = _layer_args(layer)
        n, items, rd, perm = _schedule(dims, masks, uni, bins, D, C)
        assert n > 0, n
        L = len(masks)
        P = 3 * bins - 1 if uni == E.ZK_UNI_RQS else 2
        DPC = (4 if bins == 8 else 2) if uni == E.ZK_UNI_RQS else 64
        perms, o = [], 0
        for l in range(L - 1):
            perms.append(perm[o : o + dims[l + 1]])
            o += dims[l + 1]
            assert sorted(perms[-1].tolist()) == list(range(dims[l + 1]))
        flags, n0s = items[:, 0], items[:, 1]
        layer_of = (flags >> 16) & 7
        assert np.all(np.diff(layer_of.astype(int)) >= 0)  # layers in order
        for l in range(L):
            M = masks[l].astype(bool)
            if l < L - 1:
                M = M[perms[l]]
            if l > 0:
                M = M[:, perms[l - 1]]
            sel = layer_of == l
            fl, nn0 = flags[sel], n0s[sel]
            width = DPC * P if l == L - 1 else 128
            tiles = {(int(a) // width, int(f) & 7) for f, a in zip(fl, nn0)}
            assert len(tiles) == sel.sum()  # no tile twice
            rows, cols = np.nonzero(M)
            need = set(zip((rows // width).tolist(), (cols // 64).tolist()))
            assert need <= tiles
            # extra tiles only as the single "define the accumulator" tile of an all-zero chunk
            for ch, kb in tiles - need:
                assert kb == 0 and not any(c == ch for c, _ in need)
            assert int(rd[l]) == sum(1 << kb for kb in {kb for _, kb in tiles})
            assert bool(fl[0] & OUT) == (l == L - 1)
            # descending chunk / K-block order, first / last flags, first-read and last-read flags
            order = [(int(a) // width, int(f) & 7) for f, a in zip(fl, nn0)]
            assert order == sorted(order, reverse=True)
            seen_kb, last_idx = set(), {}
            for i, (ch, kb) in enumerate(order):
                assert bool(fl[i] & FIRST) == (i == 0 or order[i - 1][0] != ch)
                assert bool(fl[i] & LAST) == (i == len(order) - 1 or order[i + 1][0] != ch)
                assert bool(fl[i] & AWAIT) == (kb not in seen_kb)
                seen_kb.add(kb)
                last_idx[kb] = i
            for i, (ch, kb) in enumerate(order):
                assert bool(fl[i] & AFREE) == (last_idx[kb] == i)
            # issued 16-column MMA steps: exactly up to the last column of the block that any row of the chunk reads
            for i, (ch, kb) in enumerate(order):
                blk = M[ch * width : (ch + 1) * width, kb * 64 : (kb + 1) * 64]
                used = np.nonzero(blk.any(axis=0))[0]
                steps = 4 - ((int(fl[i]) >> 20) & 3)
                assert steps == (int(used.max()) // 16 + 1 if used.size else 1), (l, ch, kb)


def _replay(items, rd, L, nch, KB0, n_last, seed, tiles=3):
    """Independent model of the kernel's barrier protocol with a random scheduler."""
    rng = random.Random(seed)
    n = len(items)
    done = {("df", b): 0 for b in range(2)} | {("de", b): 0 for b in range(2)}
    done |= {("ar", k): 0 for k in range(8)} | {("af", k): 0 for k in range(8)}
    seen = dict.fromkeys(done, 0)

    def complete(key):
        assert done[key] == seen[key], f"{key} would run two phases ahead of its waiter"
        done[key] += 1

    def consume(key):
        if done[key] <= seen[key]:
            return False
        seen[key] += 1
        return True

    # MMA side: entries; completion of an entry's MMAs is a separate, later event (in order)
    state = {"mi": 0, "c": 0, "pending": []}

    def mma_issue():
        if state["mi"] >= tiles * n:
            return False
        f = int(items[state["mi"] % n][0])
        c = state["c"] + (1 if state["mi"] > 0 and f & FIRST else 0)
        buf = c & 1
        if f & FIRST and done[("de", buf)] < (c >> 1):
            return False
        unread = [k for k in range(8) if (f >> (8 + k)) & 1]
        if any(done[("ar", k)] <= seen[("ar", k)] for k in unread):
            return False
        if f & AWAIT and done[("ar", f & 7)] <= seen[("ar", f & 7)]:
            return False
        if f & FIRST:
            seen[("de", buf)] = c >> 1
        for k in unread:
            seen[("ar", k)] += 1
        if f & AWAIT:
            seen[("ar", f & 7)] += 1
        ev = []
        if f & AFREE:
            ev.append(("af", f & 7))
        if f & LAST:
            ev.append(("df", buf))
        state["pending"].append(ev)
        state["c"] = c
        state["mi"] += 1
        return True

    def mma_complete():
        if not state["pending"]:
            return False
        for key in state["pending"].pop(0):
            complete(key)
        return True

    steps = []
    for t in range(tiles):
        chunk = t * ((L - 1) * nch + n_last)
        steps.append(("stage",))
        for l in range(L - 1):
            for ch in range(nch - 1, -1, -1):
                steps += [("wait_df", chunk), ("drain", chunk & 1)]
                steps += [("wait_af", kb) for kb in (2 * ch, 2 * ch + 1) if (int(rd[l]) >> kb) & 1]
                steps.append(("write", ch))
                chunk += 1
        for ch in range(n_last - 1, -1, -1):
            steps += [("wait_df", chunk), ("drain", chunk & 1)]
            chunk += 1
        steps += [("wait_af", kb) for kb in range(8) if (int(rd[L - 1]) >> kb) & 1]
    ei = [0]

    def epi():
        if ei[0] >= len(steps):
            return False
        s = steps[ei[0]]
        if s[0] == "stage":
            for kb in range(KB0):
                complete(("ar", kb))
        elif s[0] == "wait_df":
            buf, need = s[1] & 1, (s[1] >> 1) + 1
            if done[("df", buf)] < need:
                return False
            seen[("df", buf)] = need
        elif s[0] == "drain":
            complete(("de", s[1]))
        elif s[0] == "wait_af":
            if not consume(("af", s[1])):
                return False
        elif s[0] == "write":
            complete(("ar", 2 * s[1]))
            complete(("ar", 2 * s[1] + 1))
        ei[0] += 1
        return True

    agents = [mma_issue, mma_complete, epi]
    while True:
        order = agents[:]
        rng.shuffle(order)
        if not any(a() for a in order):
            break
    assert state["mi"] == tiles * n and not state["pending"] and ei[0] == len(steps), "deadlock"


@pytest.mark.parametrize("name", list(CASES))
def test_schedule_survives_random_replays_of_the_protocol(name):
    torch.manual_seed(0)
    flow = CASES[name]()
    for layer in flow.transform.transforms:
        dims, masks = _conditioner(layer)
        D, C = layer.features, layer.context
        uni, bins = _layer_args(layer)
        n, items, rd, _ = _schedule(dims, masks, uni, bins, D, C)
        assert n > 0
        DPC = (4 if bins == 8 else 2) if uni == E.ZK_UNI_RQS else 64
        for seed in range(20):
            _replay(items, rd, len(masks), dims[1] // 128, (dims[0] + 63) // 64, (D + DPC - 1) // DPC, seed)


def test_dense_wide_conditioner_is_rejected_not_deadlocked():
    """A dense [512, 512] conditioner cannot be updated in place (every chunk reads every K block):
    the dry run must reject it, so such layers stay on the per-layer GEMM path."""
    dims = [64, 512, 512, 128]
    n, *_ = _schedule(dims, [np.ones((dims[i + 1], dims[i]), np.uint8) for i in range(3)], E.ZK_UNI_AFFINE, 0, 64, 0)
    assert n == -2


def test_wide_input_with_context_is_rejected_not_deadlocked():
    """K0 = 400 puts the context columns (read by every hidden unit) into the K blocks the first
    chunk's outputs would overwrite: rejected by the dry run as well."""
    torch.manual_seed(0)
    layer = zuko.flows.MAF(300, 100, transforms=1, hidden_features=[512] * 2).transform.transforms[0]
    dims, masks = _conditioner(layer)
    assert _schedule(dims, masks, E.ZK_UNI_AFFINE, 0, 300, 100)[0] == -2


def test_unsupported_shapes_are_reported():
    ones = lambda d: [np.ones((d[i + 1], d[i]), np.uint8) for i in range(len(d) - 1)]  # noqa: E731
    assert _schedule([16, 128, 128, 32], ones([16, 128, 128, 32]), E.ZK_UNI_AFFINE, 0, 16, 0)[0] == -1  # narrow kernel's shape
    assert _schedule([16, 512, 384, 32], ones([16, 512, 384, 32]), E.ZK_UNI_AFFINE, 0, 16, 0)[0] == -1  # unequal hidden widths


# --------------------------------------------------------------------------- #
# dual-tile kernel (csrc/fused_dual.cu): hidden width 128 / 256, two sub-tiles interleaved chunk by chunk
# --------------------------------------------------------------------------- #

DUAL_CASES = {
    "cfg2_nsf": lambda: zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3),
    "nsf7_k16_h128": lambda: zuko.flows.NSF(7, 0, transforms=2, bins=16, hidden_features=[128, 128]),
    "maf32_h256": lambda: zuko.flows.MAF(32, 0, transforms=2, hidden_features=[256] * 2),
    "nsf24_k8_h256": lambda: zuko.flows.NSF(24, 0, transforms=2, bins=8, hidden_features=[256, 256]),  # classes not aligned to chunks
}


def _dual_schedule(dims, masks, uni, bins, D, C):
    L = len(masks)
    keep = [np.ascontiguousarray(m) for m in masks]
    ptrs = (ctypes.c_void_p * L)(*[m.ctypes.data for m in keep])
    cdims = (ctypes.c_int * (L + 1))(*dims)
    out = np.zeros(2 * 4096, np.uint32)
    rd = np.zeros(8, np.uint32)
    n = E.lib().zk_debug_dual_schedule(L, cdims, ptrs, uni, bins, D, C, out.ctypes.data, 4096, rd.ctypes.data, None)
    return n, out[: 2 * max(n, 0)].reshape(-1, 2), rd


@pytest.mark.parametrize("name", list(DUAL_CASES))
def test_dual_schedule_is_two_interleaved_copies_of_the_tile_walk(name):
    """Each sub-tile's entries, taken alone, are the same descending walk over the non-zero tiles (same
    tiles, same order, same first / last / first-read / last-read flags); the two alternate chunk by chunk;
    and the schedule survives a randomly interleaved replay of the three agents (MMA side, two epilogue groups)."""
    torch.manual_seed(0)
    flow = DUAL_CASES[name]()
    for layer in flow.transform.transforms:
        dims, masks = _conditioner(layer)
        uni, bins = _layer_args(layer)
        n, items, rd = _dual_schedule(dims, masks, uni, bins, layer.features, layer.context)
        assert n > 0 and n % 2 == 0, n
        flags = items[:, 0].astype(np.int64)
        sub = (flags >> 2) & 1
        strip = lambda f: f & ~np.int64(4) & ~np.int64(0xFF00)  # noqa: E731  (sub-tile bit, unread masks)
        a, b = items[sub == 0], items[sub == 1]
        assert len(a) == len(b) == n // 2
        assert np.array_equal(strip(a[:, 0].astype(np.int64)), strip(b[:, 0].astype(np.int64))) and np.array_equal(a[:, 1], b[:, 1])
        # chunks alternate: a run of sub-tile 0 entries (one chunk) is followed by the same chunk of sub-tile 1
        runs = []
        for f in flags:
            if f & FIRST:
                runs.append(int((f >> 2) & 1))
        assert runs == [0, 1] * (len(runs) // 2)
        # descending tile order per sub-tile and layer
        layer_of = (a[:, 0].astype(np.int64) >> 16) & 7
        for l in range(len(masks)):
            sel = layer_of == l
            order = [(int(y), int(f) & 3) for f, y in zip(a[sel, 0], a[sel, 1])]
            assert order == sorted(order, reverse=True)
        L, nch, KB0 = len(masks), dims[1] // 128, (dims[0] + 63) // 64
        DPC = (4 if bins == 8 else 2) if uni == E.ZK_UNI_RQS else 64
        n_last = (layer.features + DPC - 1) // DPC
        for seed in range(20):
            _replay_dual(items, rd, L, nch, KB0, n_last, seed)


def _replay_dual(items, rd, L, nch, KB0, n_last, seed, tiles=3):
    rng = random.Random(seed)
    n = len(items)
    keys = [("df", u) for u in range(2)] + [("de", u) for u in range(2)] + [("ar", k) for k in range(8)] + [("af", k) for k in range(8)]
    done, seen = dict.fromkeys(keys, 0), dict.fromkeys(keys, 0)

    def complete(key):
        assert done[key] == seen[key], f"{key} would run two phases ahead of its waiter"
        done[key] += 1

    st = {"mi": 0, "cnt": [0, 0], "pending": []}

    def mma_issue():
        if st["mi"] >= tiles * n:
            return False
        f = int(items[st["mi"] % n][0])
        u = (f >> 2) & 1
        if f & FIRST and done[("de", u)] < st["cnt"][u]:
            return False
        unread = [k for k in range(8) if (f >> (8 + k)) & 1]
        if any(done[("ar", k)] <= seen[("ar", k)] for k in unread):
            return False
        kb8 = 4 * u + (f & 3)
        if f & AWAIT and done[("ar", kb8)] <= seen[("ar", kb8)]:
            return False
        if f & FIRST:
            seen[("de", u)] = st["cnt"][u]
            st["cnt"][u] += 1
        for k in unread:
            seen[("ar", k)] += 1
        if f & AWAIT:
            seen[("ar", kb8)] += 1
        ev = ([("af", kb8)] if f & AFREE else []) + ([("df", u)] if f & LAST else [])
        st["pending"].append(ev)
        st["mi"] += 1
        return True

    def mma_complete():
        if not st["pending"]:
            return False
        for key in st["pending"].pop(0):
            complete(key)
        return True

    steps = [[], []]
    for u in range(2):
        chunk = 0
        for _ in range(tiles):
            steps[u].append(("stage",))
            for l in range(L - 1):
                for ch in range(nch - 1, -1, -1):
                    steps[u] += [("wait_df", chunk), ("drain",)]
                    steps[u] += [("wait_af", kb) for kb in (2 * ch, 2 * ch + 1) if (int(rd[l]) >> kb) & 1]
                    steps[u].append(("write", ch))
                    chunk += 1
            for ch in range(n_last - 1, -1, -1):
                steps[u] += [("wait_df", chunk), ("drain",)]
                chunk += 1
            steps[u] += [("wait_af", kb) for kb in range(4) if (int(rd[L - 1]) >> kb) & 1]
    ei = [0, 0]

    def epi(u):
        def run():
            if ei[u] >= len(steps[u]):
                return False
            s = steps[u][ei[u]]
            if s[0] == "stage":
                for kb in range(KB0):
                    complete(("ar", 4 * u + kb))
            elif s[0] == "wait_df":
                if done[("df", u)] < s[1] + 1:
                    return False
                seen[("df", u)] = s[1] + 1
            elif s[0] == "drain":
                complete(("de", u))
            elif s[0] == "wait_af":
                k = ("af", 4 * u + s[1])
                if done[k] <= seen[k]:
                    return False
                seen[k] += 1
            elif s[0] == "write":
                complete(("ar", 4 * u + 2 * s[1]))
                complete(("ar", 4 * u + 2 * s[1] + 1))
            ei[u] += 1
            return True

        return run

    agents = [mma_issue, mma_complete, epi(0), epi(1)]
    while True:
        order = agents[:]
        rng.shuffle(order)
        if not any(a() for a in order):
            break
    assert st["mi"] == tiles * n and not st["pending"] and ei[0] == len(steps[0]) and ei[1] == len(steps[1]), "deadlock"
