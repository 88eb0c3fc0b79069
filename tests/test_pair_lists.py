"""Host-side work lists of the sparse-visibility Schur update (csrc/tscm_pair_lists.h):
every (frame, pair of cameras that both see it) exactly once, items of at most `chunk`
consecutive common frames, launched in the order of the first frame they touch, and per pair
an item list in frame order (what k_reduce_pairs sums in a fixed order)."""
import ctypes as C

import numpy as np
import pytest

from tscm_calib_b200 import synth

ip = C.POINTER(C.c_int)


def build(hostmath, vis, chunk):
    Cn, F = vis.shape
    cams, frames = np.nonzero(vis)                       # camera-major views, as tscm_problem wants
    vc, vf = cams.astype(np.int32), frames.astype(np.int32)
    V = len(vc)
    npairs = Cn * (Cn + 1) // 2
    cap_ent = F * npairs
    cap_items = cap_ent // max(1, chunk) + npairs * (F // max(1, chunk) + 2)
    sizes = np.zeros(2, np.int32)
    ent, rng = np.zeros((cap_ent, 2), np.int32), np.zeros((cap_items, 2), np.int32)
    pair_item, pair_items = np.zeros(npairs + 1, np.int32), np.zeros(cap_items, np.int32)
    rc = hostmath.hostmath_pair_lists(Cn, F, V, vc.ctypes.data_as(ip), vf.ctypes.data_as(ip), chunk, cap_ent,
                                      cap_items, sizes.ctypes.data_as(ip), ent.ctypes.data_as(ip),
                                      rng.ctypes.data_as(ip), pair_item.ctypes.data_as(ip),
                                      pair_items.ctypes.data_as(ip))
    assert rc == 0
    ne, ni = int(sizes[0]), int(sizes[1])
    return vc, vf, ent[:ne], rng[:ni], pair_item, pair_items[:ni]


@pytest.mark.parametrize("case,chunk", [("ring", 128), ("ring", 7), ("rig4", 16), ("dense", 5)])
def test_pair_lists_cover_every_common_frame_once(hostmath, case, chunk):
    vis = {"ring": lambda: synth.config(4, num_frames=600).visible,
           "rig4": lambda: synth.config(2).visible,
           "dense": lambda: np.ones((3, 23), dtype=bool)}[case]()
    Cn, F = vis.shape
    vc, vf, ent, rng, pair_item, pair_items = build(hostmath, vis, chunk)
    pairs = [(a, b) for a in range(Cn) for b in range(a, Cn)]
    # entries: the two views belong to the pair's cameras and to the same frame; the number of
    # entries is the number of (frame, pair) incidences
    expected = sum(int((vis[a] & vis[b]).sum()) for a, b in pairs)
    assert len(ent) == expected
    np.testing.assert_array_equal(vf[ent[:, 0]], vf[ent[:, 1]])
    # items: a partition of the entries, each within one pair, at most `chunk` long, frames increasing
    lengths = rng[:, 1] - rng[:, 0]
    assert lengths.min() >= 1 and lengths.max() <= chunk and lengths.sum() == len(ent)
    covered = np.zeros(len(ent), dtype=np.int32)
    first_frame = np.zeros(len(rng), dtype=np.int64)
    for k, (b0, b1) in enumerate(rng):
        covered[b0:b1] += 1
        e = ent[b0:b1]
        assert len(set(zip(vc[e[:, 0]], vc[e[:, 1]]))) == 1                 # one camera pair
        assert np.all(np.diff(vf[e[:, 0]]) > 0)                             # consecutive common frames
        first_frame[k] = vf[e[0, 0]]
    assert np.all(covered == 1)
    assert np.all(np.diff(first_frame) >= 0)                                # launch order = first frame
    # per pair: its items, in frame order, and all of them
    seen = np.zeros(len(rng), dtype=np.int32)
    for pr, (a, b) in enumerate(pairs):
        ids = pair_items[pair_item[pr]:pair_item[pr + 1]]
        seen[ids] += 1
        fr = []
        for k in ids:
            e = ent[rng[k, 0]:rng[k, 1]]
            assert vc[e[0, 0]] == a and vc[e[0, 1]] == b
            fr.append(vf[e[:, 0]])
        if len(ids):
            fr = np.concatenate(fr)
            np.testing.assert_array_equal(fr, np.flatnonzero(vis[a] & vis[b]))  # every common frame, in order
        else:
            assert not (vis[a] & vis[b]).any()
    assert np.all(seen == 1)
