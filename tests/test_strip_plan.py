"""Row-strip split of one oversized layer: the exchange plan on CPU.

bidi_strip_plan is pure geometry (no device).  Checked here against a brute-force
restatement of what every phase of the split step reads (window centres as
sdr/RSDR.cpp:33-41 computes them), and exercised over two gloo ranks: each rank ships
exactly the rows the plan says its neighbour needs and receives exactly the rows it needs
itself -- the host logic of the NCCL transport, which the GPU tests then run for real."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bidinet_b200 import config as cf  # noqa: E402
from bidinet_b200 import strip_plan  # noqa: E402

GEOMS = [  # (in_w, in_h, hid_w, hid_h, r_ff, r_rec, r_lat, r_pred)
    (32, 32, 32, 32, 6, 6, 6, 6),
    (17, 29, 11, 23, 3, 2, 2, 3),
    (9, 40, 12, 16, 2, -1, 1, 4),
    (8, 12, 8, 36, 1, 3, 5, 0),
]


def _cfg(g):
    iw, ih, hw, hh, rff, rrec, rlat, rpred = g
    return cf.make_config(cf.KWINNERS, iw, ih, [dict(width=hw, height=hh, r_ff=rff, r_rec=rrec, r_lat=rlat, r_pred=rpred)])


def _centres(hh, vh):
    r = np.float32(vh) / np.float32(hh)
    v = (np.arange(hh, dtype=np.float32) * r).astype(np.float32)
    return np.floor(np.abs(v) + np.float32(0.5)).astype(int)  # std::round: half away from zero


def _clip(lo, hi, n):
    return max(0, min(n, lo)), max(0, min(n, hi))


def _within(need, have):
    return need[0] >= need[1] or (have[0] <= need[0] and need[1] <= have[1])


@pytest.mark.parametrize("g", GEOMS)
@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_plan_covers_what_each_phase_reads(g, n):
    iw, ih, hw, hh, rff, rrec, rlat, rpred = g
    cfg, ld = _cfg(g)
    cy = _centres(hh, ih)
    rr = max(rrec, 0)
    plans = [strip_plan(cfg, ld, p, n) for p in range(n)]
    # owned rows partition both grids
    assert [pl["OWN_HID"][0] for pl in plans] + [hh] == [0] + [pl["OWN_HID"][1] for pl in plans]
    assert [pl["OWN_VIS"][0] for pl in plans] + [ih] == [0] + [pl["OWN_VIS"][1] for pl in plans]
    for pl in plans:
        (o0, o1), (v0, v1) = pl["OWN_HID"], pl["OWN_VIS"]
        assert o1 > o0 and v1 > v0
        # units whose feed-forward window contains an owned visible row: their winners and
        # weights are read by the gather-form reconstruction / input prediction
        ef = [y for y in range(hh) if cy[y] + rff >= v0 and cy[y] - rff < v1]
        ef_span = (min(ef), max(ef) + 1) if ef else (o0, o0)
        assert _within(ef_span, pl["PRED_STATE"]) and _within(ef_span, pl["EXT_HID"])
        assert _within(_clip(o0 - rr, o1 + rr, hh), pl["EXT_HID"]) and _within((o0, o1), pl["EXT_HID"])
        e0, e1 = pl["EXT_HID"]
        assert _within(_clip(o0 - rlat, o1 + rlat, hh), pl["ACT"])
        for need in (pl["EXT_HID"], _clip(o0 - rpred, o1 + rpred, hh), _clip(e0 - rr, e1 + rr, hh)):
            assert _within(need, pl["STATE"])
        assert _within(_clip(e0 - rr, e1 + rr, hh), pl["HID_RECON"])
        lo = min(cy[y] - rff for y in range(e0, e1))
        hi = max(cy[y] + rff for y in range(e0, e1)) + 1
        assert _within(_clip(lo, hi, ih), pl["VIS_RECON"])


def test_plan_rejects_what_the_split_does_not_handle():
    import bidinet_b200
    cfg, ld = cf.config_c1(cf.KWINNERS)  # three layers
    with pytest.raises(bidinet_b200.BidiError, match="single-layer"):
        strip_plan(cfg, ld, 0, 2)
    cfg, ld = _cfg(GEOMS[0])
    with pytest.raises(bidinet_b200.BidiError, match="more parts than rows"):
        strip_plan(cfg, ld, 0, 33)


def test_c5_halo_is_radius_rows():
    """C5 (1024x1024, every radius 6) over 8 parts: an interior part needs 6 rows of
    activations either side (r x 1024 x 4 B = 24.6 kB per direction)."""
    cfg, ld = cf.config_c5(1024)
    pl = strip_plan(cfg, ld, 3, 8)
    assert pl["OWN_HID"] == (384, 512) and pl["ACT"] == (378, 518)
    assert pl["EXT_HID"] == (378, 518) and pl["STATE"] == (372, 524)


# ---- two gloo ranks move the rows the plan names
ITEMS = [("act", "ACT", "OWN_HID"), ("state", "STATE", "OWN_HID"), ("vis_recon", "VIS_RECON", "OWN_VIS"),
         ("recon", "HID_RECON", "OWN_HID"), ("p_state", "PRED_STATE", "OWN_HID")]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = GEOMS[1]
    iw, ih, hw, hh = g[:4]
    cfg, ld = _cfg(g)
    plans = [strip_plan(cfg, ld, p, world) for p in range(world)]
    me = plans[rank]
    ok = True
    for name, need, own in ITEMS:
        width, rows = (iw, ih) if own == "OWN_VIS" else (hw, hh)
        truth = torch.arange(rows * width, dtype=torch.float32).reshape(rows, width) + 1000.0 * len(name)
        mine = torch.full((rows, width), float("nan"))
        mine[me[own][0]:me[own][1]] = truth[me[own][0]:me[own][1]]   # a rank starts with its own rows only
        ops, recvs = [], []
        for peer in range(world):
            if peer == rank:
                continue
            lo, hi = max(plans[peer][need][0], me[own][0]), min(plans[peer][need][1], me[own][1])
            if lo < hi:
                ops.append(dist.P2POp(dist.isend, mine[lo:hi].clone(), peer))
            lo, hi = max(me[need][0], plans[peer][own][0]), min(me[need][1], plans[peer][own][1])
            if lo < hi:
                buf = torch.empty((hi - lo, width))
                recvs.append((lo, hi, buf))
                ops.append(dist.P2POp(dist.irecv, buf, peer))
        for w in dist.batch_isend_irecv(ops) if ops else []:
            w.wait()
        for lo, hi, buf in recvs:
            mine[lo:hi] = buf
        lo, hi = me[need]
        ok = ok and bool(torch.equal(mine[lo:hi], truth[lo:hi]))          # everything needed arrived, bit for bit
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_gloo_ranks_exchange_the_planned_halos():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, True), (1, True)]
