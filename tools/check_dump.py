#!/usr/bin/env python
"""Replay the window batches a caller sent to clb_popoa_batch (written with CLB_DUMP_DIR=dir) and compare every window
on the GPU -- warp-per-window kernel and strip kernel -- with the C oracle.  Debugging aid for end-to-end mismatches:
    CLB_DUMP_DIR=/tmp/d oracle/_ref/centrolign_b200 ... ; python tools/check_dump.py /tmp/d"""
import glob, os, sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from centrolign_b200.batch import AlignmentParameters, GraphSide, WindowBatch  # noqa: E402
from centrolign_b200.popoa import po_poa_batch  # noqa: E402
from checkers import CpuChecker  # noqa: E402


def read_batch(path):
    buf = open(path, "rb").read()
    pos = 0

    def take(dt, n):
        nonlocal pos
        a = np.frombuffer(buf, dtype=dt, count=n, offset=pos).copy()
        pos += a.nbytes
        return a

    nw = int(take(np.int64, 1)[0])
    prm = take(np.int32, 1 + 2 + 3 + 3)
    params = AlignmentParameters(int(prm[1]), int(prm[2]), tuple(int(x) for x in prm[3:3 + prm[0]]), tuple(int(x) for x in prm[6:6 + prm[0]]))
    sides = []
    for _ in range(2):
        N, E, S, K = (int(x) for x in take(np.int64, 4))
        node_off = take(np.int64, nw + 1); label = take(np.uint8, N); edge_off = take(np.int64, nw + 1)
        pred_off = take(np.uint32, N + nw); pred = take(np.uint32, E); src_off = take(np.int64, nw + 1)
        src = take(np.uint32, S); snk_off = take(np.int64, nw + 1); snk = take(np.uint32, K)
        sides.append(GraphSide(node_off, label, edge_off, pred_off, pred, src_off, src, snk_off, snk))
    return WindowBatch(sides[0], sides[1]), params


def main():
    files = sorted(glob.glob(os.path.join(sys.argv[1], "popoa_*.bin")), key=lambda p: int(p.split("_")[-1].split(".")[0]))
    oracle = CpuChecker("port")
    os.environ.pop("CLB_DUMP_DIR", None)
    bad = tot = 0
    for path in files:
        batch, params = read_batch(path)
        res = {}
        for name, env in (("small", None), ("strip", "1")):
            if env:
                os.environ["CLB_NO_SMALL_WINDOWS"] = env
            try:
                res[name] = po_poa_batch(batch, params)
            finally:
                os.environ.pop("CLB_NO_SMALL_WINDOWS", None)
        for w in range(batch.n_windows):
            tot += 1
            s, a = oracle.po_poa(batch, w, params)
            for name in res:
                if s != res[name][0][w] or not np.array_equal(a, res[name][1][w]):
                    bad += 1
                    if bad <= 20:
                        lab1, po1, pr1, src1, snk1 = batch.g1.window(w)
                        lab2, po2, pr2, src2, snk2 = batch.g2.window(w)
                        print(f"{os.path.basename(path)} window {w} [{name}]: n1={len(lab1)} n2={len(lab2)} num_pw={params.num_pw} score gpu {res[name][0][w]} oracle {s}; "
                              f"aln equal {np.array_equal(a, res[name][1][w])}")
                        print("   g1", bytes(lab1).decode(), po1.tolist(), pr1.tolist(), src1.tolist(), snk1.tolist())
                        print("   g2", bytes(lab2).decode(), po2.tolist(), pr2.tolist(), src2.tolist(), snk2.tolist())
                        print("   gpu", res[name][1][w].tolist()); print("   ora", a.tolist())
    print(f"checked {tot} windows of {len(files)} batches: {bad} mismatches")


if __name__ == "__main__":
    main()
