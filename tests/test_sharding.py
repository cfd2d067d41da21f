"""Host-side multi-GPU logic on CPU: cell-balanced partitioning and the world_size-2 result path
over the gloo backend (the DP itself needs no collective; see centrolign_b200/sharding.py)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from centrolign_b200.batch import AlignmentParameters, successor_form, synth_windows
from centrolign_b200.sharding import balanced_partition, run_sharded, stream_shard
from checkers import CpuChecker  # test infrastructure: tests/checkers.py


def test_balanced_partition_properties():
    rng = np.random.default_rng(1)
    cells = rng.integers(1, 10**8, 500)
    for parts_n in (1, 2, 4, 8):
        parts = balanced_partition(cells, parts_n)
        allw = np.concatenate(parts)
        assert sorted(allw.tolist()) == list(range(500))  # a partition: disjoint and complete
        loads = np.array([cells[p].sum() for p in parts])
        assert loads.max() - loads.min() <= cells.max()  # LPT bound
    assert [p.tolist() for p in balanced_partition([5, 5, 5], 3)] == [[0], [1], [2]]
    assert balanced_partition([], 2)[0].size == 0


def test_stream_shards_are_disjoint():
    spans = [stream_shard(r, 1000) for r in range(8)]
    assert spans == [(r * 1000, 1000) for r in range(8)]


def _oracle_runner(batch, params):
    chk = CpuChecker("port")
    res = [chk.po_poa(batch, w, params) for w in range(batch.n_windows)]
    return np.asarray([r[0] for r in res], np.int64), [r[1] for r in res]


def _pwfa_oracle_runner(batch, params):
    """The wavefront variant shards the same way (windows in successor form; prune limit of the Stitcher)."""
    chk = CpuChecker("port")
    res = [chk.pwfa_po_poa(batch, w, params, 50) for w in range(batch.n_windows)]
    return np.asarray([r[0] for r in res], np.int64), [r[1] for r in res]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = synth_windows(9, first_index=40, seed=13, len_min=20, len_max=200, alt_len=21, alt_period=100)
    out = run_sharded(batch, AlignmentParameters(), runner=_oracle_runner)
    wout = run_sharded(successor_form(batch), AlignmentParameters(), runner=_pwfa_oracle_runner)
    if rank == 0:
        q.put((out[0].tolist(), [a.tolist() for a in out[1]], wout[0].tolist(), [a.tolist() for a in wout[1]]))
    else:
        assert out is None and wout is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scores, alns, wscores, walns = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    batch = synth_windows(9, first_index=40, seed=13, len_min=20, len_max=200, alt_len=21, alt_period=100)
    want_s, want_a = _oracle_runner(batch, AlignmentParameters())
    assert scores == want_s.tolist()
    assert alns == [a.tolist() for a in want_a]
    want_ws, want_wa = _pwfa_oracle_runner(successor_form(batch), AlignmentParameters())
    assert wscores == want_ws.tolist()
    assert walns == [a.tolist() for a in want_wa]


def test_c_partition_matches_python_partition():
    """clb_balanced_partition (what clb_popoa_batch_multi deals windows to devices with) is the same
    longest-processing-time-first rule as sharding.balanced_partition."""
    from centrolign_b200.popoa import balanced_partition_c
    from centrolign_b200.sharding import balanced_partition

    rng = np.random.default_rng(5)
    for n, parts in ((0, 3), (1, 4), (57, 2), (1000, 8), (33, 5)):
        cells = (rng.integers(1, 2000, n) ** 2).astype(np.int64)
        if n > 10:
            cells[rng.integers(0, n, n // 3)] = cells[0]  # ties
        got = balanced_partition_c(cells, parts)
        want = np.zeros(n, np.int32)
        for k, ids in enumerate(balanced_partition(cells, parts)):
            want[ids] = k
        assert np.array_equal(got, want)
        if n >= 57:
            loads = np.bincount(got, weights=cells.astype(np.float64), minlength=parts)
            assert loads.max() - loads.min() <= cells.max()  # LPT bound
