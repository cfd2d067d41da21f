"""GPU parity tests of the sparse anchor-chaining DP: ``clb_chain_dp`` (hand-written sm_100a kernel) must
return exactly the chain the unmodified reference returns (tests/golden/chain_golden.npz, produced by
oracle/chain_shim.cpp from Anchorer::sparse_chain_dp / sparse_affine_chain_dp), match for match -- ties
between equal-scoring chains included, which is where the reference's search-tree traversal order shows."""
import numpy as np
import pytest

from centrolign_b200.chain import ChainProblem, ChainStats, chain_dp
from checkers import chain_oracle  # test infrastructure: tests/checkers.py
from golden_io import load_chain_golden

pytestmark = pytest.mark.gpu

GOLD = load_chain_golden()
CASES = [(c, k) for c in sorted(GOLD) for k in sorted(GOLD[c])]


def _check_chain_consistency(prob, chain, dp, bp):
    """Size-independent properties: the chain follows back-pointers, and every DP value is reproduced by its
    back-pointer's value plus its own weight plus a non-positive gap term (exactly zero for the gap-free DP)."""
    for a, b in zip(chain[:-1], chain[1:]):
        assert bp[b] == a
    if len(chain):
        assert bp[chain[0]] == -1
    w = prob.arrays["weight"]
    has = bp >= 0
    idx = np.nonzero(has)[0]
    gap = dp[idx].astype(np.float64) - w[idx] - dp[bp[idx]].astype(np.float64)
    assert np.all(gap <= 1e-3)
    if prob.num_pw == 0:
        assert np.all(dp[idx] == (dp[bp[idx]] + w[idx]).astype(np.float32))


@pytest.mark.parametrize("case,kind", CASES)
def test_chain_equals_reference(case, kind):
    prob = GOLD[case][kind]
    st = ChainStats()
    chain, dp, bp, opt = chain_dp(prob, stats=st)
    small = st.tree_bytes <= 200 * 1024  # whole problem in shared memory: one kernel; else preparation + persistent DP kernel
    assert st.kernel_launches == (1 if small else 2) and st.steps == prob.n_step
    if case == "pair600" and kind == "gapfree":
        assert small, "the tiny gap-free fixture is meant to exercise chain_small_kernel"
    assert len(chain) == len(prob.expect_chain), f"{case}/{kind}: chain length {len(chain)} != {len(prob.expect_chain)}"
    assert np.array_equal(chain, prob.expect_chain), f"{case}/{kind}: chain differs from the reference's"
    _check_chain_consistency(prob, chain, dp, bp)


@pytest.mark.parametrize("grid", [1, 8])
def test_multi_cta_gives_the_same_values(grid, monkeypatch):
    """One CTA (block barriers) and a cooperative multi-CTA grid (grid-wide barriers) must agree bit for bit."""
    prob = GOLD["msa4_3k"]["affine"]
    ref_chain, ref_dp, ref_bp, ref_opt = chain_dp(prob)
    monkeypatch.setenv("CLB_CHAIN_GRID", str(grid))
    chain, dp, bp, opt = chain_dp(prob)
    assert np.array_equal(chain, ref_chain) and np.array_equal(bp, ref_bp)
    assert np.array_equal(dp.view(np.uint32), ref_dp.view(np.uint32)) and opt == ref_opt


@pytest.mark.parametrize("mode", ["CLB_CHAIN_CLUSTER=1", "CLB_CHAIN_CLUSTER=2", "CLB_CHAIN_CLUSTER=8", "CLB_CHAIN_CLUSTER=16", "CLB_CHAIN_GRID=8",
                                  "CLB_CHAIN_GRID=148"])
@pytest.mark.parametrize("num_pw", [0, 1, 2, 3])
def test_every_way_to_separate_the_phases_and_every_piece_count(mode, num_pw, monkeypatch):
    """One CTA, a thread-block cluster of 2 / 8 / 16 CTAs and a cooperative grid must give every DP value and back-pointer
    of the oracle, for the gap-free instance of the step loop (1 item kind), the three-piece instance (7) and the generic one
    (1 or 2 gap pieces: the same matches chained with truncated gap parameters)."""
    src = GOLD["msa4_3k"]["gapfree" if num_pw == 0 else "affine"]
    k = max(1, num_pw)
    prob = ChainProblem(num_pw, tuple(src.gap_open[:k]), tuple(src.gap_extend[:k]), src.scale, src.n_chain1, src.n_chain2, src.min_score, src.arrays)
    name, value = mode.split("=")
    monkeypatch.setenv(name, value)
    monkeypatch.setenv("CLB_CHAIN_NO_SMALL", "1")  # not the shared-memory kernel: that one is a single CTA by construction
    chain, dp, bp, opt = chain_dp(prob)
    ochain, odp, obp, oopt = chain_oracle(prob)
    assert np.array_equal(chain, ochain) and opt == oopt
    assert np.array_equal(dp.view(np.uint32), odp.view(np.uint32)) and np.array_equal(bp, obp)


@pytest.mark.parametrize("case,kind", CASES)
def test_every_dp_value_and_backpointer_equals_the_oracle(case, kind):
    """Stronger than chain identity: all DP values bit for bit and all back-pointers, against the literal C
    restatement of the reference's search trees (oracle/chain_oracle.c)."""
    prob = GOLD[case][kind]
    chain, dp, bp, opt = chain_dp(prob)
    ochain, odp, obp, oopt = chain_oracle(prob)
    assert np.array_equal(chain, ochain) and opt == oopt
    assert np.array_equal(dp.view(np.uint32), odp.view(np.uint32))
    assert np.array_equal(bp, obp)


def _handmade(num_pw, with_query=True, n=2):
    """Two matches on one path pair; the second starts behind the first (offset 3 < query offset 5)."""
    a = dict(weight=np.array([1.5, 2.0], np.float32)[:n], dp_init=np.array([1.5, 2.0], np.float32)[:n],
             final_term=np.zeros(n, np.float32), end_off=np.array([0, 1, 2], np.int64)[:n + 1], end_match=np.arange(n, dtype=np.uint32),
             qry_off=np.array([0, 0, 1 if with_query else 0], np.int64)[:n + 1],
             qry_match=np.array([1] if with_query and n == 2 else [], np.uint32),
             qry_chain1=np.array([0] if with_query and n == 2 else [], np.uint32), ins_off=np.array([0, 1, 2], np.int64)[:n + 1],
             ins_p1=np.zeros(n, np.uint32), ins_p2=np.zeros(n, np.uint32), ins_shift=np.zeros(n, np.int32),
             ins_offset=np.array([3, 9], np.uint32)[:n], ins_active=np.ones(n, np.uint8), qa1=np.zeros(n, np.int32),
             qa2=np.zeros(n, np.int32), qoff=np.array([0, 5], np.uint32)[:n])
    k = max(1, num_pw)
    return ChainProblem(num_pw, (1.0, 2.0, 3.0)[:k], (0.5, 0.2, 0.1)[:k], 1.0, 1, 1, 0.0, a)


@pytest.mark.parametrize("num_pw", [0, 1, 3])
def test_handmade_edge_cases(num_pw):
    for prob in (_handmade(num_pw), _handmade(num_pw, with_query=False), _handmade(num_pw, n=1)):
        chain, dp, bp, opt = chain_dp(prob)
        ochain, odp, obp, oopt = chain_oracle(prob)
        assert np.array_equal(chain, ochain) and np.array_equal(dp, odp) and np.array_equal(bp, obp) and opt == oopt
    chain, dp, bp, opt = chain_dp(_handmade(num_pw))
    assert chain.tolist() == [0, 1] and dp.tolist() == [1.5, 3.5] and bp.tolist() == [-1, 0]
    empty = _handmade(num_pw, n=1)
    empty.arrays = {k: v[:0] if k not in ("end_off", "qry_off", "ins_off") else v[:1] for k, v in empty.arrays.items()}
    chain, dp, bp, opt = chain_dp(empty)  # no matches at all: the empty chain
    assert len(chain) == 0


def test_batched_call_equals_one_by_one():
    """clb_chain_dp_batch (one launch, a CTA per problem for everything that fits shared memory; the rest one by one)
    must return, per problem, exactly what clb_chain_dp returns: chains, every DP value, every back-pointer."""
    from centrolign_b200.chain import ChainStats, chain_dp_batch

    probs = []
    for case in sorted(GOLD):
        for kind in sorted(GOLD[case]):
            probs.append(GOLD[case][kind])
    for num_pw in (0, 1, 3):  # tiny problems, incl. one without matches
        probs += [_handmade(num_pw), _handmade(num_pw, with_query=False), _handmade(num_pw, n=1)]
        empty = _handmade(num_pw, n=1)
        empty.arrays = {k: v[:0] if k not in ("end_off", "qry_off", "ins_off") else v[:1] for k, v in empty.arrays.items()}
        probs.append(empty)
    probs = probs * 3  # several CTAs' worth
    st = ChainStats()
    got = chain_dp_batch(probs, stats=st)
    assert len(got) == len(probs)
    n_small = 0
    for prob, (chain, dp, bp, opt) in zip(probs, got):
        rchain, rdp, rbp, ropt = chain_dp(prob)
        assert np.array_equal(chain, rchain) and opt == ropt
        assert np.array_equal(dp.view(np.uint32), rdp.view(np.uint32)) and np.array_equal(bp, rbp)
        if prob.expect_chain is not None and len(prob.expect_chain):
            assert np.array_equal(chain, prob.expect_chain)
        n_small += 1
    assert st.kernel_launches >= 1
    assert chain_dp_batch([]) == []


def test_jobs_created_on_threads_and_run_in_one_launch():
    """clb_chain_job_create on several host threads + ONE clb_chain_jobs_run (what the drop-in Anchorer's fill-in pool does,
    hostcpp/chain_batcher.hpp) returns per problem exactly what clb_chain_dp returns; problems that fit shared memory, problems
    that run from global memory inside the batch, and problems solved at creation are all in the mix."""
    from centrolign_b200.chain import chain_dp_jobs

    probs = []
    for case in sorted(GOLD):
        for kind in sorted(GOLD[case]):
            probs.append(GOLD[case][kind])
    for num_pw in (0, 1, 3):
        probs += [_handmade(num_pw), _handmade(num_pw, with_query=False), _handmade(num_pw, n=1)]
    probs = probs * 2
    for threads in (1, 8):
        got = chain_dp_jobs(probs, threads=threads)
        for prob, (chain, dp, bp, opt) in zip(probs, got):
            rchain, rdp, rbp, ropt = chain_dp(prob)
            assert np.array_equal(chain, rchain) and opt == ropt
            assert np.array_equal(dp.view(np.uint32), rdp.view(np.uint32)) and np.array_equal(bp, rbp)
    assert chain_dp_jobs([]) == []
