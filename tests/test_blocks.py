"""Block construction: COO -> CSC.  CPU part: the oracle against a numpy stable argsort.
-m gpu part: lg_block_csc through the C ABI against the oracle, on real sampler output and on edge cases."""
import numpy as np
import pytest

from conftest import make_sets, small_graph


def _numpy_csc(src, dst, num_dst):
    order = np.argsort(dst, kind="stable").astype(np.int32)
    indptr = np.zeros(num_dst + 1, np.int32)
    np.cumsum(np.bincount(dst, minlength=num_dst), out=indptr[1:])
    return indptr, src[order], order


@pytest.mark.parametrize("e,num_dst", [(0, 0), (0, 5), (1, 1), (1000, 17), (5000, 5000), (4096, 3)])
def test_oracle_block_csc_is_a_stable_sort(oracle, e, num_dst):
    rng = np.random.default_rng(e + num_dst)
    dst = rng.integers(0, max(num_dst, 1), e).astype(np.int32)
    src = rng.integers(0, 100000, e).astype(np.int32)
    ip, ix, eid = oracle.block_csc(src, dst, num_dst)
    w_ip, w_ix, w_eid = _numpy_csc(src, dst, num_dst)
    assert np.array_equal(ip, w_ip) and np.array_equal(ix, w_ix) and np.array_equal(eid, w_eid)


@pytest.mark.gpu
@pytest.mark.parametrize("e,num_dst", [(0, 0), (0, 5), (1, 1), (1000, 17), (5000, 5000), (4096, 3), (200000, 70000)])
def test_block_csc_matches_oracle_random(oracle, e, num_dst):
    torch = pytest.importorskip("torch")
    from legion_b200.blocks import BlockBuilder
    rng = np.random.default_rng(e * 7 + num_dst)
    dst = rng.integers(0, max(num_dst, 1), e).astype(np.int32)
    src = rng.integers(0, 1 << 22, e).astype(np.int32)
    bb = BlockBuilder(max(e, 1))
    ip, ix, eid = bb.csc(torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda(), num_dst)
    torch.cuda.synchronize()
    w_ip, w_ix, w_eid = oracle.block_csc(src, dst, num_dst)
    assert np.array_equal(ip.cpu().numpy(), w_ip)
    assert np.array_equal(ix.cpu().numpy(), w_ix) and np.array_equal(eid.cpu().numpy(), w_eid)
    ip2, ix2, none = bb.csc(torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda(), num_dst, with_eids=False)
    assert none is None and torch.equal(ip2, ip) and torch.equal(ix2, ix)


@pytest.mark.gpu
@pytest.mark.parametrize("fanout", [[25, 10], [15, 10, 5]])
def test_blocks_of_a_sampled_batch(oracle, fanout):
    """every block of a real batch (cumulative COO views, training_backend/ipc_cuda_kernel.cu:200-232): CSC equals the
    oracle's, and a segment-sum over the CSC equals the COO scatter-add (what the GNN layer computes)"""
    torch = pytest.importorskip("torch")
    from gpu_util import Rig
    from legion_b200 import synth
    from legion_b200.blocks import BlockBuilder
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = synth.features(0, N, 8, 5)
    ids, labels = make_sets(N)
    B, H = 128, len(fanout)
    rig = Rig(indptr, indices, feat, fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    rig.dp.run_once(rig.dp.params(d_ids, d_lab, B, 1, seed=77, batch_id=1), buf)
    torch.cuda.synchronize()
    nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
    bb = BlockBuilder(int(ec[9 + H]))
    # all blocks in one set of launches, sizes read on the device (what the server runs): bounds = the buffers' capacity
    max_edges, max_dst, per, nodes, edges = [], [], B, B, 0
    for f in fanout:
        max_dst.append(nodes)
        per *= f
        edges += per
        nodes += per
        max_edges.append(edges)
    multi = bb.csc_batch(buf.c, H, max_edges, max_dst)
    torch.cuda.synchronize()
    for h in range(H, 0, -1):
        e, num_dst = int(ec[9 + h]), int(nc[9 + h - 1])
        w_ip, w_ix, w_eid = oracle.block_csc(buf.agg_src[:e].cpu().numpy(), buf.agg_dst[:e].cpu().numpy(), num_dst)
        ip, ix, eid = multi[h - 1]
        assert np.array_equal(ip[:num_dst + 1].cpu().numpy(), w_ip), h
        assert np.array_equal(ix[:e].cpu().numpy(), w_ix) and np.array_equal(eid[:e].cpu().numpy(), w_eid), h
    for h in range(H, 0, -1):
        e, num_src, num_dst = int(ec[9 + h]), int(nc[9 + h]), int(nc[9 + h - 1])
        src, dst = buf.agg_src[:e], buf.agg_dst[:e]
        ip, ix, eid = bb.csc(src, dst, num_dst)
        torch.cuda.synchronize()
        w_ip, w_ix, w_eid = oracle.block_csc(src.cpu().numpy(), dst.cpu().numpy(), num_dst)
        assert np.array_equal(ip.cpu().numpy(), w_ip) and np.array_equal(ix.cpu().numpy(), w_ix)
        assert np.array_equal(eid.cpu().numpy(), w_eid)
        assert int(ip[-1]) == e and int(ix.max()) < num_src
        # mean-aggregation numerator both ways (integer-valued so the sums are exact in fp32)
        x = torch.arange(num_src, device="cuda", dtype=torch.float32) % 13
        coo = torch.zeros(num_dst, device="cuda").index_add_(0, dst.long(), x[src.long()])
        seg = torch.segment_reduce(x[ix.long()], "sum", lengths=(ip[1:] - ip[:-1]).long(), unsafe=True) if e else coo
        assert torch.equal(coo, seg)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["one_run", "runs_across_tiles", "sorted_already", "reversed", "three_passes", "hub", "full_size"])
def test_block_csc_structured_inputs(oracle, case):
    """the run-compressed radix path (csrc/blocks.cu) on the shapes it is built for and on the degenerate ones: one single
    run, runs that straddle the 4096-edge tiles, destinations beyond 2^22 (three digit passes), one hub destination whose
    runs are spread over the whole list, and a full-size [25,10] block (2.2 M edges in runs of <= 25)"""
    torch = pytest.importorskip("torch")
    from legion_b200.blocks import BlockBuilder
    rng = np.random.default_rng(len(case))
    if case == "one_run":
        dst, num_dst = np.full(10000, 3, np.int32), 7
    elif case == "runs_across_tiles":
        lens = rng.integers(1, 900, 60)
        dst, num_dst = np.repeat(rng.integers(0, 25, 60), lens).astype(np.int32), 25
    elif case == "sorted_already":
        dst, num_dst = np.repeat(np.arange(3000, dtype=np.int32), rng.integers(0, 12, 3000)), 3000
    elif case == "reversed":
        dst, num_dst = np.repeat(np.arange(3000, dtype=np.int32)[::-1], rng.integers(1, 12, 3000)), 3100
    elif case == "three_passes":
        num_dst = 5_000_000
        dst = np.repeat(rng.integers(0, num_dst, 20000), rng.integers(1, 6, 20000)).astype(np.int32)
    elif case == "hub":
        keys = rng.integers(0, 50000, 120000)
        keys[::3] = 4242  # one destination owns a third of the runs
        dst, num_dst = np.repeat(keys, rng.integers(1, 11, len(keys))).astype(np.int32), 50000
    else:  # hop 1: 8000 seeds x <= 25, hop 2: 200 k frontier entries x <= 10
        k1 = np.arange(8000)
        k2 = rng.integers(0, 130000, 200000)
        dst = np.concatenate([np.repeat(k1, rng.integers(0, 26, 8000)), np.repeat(k2, rng.integers(0, 11, 200000))]).astype(np.int32)
        num_dst = 130000
    dst = np.ascontiguousarray(dst)
    src = rng.integers(0, 1 << 21, len(dst)).astype(np.int32)
    bb = BlockBuilder(len(dst))
    ip, ix, eid = bb.csc(torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda(), num_dst)
    torch.cuda.synchronize()
    w_ip, w_ix, w_eid = oracle.block_csc(src, dst, num_dst)
    assert np.array_equal(ip.cpu().numpy(), w_ip)
    assert np.array_equal(eid.cpu().numpy(), w_eid) and np.array_equal(ix.cpu().numpy(), w_ix)
