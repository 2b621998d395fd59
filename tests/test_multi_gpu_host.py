"""Host logic of the multi-GPU SpGEMM on CPU: world_size-2 gloo processes, each computing its row block
with the CPU oracle and gathering through nsparse_b200.multi_gpu.allgatherv_csr.  (The GPU kernels are
not involved: this covers partitioning, displacements and the row-pointer rebasing.)"""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mat(seed=0, m=300, n=300, dens=0.03):
    rng = np.random.default_rng(seed)
    a = sp.random(m, n, density=dens, random_state=rng, format="csr")
    a.data = rng.integers(1, 5, a.nnz).astype(np.float64)
    a.sort_indices()
    # make the load very uneven so the cuts are not the trivial halves
    a = sp.vstack([a[:10].toarray().repeat(1, axis=0) * 0 + 1, a[10:]]).tocsr()
    a.sort_indices()
    return a


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    from nsparse_b200 import multi_gpu as mg
    from nsparse_b200.csr import CSR
    from oracle import oracle

    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = _mat()
    A = CSR.from_scipy(a, np.float64)
    cuts, total = mg.partition_rows_by_ip(A.rpt, A.col, A.rpt, world)
    blk = mg.row_block(A, cuts[rank], cuts[rank + 1])
    rpt, col, val = oracle.spgemm(blk.rpt, blk.col, blk.val, A.rpt, A.col, A.val)
    out = mg.allgatherv_csr(torch.from_numpy(rpt), torch.from_numpy(col), torch.from_numpy(val), int(rpt[-1]),
                            cuts, A.M)
    full = oracle.spgemm(A.rpt, A.col, A.val, A.rpt, A.col, A.val)
    ok = (np.array_equal(out[0].numpy(), full[0]) and np.array_equal(out[1].numpy()[:out[3]], full[1])
          and np.array_equal(out[2].numpy()[:out[3]], full[2]) and out[3] == int(full[0][-1]))
    q.put((rank, ok, cuts, total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_block_allgatherv_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    cuts = res[0][2]
    assert cuts[0] == 0 and cuts[-1] == 300 and all(c0 <= c1 for c0, c1 in zip(cuts, cuts[1:]))


def test_partition_balances_intermediate_products():
    sys.path.insert(0, ROOT)
    from nsparse_b200 import multi_gpu as mg
    from oracle import oracle

    a = _mat(3, 2000, 2000, 0.01)
    ip = oracle.spgemm_intprod(a.indptr.astype(np.int32), a.indices.astype(np.int32), a.indptr.astype(np.int32))
    for parts in (1, 2, 4, 8):
        cuts, total = mg.partition_rows_by_ip(a.indptr, a.indices, a.indptr, parts)
        assert total == int(ip.sum()) and len(cuts) == parts + 1
        loads = [int(ip[cuts[i]:cuts[i + 1]].sum()) for i in range(parts)]
        assert sum(loads) == total
        assert max(loads) <= total / parts + ip.max()
    # degenerate: more parts than rows, empty matrix
    cuts, total = mg.partition_rows_by_ip(np.array([0, 1, 2]), np.array([0, 1]), np.array([0, 1, 2]), 4)
    assert cuts[0] == 0 and cuts[-1] == 2 and total == 2
    cuts, total = mg.partition_rows_by_ip(np.zeros(4, np.int64), np.zeros(0, np.int64), np.zeros(4, np.int64), 2)
    assert cuts == [0, 0, 3] and total == 0


def test_partition_rows_by_cost_balances_products_plus_output():
    """partition_rows_by_cost: blocks of ~equal  products + w * nnz(C_i); w = 0 is the equal-product cut."""
    import numpy as np

    from nsparse_b200 import gen, partition_rows_by_cost, partition_rows_by_ip
    from oracle import oracle

    a = gen.rmat_csr(11, 8, seed=2, dtype=np.float64, native=False, values="ones")
    c_rpt = oracle.spgemm(a.rpt, a.col, a.val, a.rpt, a.col, a.val, acc_double=True)[0].astype(np.int64)
    cuts0, ip0 = partition_rows_by_ip(a.rpt, a.col, a.rpt, 4)
    cuts1, ip1 = partition_rows_by_cost(a.rpt, a.col, a.rpt, c_rpt, 4, 0.0)
    assert cuts1 == cuts0 and ip1 == ip0
    w = 5.0
    cuts, _ = partition_rows_by_cost(a.rpt, a.col, a.rpt, c_rpt, 4, w)
    assert cuts[0] == 0 and cuts[-1] == a.M and all(x <= y for x, y in zip(cuts, cuts[1:]))
    blen = np.diff(a.rpt).astype(np.int64)
    ip_row = np.add.reduceat(np.concatenate([blen[a.col], [0]]), np.minimum(a.rpt[:-1], len(a.col)))
    ip_row[np.diff(a.rpt) == 0] = 0
    cost = ip_row + w * np.diff(c_rpt)
    per = [cost[cuts[i]:cuts[i + 1]].sum() for i in range(4)]
    assert max(per) - min(per) <= 2 * cost.max() + 1e-9     # within one (heaviest) row of each other


def test_partition_by_measured_times():
    """Feedback cut: equal times keep the equal-products cut; a rank that was twice as slow gets about half the
    products; cuts stay monotone and cover all rows."""
    import numpy as np

    import nsparse_b200 as ns
    from nsparse_b200 import gen

    a = gen.rmat_csr(12, 16, seed=3, dtype=np.float32)
    cuts, ip = ns.partition_rows_by_ip(a.rpt, a.col, a.rpt, 4)
    same, ip2 = ns.partition_rows_by_measured(a.rpt, a.col, a.rpt, cuts, [1.0] * 4, 4)
    assert ip2 == ip and all(abs(x - y) <= 1 for x, y in zip(same, cuts))      # (float targets: a row either way)
    new, _ = ns.partition_rows_by_measured(a.rpt, a.col, a.rpt, cuts, [2.0, 1.0, 1.0, 1.0], 4)
    assert new[0] == 0 and new[-1] == a.M and all(x <= y for x, y in zip(new, new[1:]))
    blen = np.diff(a.rpt).astype(np.int64)
    row_ip = np.add.reduceat(np.r_[blen[a.col], 0], np.minimum(a.rpt[:-1], a.nnz))
    row_ip[np.diff(a.rpt) == 0] = 0
    share0_old = row_ip[:cuts[1]].sum() / ip
    share0_new = row_ip[:new[1]].sum() / ip
    assert 0.2 < share0_old < 0.3 and share0_new < 0.75 * share0_old


def test_partition_minmax_bounds_compute_and_transfer():
    """partition_rows_minmax: with a fast link the cut balances the measured compute, with a slow link the entries of
    C every rank has to send; the largest block cost is never above that of the equal-products cut, with and without
    the faster tail (SM stores once the kernels have ended); cuts are monotone and cover all rows."""
    import numpy as np

    import nsparse_b200 as ns
    from nsparse_b200 import gen
    from oracle import oracle

    a = gen.rmat_csr(11, 16, seed=5, dtype=np.float32)
    c_rpt = oracle.spgemm(a.rpt, a.col, a.val, a.rpt, a.col, a.val)[0].astype(np.int64)
    n = 4
    cuts, ip = ns.partition_rows_by_ip(a.rpt, a.col, a.rpt, n)
    secs = [1e-3] * n
    blen = np.diff(a.rpt).astype(np.int64)
    ipp = np.concatenate([[0], np.cumsum(blen[a.col])])[a.rpt].astype(np.float64)

    def worst(cs, out_gbs, tail_gbs):
        w = 0.0
        for r in range(n):
            t = (ipp[cs[r + 1]] - ipp[cs[r]]) / ip * n * 1e-3
            ob = float(c_rpt[cs[r + 1]] - c_rpt[cs[r]]) * 8 * (n - 1)
            w = max(w, t + max(0.0, ob - out_gbs * 1e9 * t) / ((tail_gbs or out_gbs) * 1e9))
        return w

    for out_gbs, tail in ((1e6, None), (1e-3, None), (1e-3, 2e-3), (0.05, None), (0.05, 0.1)):
        new, ip2 = ns.partition_rows_minmax(a.rpt, a.col, a.rpt, c_rpt, cuts, secs, n, 8, out_gbs, tail)
        assert ip2 == ip and new[0] == 0 and new[-1] == a.M and all(x <= y for x, y in zip(new, new[1:]))
        assert worst(new, out_gbs, tail) <= worst(cuts, out_gbs, tail) * (1 + 1e-9)
        if out_gbs == 1e6:      # compute only: the equal-products cut again (a row either way)
            assert all(abs(x - y) <= 1 for x, y in zip(new, cuts))
        if out_gbs == 1e-3:     # transfer only: equal entries of C per rank, within the heaviest row
            per = np.diff(c_rpt[new])
            assert per.max() - per.min() <= 2 * np.diff(c_rpt).max() + 1


def test_device_partition_and_generators_on_cpu_tensors():
    """The torch versions used for the full-size C4 / C5 inputs (nsparse_b200.gen.*_device, partition_rows_by_ip_device)
    are plain torch: on CPU tensors they must reproduce the numpy / native results bit for bit."""
    import numpy as np
    import torch

    import nsparse_b200 as ns
    from nsparse_b200 import gen

    dev = torch.device("cpu")
    d = gen.rmat_csr_device(11, 16, 12345, np.float64, dev)
    h = gen.rmat_csr(11, 16, 12345, np.float64)
    assert np.array_equal(d.d_rpt.numpy(), h.rpt) and np.array_equal(d.d_col.numpy(), h.col) and np.array_equal(d.d_val.numpy(), h.val)
    assert ns.partition_rows_by_ip_device(d, d, 3) == ns.partition_rows_by_ip(h.rpt, h.col, h.rpt, 3)
    blk, hb = ns.row_block(d, 100, 700), ns.row_block(h, 100, 700)
    assert np.array_equal(blk.d_rpt.numpy(), hb.rpt) and np.array_equal(blk.d_col.numpy(), hb.col)
    e = gen.er_csr_device(3000, 500, 4, device=dev)
    c = e.d_col.numpy().reshape(3000, 4)
    assert (np.diff(c, axis=1) > 0).all() and c.min() >= 0 and c.max() < 500
    p = gen.powerlaw_csr_device(1 << 13, 64, 2048, device=dev)
    lens = np.diff(p.d_rpt.numpy())
    assert lens.max() == 2048 and lens.min() >= 1
    sub = p.rows_to_host([0, 5, 77])
    assert sub.M == 3 and sub.nnz == int(lens[[0, 5, 77]].sum())
