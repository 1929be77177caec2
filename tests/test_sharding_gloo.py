"""world_size-2 (and 3) gloo tests of the multi-GPU host logic, on CPU.

The engine shards genes and exchanges exactly three things: the partial Gram (sum), the top eigenvectors (from
rank 0) and the control-gene partials [A | Gm] (sum).  Here each rank plays one shard with numpy standing in for
the kernels, torch.distributed/gloo for NCCL, and the assembled result must equal the unsharded oracle --
which pins the partitioning rule, what is reduced, and that r = min(m - p, sum(ctl)) uses the GLOBAL control
count even when one shard holds no control gene at all.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from ruviii_b200.sharding import assemble_columns, gene_range, shard_columns  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir, ctl_in_first_shard_only):
    import torch
    import torch.distributed as dist
    from oracle import ruviii_oracle as O
    from ruviii_b200.synth import make_problem
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    p = make_problem(m1=90, n=403, groups=9, rep=3, n_ctl=50, k=4, seed=77)
    ctl = p.ctl.copy()
    if ctl_in_first_shard_only:
        g0, g1 = gene_range(403, 0, world)
        ctl[g1:] = False
    Y = np.vstack([p.Y, p.replicate_data])
    gid = O.group_ids_from_names(p.sample_names + p.replicate_names)
    Ys, ctls = shard_columns(Y, rank, world), shard_columns(ctl, rank, world)
    m = Y.shape[0]
    # K1 + K2 on the shard, exchange 1: allreduce of the partial Gram
    active, Y0a, pcols = O._group_mean_residual(Ys, gid)
    G = torch.from_numpy(Y0a @ Y0a.T)
    dist.all_reduce(G)
    n_ctl = torch.tensor([float(ctls.sum())], dtype=torch.float64)
    dist.all_reduce(n_ctl)
    r = min(m - pcols, int(n_ctl.item()))
    k = min(4, r)
    # K3 on rank 0, exchange 2: broadcast of U
    U = torch.zeros((Y0a.shape[0], k), dtype=torch.float64)
    if rank == 0:
        lam, V = np.linalg.eigh(G.numpy())
        U = torch.from_numpy(np.ascontiguousarray(V[:, np.argsort(lam)[::-1][:k]]))
    dist.broadcast(U, src=0)
    # K4 + K5 on the shard, exchange 3: allreduce of [A | Gm]
    alpha = U.numpy().T @ Y0a
    ac = alpha[:, ctls]
    AG = torch.from_numpy(np.concatenate([(Ys[:, ctls] @ ac.T).ravel(), (ac @ ac.T).ravel()]))
    dist.all_reduce(AG)
    A = AG.numpy()[:m * k].reshape(m, k)
    Gm = AG.numpy()[m * k:].reshape(k, k)
    A = A - A.mean(axis=0, keepdims=True)
    W = np.linalg.solve(Gm, A.T).T
    newY = Ys - W @ alpha                       # K6 on the shard, no exchange
    np.save(os.path.join(outdir, "newY_%d.npy" % rank), newY)
    np.save(os.path.join(outdir, "W_%d.npy" % rank), W)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,first_only", [(2, False), (3, False), (2, True)])
def test_gene_sharded_exchange_matches_unsharded_oracle(tmp_path, world, first_only):
    import torch.multiprocessing as mp
    from oracle import ruviii_oracle as O
    from ruviii_b200.synth import make_problem
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), first_only), nprocs=world, join=True)
    p = make_problem(m1=90, n=403, groups=9, rep=3, n_ctl=50, k=4, seed=77)
    ctl = p.ctl.copy()
    if first_only:
        ctl[gene_range(403, 0, world)[1]:] = False
    ref = O.ruvIII_literal(p.assay, p.sample_names, p.replicate_data, p.replicate_names, ctl, 4, apply_log=False)
    newY = assemble_columns([np.load(tmp_path / ("newY_%d.npy" % r)) for r in range(world)])
    assert O.rel_frob(newY, ref["newY"]) < 1e-10
    for r in range(1, world):                   # W is replicated, bit-identical after the allreduce
        assert np.array_equal(np.load(tmp_path / "W_0.npy"), np.load(tmp_path / ("W_%d.npy" % r)))


def _worker_gene_side(rank, world, port, outdir):
    """The gene-side regime over several shards, as engine.cu::do_run does it: all-to-all of row blocks of Z, a full
    n x n partial Gram per rank over its rows, allreduce, eigensolve on rank 0, alpha = diag(sigma) V' (appendix A-5)."""
    import torch
    import torch.distributed as dist
    from oracle import ruviii_oracle as O
    from ruviii_b200.synth import make_dense_design
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n, k = 61, 3
    Y, gid, ctl = make_dense_design(m=150, n=n, group_size=3, n_ctl=20, k=k, seed=41)
    Ys, ctls = shard_columns(Y, rank, world), shard_columns(ctl, rank, world)
    m = Y.shape[0]
    active, Z, pcols = O._group_mean_residual(Ys, gid)           # residual rows of this gene slab (R x n_loc)
    R = Z.shape[0]
    zrow = [R * j // world for j in range(world + 1)]            # the row split of engine.cu (h->zrow)
    edges = [gene_range(n, j, world) for j in range(world)]
    rows_me = zrow[rank + 1] - zrow[rank]
    # exchange 1: all-to-all -- rank q receives rows [zrow[q], zrow[q+1]) of every slab
    stage = [torch.zeros((rows_me, b - a), dtype=torch.float64) for a, b in edges]
    reqs = []
    for q in range(world):
        blk = torch.from_numpy(np.ascontiguousarray(Z[zrow[q]:zrow[q + 1]]))
        if q == rank:
            stage[q].copy_(blk)
        else:
            reqs.append(dist.isend(blk, dst=q))
            reqs.append(dist.irecv(stage[q], src=q))
    for r_ in reqs:
        r_.wait()
    Zrows = np.concatenate([t.numpy() for t in stage], axis=1)   # rows_me x n: all genes of my rows
    G = torch.from_numpy(Zrows.T @ Zrows)                        # exchange 2: allreduce of the n x n partials
    dist.all_reduce(G)
    n_ctl = torch.tensor([float(ctls.sum())], dtype=torch.float64)
    dist.all_reduce(n_ctl)
    assert min(m - pcols, int(n_ctl.item())) >= k
    V = torch.zeros((n, k), dtype=torch.float64)
    lam = torch.zeros(k, dtype=torch.float64)
    if rank == 0:
        w, Q = np.linalg.eigh(G.numpy())
        top = np.argsort(w)[::-1][:k]
        V, lam = torch.from_numpy(np.ascontiguousarray(Q[:, top])), torch.from_numpy(np.ascontiguousarray(w[top]))
    dist.broadcast(V, src=0)                                     # exchange 3: V and the eigenvalues
    dist.broadcast(lam, src=0)
    a, b = edges[rank]
    alpha = np.sqrt(lam.numpy())[:, None] * V.numpy()[a:b].T     # this shard's columns of diag(sigma) V'
    ac = alpha[:, ctls]
    AG = torch.from_numpy(np.concatenate([(Ys[:, ctls] @ ac.T).ravel(), (ac @ ac.T).ravel()]))
    dist.all_reduce(AG)                                          # exchange 4: [A | Gm]
    A = AG.numpy()[:m * k].reshape(m, k)
    Gm = AG.numpy()[m * k:].reshape(k, k)
    A = A - A.mean(axis=0, keepdims=True)
    W = np.linalg.solve(Gm, A.T).T
    np.save(os.path.join(outdir, "newY_%d.npy" % rank), Ys - W @ alpha)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gene_side_row_resplit_matches_unsharded_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    from oracle import ruviii_oracle as O
    from ruviii_b200.synth import make_dense_design
    mp.spawn(_worker_gene_side, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    Y, gid, ctl = make_dense_design(m=150, n=61, group_size=3, n_ctl=20, k=3, seed=41)
    ref = O.ruvIII_literal(Y.T, ["g%d" % g for g in gid], np.zeros((0, 61)), [], ctl, 3, apply_log=False)
    newY = assemble_columns([np.load(tmp_path / ("newY_%d.npy" % r)) for r in range(world)])
    assert O.rel_frob(newY, ref["newY"]) < 1e-10


def _worker_pca(rank, world, port, outdir):
    """computePCA's fast path over gene shards as ruviii_pca does it: centre per gene (local), partial sample-side Gram,
    allreduce, eigensolve on rank 0, broadcast of u and d^2, local projection v = t(Xc) u / d."""
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(12)
    X = rng.normal(size=(40, 97)) @ np.diag(np.linspace(3, 0.2, 97)) + rng.normal(size=(40, 1))
    k = 4
    Xs = shard_columns(X, rank, world)
    Xc = Xs - Xs.mean(axis=0, keepdims=True)
    G = torch.from_numpy(Xc @ Xc.T)
    dist.all_reduce(G)
    U = torch.zeros((40, k), dtype=torch.float64)
    lam = torch.zeros(k, dtype=torch.float64)
    if rank == 0:
        w, Q = np.linalg.eigh(G.numpy())
        top = np.argsort(w)[::-1][:k]
        U, lam = torch.from_numpy(np.ascontiguousarray(Q[:, top])), torch.from_numpy(np.ascontiguousarray(w[top]))
    dist.broadcast(U, src=0)
    dist.broadcast(lam, src=0)
    d = np.sqrt(lam.numpy())
    np.save(os.path.join(outdir, "v_%d.npy" % rank), (Xc.T @ U.numpy()) / d)
    if rank == 0:
        np.save(os.path.join(outdir, "u.npy"), U.numpy())
        np.save(os.path.join(outdir, "d.npy"), d)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pca_matches_unsharded_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    from oracle import pca_oracle as PCAO
    from oracle import ruviii_oracle as O
    mp.spawn(_worker_pca, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(12)
    X = rng.normal(size=(40, 97)) @ np.diag(np.linspace(3, 0.2, 97)) + rng.normal(size=(40, 1))
    ref = PCAO.fast_pca(X.T, 4, apply_log=False)
    v = np.concatenate([np.load(tmp_path / ("v_%d.npy" % r)) for r in range(world)], axis=0)
    u, d = np.load(tmp_path / "u.npy"), np.load(tmp_path / "d.npy")
    assert np.allclose(d, ref["d"], rtol=1e-12, atol=0)
    assert O.rel_frob((u * d) @ v.T, (ref["u"] * ref["d"]) @ ref["v"].T) < 1e-11


def test_gene_ranges_tile_the_genome():
    for n in (1, 7, 403, 20000):
        for world in (1, 2, 3, 4, 8):
            edges = [gene_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gene_range(10, 2, 2)
