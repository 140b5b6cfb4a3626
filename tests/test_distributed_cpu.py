"""World-size-2 `gloo` test of the data-parallel host logic (no GPU): graph sharding, the
1/global_batch gradient scaling and the single flat all-reduce.  The per-shard forward/backward
is played by the float64 oracle here; on the GPU box the same `distributed` functions wrap the
CUDA train step (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gcn_string_b200 as g
        from gcn_string_b200 import distributed as gd, synthetic
        from oracle import batching_ref, model_ref_np as O1

        assert gd.world() == (rank, world)
        ds = synthetic.make_dataset(7, seed=9, n_mean=40, deg=6, n_feat=6)       # 7 graphs: unequal shards (4 + 3)
        cfg = g.GNNConfig(in_features=6, output=2, activation="softmax", hidden=8, message_passing=2)
        specs = g.block_specs(cfg)
        w, s = g.init_params(cfg, seed=3, perturb=True)
        global_batch = 7

        def shard_grads(r):
            lo, hi = gd.shard_bounds(0, global_batch, r, world)
            graphs = [ds.graph(k) for k in range(lo, hi)]
            (x, (idx, _, _), seg), y = batching_ref.collate(graphs)
            res = O1.loss_and_grads(cfg, specs, w, s, x, idx[:, 0], idx[:, 1], seg, y, hi - lo)
            # the train step back-propagates with grad_scale = 1/global_batch instead of 1/local
            return res["grads"] * (hi - lo) / global_batch, hi - lo, res["loss"]

        mine, n_local, _ = shard_grads(rank)
        flat = torch.from_numpy(mine.copy())
        gd.allreduce_gradients(flat)                                            # the only collective
        both = [shard_grads(r) for r in range(world)]
        expect = sum(gr for gr, _, _ in both)                                   # = count-weighted mean of shard gradients
        counts = [n for _, n, _ in both]
        assert counts == [4, 3] and sum(counts) == global_batch
        np.testing.assert_allclose(flat.numpy(), expect, rtol=1e-12, atol=1e-15)
        weighted = sum(gr / (n / global_batch) * (n / global_batch) for gr, n, _ in both)
        np.testing.assert_allclose(expect, weighted, rtol=1e-12, atol=1e-15)
        # identical parameters on every rank after the step
        w_new = w.astype(np.float64) - 0.02 * flat.numpy()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(w_new))
        assert torch.equal(gathered[0], gathered[1])
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok_{r}.npy") for r in range(2))


def test_shard_bounds_tile_the_batch():
    sys.path.insert(0, ROOT)
    from gcn_string_b200.distributed import shard_bounds
    for n, world in ((10, 4), (7, 2), (3, 8), (1024, 8), (50, 3)):
        parts = [shard_bounds(100, 100 + n, r, world) for r in range(world)]
        assert parts[0][0] == 100 and parts[-1][1] == 100 + n
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_balanced_shards_partition_the_batch_and_even_out_the_work():
    """§8e: shards balanced by work.  Every graph on exactly one rank, the same shard sizes as the contiguous
    split, identical result on every rank, and a smaller spread of per-rank work on log-normal graph sizes."""
    sys.path.insert(0, ROOT)
    from gcn_string_b200 import distributed as gd
    rng = np.random.default_rng(0)
    for n, world in [(1024, 8), (1000, 8), (7, 2), (5, 4), (64, 1)]:
        ids = rng.permutation(10 * n)[:n]
        cost = np.rint(rng.lognormal(8.0, 0.5, n)).astype(np.int64)
        parts = [gd.balanced_shard(ids, cost, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == sorted(ids.tolist())
        want_sizes = [gd.shard_bounds(0, n, r, world)[1] - gd.shard_bounds(0, n, r, world)[0] for r in range(world)]
        assert [len(p) for p in parts] == want_sizes
        again = [gd.balanced_shard(ids, cost, r, world) for r in range(world)]
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))
        if n >= 1000:
            lookup = dict(zip(ids.tolist(), cost.tolist()))
            work = np.array([sum(lookup[i] for i in p.tolist()) for p in parts], dtype=np.float64)
            contiguous = np.array([cost[slice(*gd.shard_bounds(0, n, r, world))].sum() for r in range(world)], dtype=np.float64)
            assert work.max() / work.mean() < 1.002 < contiguous.max() / contiguous.mean()
