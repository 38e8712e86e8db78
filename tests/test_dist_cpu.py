"""CPU, world_size 2, gloo: the host-side sharding / gather logic of ml_unigen_b200.dist."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, N, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_unigen_b200 import dist as ugd
    s, e = ugd.shard_range(total, rank, world)
    ids = (torch.arange(s, e)[:, None] * 1000 + torch.arange(N)[None, :]).to(torch.int64)
    scores = torch.arange(s, e, dtype=torch.float32) * -0.25 + (torch.arange(s, e) % 3).float()
    all_ids, all_scores = ugd.gather_ids_scores(ids, scores, total)
    exp_ids = (torch.arange(total)[:, None] * 1000 + torch.arange(N)[None, :]).to(torch.int64)
    exp_scores = torch.arange(total, dtype=torch.float32) * -0.25 + (torch.arange(total) % 3).float()
    ok = torch.equal(all_ids, exp_ids) and torch.equal(all_scores, exp_scores)
    top = ugd.select_topk(all_scores, 4)
    ok = ok and torch.equal(top, torch.sort(exp_scores, descending=True, stable=True).indices[:4])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_ids_scores_world2(total):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + total
    mp.spawn(_worker, args=(2, port, total, 5, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_range_partitions():
    from ml_unigen_b200.dist import shard_range
    for total in (1, 7, 16, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_candidate_scores_ignores_known_tokens():
    from ml_unigen_b200.dist import candidate_scores
    p = torch.tensor([[0.5, 0.25, torch.finfo(torch.float32).max], [1.0, 1.0, 1.0]])
    s = candidate_scores(p)
    assert abs(s[0].item() - (torch.log(torch.tensor(0.5)) + torch.log(torch.tensor(0.25))).item() / 3) < 1e-6
    assert s[1].item() == 0.0


class _FakeModel:
    """Stands in for UniGenB200 on the CPU: every image is a pure function of (philox key, global image index)."""
    device = torch.device("cpu")

    def t2i_generate(self, input_ids, batch_offset=0, philox_seed=None, philox_offset=0, **kw):
        n = input_ids.shape[0]
        idx = torch.arange(batch_offset, batch_offset + n)
        ids = ((idx[:, None] * 131 + torch.arange(6)[None, :] * 7 + philox_seed % 1000 + philox_offset % 89) % 8192).to(torch.int64)
        self.last_sel_prob = (0.1 + ((idx[:, None] * 37 + torch.arange(6)[None, :] + philox_seed % 11) % 50).float() / 100)
        self.seen_key = (philox_seed, philox_offset)
        return ids


def _bon_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_unigen_b200 import dist as ugd
    m = _FakeModel()
    # every rank hands in its OWN generator state: only rank 0's may decide the stream (sharding invariance)
    gen = torch.Generator().manual_seed(5 if rank == 0 else 999)
    ids, scores, top = ugd.best_of_n(m, torch.arange(10), 9, topk=4, generator=gen)
    ret[rank] = (ids.tolist(), scores.tolist(), top.tolist(), m.seen_key)
    dist.barrier()
    dist.destroy_process_group()


def test_best_of_n_is_invariant_to_the_number_of_ranks():
    from ml_unigen_b200 import dist as ugd
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bon_worker, args=(2, 31500 + (os.getpid() % 2000), ret), nprocs=2, join=True)
    assert ret[0] == ret[1]                                   # same gathered result and the same Philox key on both ranks
    m = _FakeModel()                                          # the single-process run with rank 0's generator
    ids, scores, top = ugd.best_of_n(m, torch.arange(10), 9, topk=4, generator=torch.Generator().manual_seed(5))
    assert (ids.tolist(), scores.tolist(), top.tolist(), m.seen_key) == ret[0]
