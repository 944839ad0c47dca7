"""world_size-2 gloo test (CPU) of the only multi-GPU step on the path: rank-sharded encode followed by ONE
all-gather of the [B_local,32] ids.  The encode itself is stubbed (it needs a GPU); what is
tested is the host logic: int32 wire format, rank-major ordering, int64 at the boundary, every rank gets all ids."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist

    from models.seed_llama_tokenizer import SeedImageTokenMixin

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeImageTokenizer:
        def encode(self, x):            # deterministic ids derived from the shard content: [B,32] int64
            base = x.reshape(x.shape[0], -1)[:, 0].long()
            return (base[:, None] * 32 + torch.arange(32)[None]) % 8192

    class Tok(SeedImageTokenMixin):
        image_tokenizer = FakeImageTokenizer()

    B_local = 3
    shard = torch.arange(rank * B_local, (rank + 1) * B_local, dtype=torch.float32).reshape(B_local, 1, 1, 1)
    ids = Tok().encode_image_sharded(shard)
    q.put((rank, ids.dtype == torch.int64, tuple(ids.shape), ids.tolist()))
    dist.destroy_process_group()


def test_all_gather_of_ids_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = ((torch.arange(6)[:, None] * 32 + torch.arange(32)[None]) % 8192).tolist()
    for rank, is64, shape, ids in results:
        assert is64 and shape == (6, 32)
        assert ids == expect            # rank-major order, identical on every rank


def test_single_process_passthrough():
    sys.path.insert(0, REPO)
    from models.seed_llama_tokenizer import SeedImageTokenMixin

    class FakeImageTokenizer:
        def encode(self, x):
            return torch.zeros((x.shape[0], 32), dtype=torch.int64)

    class Tok(SeedImageTokenMixin):
        image_tokenizer = FakeImageTokenizer()

    assert tuple(Tok().encode_image_sharded(torch.zeros(4, 3, 2, 2)).shape) == (4, 32)
