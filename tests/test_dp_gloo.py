"""Data-parallel host logic on CPU: world_size 2 over gloo.  A stub engine stands in for the GPU (tokens are a
deterministic function of the request), so what is tested is exactly the sharding + per-step all-gather code of
visualcla/dp.py: the DP result must equal the single-process result row for row."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "visual-chinese-llama-alpaca_b200")


class _StubEngine:
    """Mimics Engine's prefill/decode contract on the CPU: next token = f(request checksum, position)."""
    device = torch.device("cpu")

    def __init__(self):
        self.state = None
        self.calls = 0

    def vision_encode(self, px):
        self.vis = px.float().sum(dim=(1, 2, 3))

    def prefill(self, ids, mode, rows, all_logits=False, last_logits=False):
        self.state = (ids.sum(1) * 7 + (self.vis * 1000).long()) % 1009
        self.pos = 0
        first = (self.state % 997).to(torch.int32)
        self.hist = [first.clone()]
        return None, first, None

    def decode_step(self, tok_in, tok_out, logits=None, use_graph=True):
        self.calls += 1
        self.pos += 1
        tok_out.copy_(((tok_in.long() * 31 + self.state + self.pos) % 997).to(torch.int32))
        self.hist.append(tok_out.clone())

    def decode_many(self, tok, n_steps):
        for _ in range(n_steps):
            self.decode_step(tok, tok)

    def read_history(self, B, n_steps):
        return torch.stack(self.hist[:n_steps], 0)


class _StubModel:
    def __init__(self):
        self._engine = _StubEngine()

    def _image_layout(self, ids, px):
        return (1 if px is not None else 0), None


def _inputs(B):
    g = torch.Generator().manual_seed(3)
    return torch.randint(3, 900, (B, 6), generator=g), torch.randn(B, 3, 4, 4, generator=g)


def _worker(rank, world, port, B, n_new, q):
    sys.path.insert(0, PKG)
    from visualcla.dp import generate_dp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids, px = _inputs(B)
    m = _StubModel()
    out = generate_dp(m, ids, px, n_new)
    q.put((rank, out.clone(), m._engine.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5])
def test_dp_world2_equals_single(B):
    sys.path.insert(0, PKG)
    from visualcla.dp import generate_dp
    n_new = 6
    ids, px = _inputs(B)
    single = generate_dp(_StubModel(), ids, px, n_new)
    assert single.shape == (B, n_new)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, n_new, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, calls in res:
        assert torch.equal(out, single), f"rank {rank} result differs from the single-process run"
        assert calls == n_new - 1       # one decode call per step, one all-gather per step
