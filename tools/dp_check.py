"""Run under torchrun (one rank per GPU): data-parallel generation over NCCL must equal the single-GPU result row for row.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-chinese-llama-alpaca_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import visualcla  # noqa: E402
from visualcla.dp import generate_dp  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
from visualcla.engine import path_config_7b  # noqa: E402
cfg = dict(path_config_7b(), v_layers=1, r_layers=1, t_hidden=512, t_heads=4, t_ffn=1408, t_layers=2, t_vocab=2003)
B, n_new = 6, 12
m = visualcla.VisualCLAModel.from_synthetic(cfg, seed=9, max_batch=B, max_seq=160)
m.image_at_head = True
g = torch.Generator().manual_seed(5)
px = torch.randn(B, 3, cfg["v_image"], cfg["v_image"], generator=g)
ids = torch.randint(3, cfg["t_vocab"] - 4, (B, 20), generator=g)
ids[:, 0], ids[:, 1], ids[:, 2] = 1, cfg["t_vocab"] - 4, cfg["t_vocab"] - 3
calls = {"n": 0}
_orig = dist.all_gather_into_tensor


def _counting(*a, **k):
    calls["n"] += 1
    return _orig(*a, **k)


dist.all_gather_into_tensor = _counting
dp = generate_dp(m, ids, px, n_new)                     # every rank: its slice; the NCCL all-gather runs inside the decode graphs
dp2 = generate_dp(m, ids, px, n_new)                    # replay of the captured graphs
tiny = generate_dp(m, ids[:1], px[:1], n_new)           # fewer requests than ranks: the other ranks only take part in the exchange
in_graph = calls["n"] == 0 and m._engine.dp_width > 0   # no torch collective per token: the exchange belongs to the native context
single = m.generate(input_ids=ids.cuda(), pixel_values=px.cuda(), do_sample=False, max_new_tokens=n_new, eos_token_id=None, pad_token_id=0)
ok = torch.equal(dp, single) and torch.equal(dp2, single) and torch.equal(tiny, single[:1])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"[dp_check] world={world} B={B}: DP == single-GPU on every rank: {bool(flag.item())}; in-graph exchange: {in_graph}")
    print(dp.tolist())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
