"""Short run of the whole path for `ncu` (tools/ncu_all.sh): every kernel at its REAL VisualCLA-7B width, but a shallow stack
(2 ViT layers, 1 Resampler layer, 2 LLaMA layers) so that one `ncu --set full` pass over every launch stays within minutes.
Kernel shapes, tiles, split-K factors and grids depend on widths and batch, not on depth.

    python tools/profile_step.py <batch> [--sample]     # eager launches (no CUDA graph): vision + prefill + 3 decode steps
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "visual-chinese-llama-alpaca_b200"))
import torch  # noqa: E402
import visualcla  # noqa: E402
from visualcla import _native as N  # noqa: E402
from visualcla.engine import path_config_7b  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(os.environ.get("VCLA_PROFILE_T", "64"))
N.load().vcla_set_pdl(int(os.environ.get("VCLA_PDL", "1")))
cfg = dict(path_config_7b(), v_layers=2, r_layers=1, t_layers=2)
m = visualcla.VisualCLAModel.from_synthetic(cfg, seed=0, max_batch=B, max_seq=T + 64 + 64, max_prefill_tokens=B * (T + 64))
m.image_at_head = True
eng = m._engine
g = torch.Generator().manual_seed(1)
px = torch.randn(B, 3, 224, 224, generator=g).half().cuda()
ids = torch.randint(3, 49954, (B, T), generator=g).cuda()
if "--sample" in sys.argv:
    eng.set_sampler(eng.sampler_spec(do_sample=True, repetition_penalty=1.1, no_repeat_ngram_size=15, temperature=0.5, top_k=40, top_p=0.9, seed=1))
eng.vision_encode(px)
_, tok, _ = eng.prefill(ids, N.IMAGE_AT_HEAD, None, all_logits=False, last_logits=False)
tok = tok.clone()
for _ in range(3):
    eng.decode_step(tok, tok, None, use_graph=False)
torch.cuda.synchronize()
print("profile_step done: B", B, "tokens", tok.tolist()[:4])
