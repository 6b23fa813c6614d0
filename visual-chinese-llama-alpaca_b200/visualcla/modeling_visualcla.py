"""VisualCLAModel on the B200-native path.

Same public surface as the reference's composite model (ref: models/visualcla/modeling_visualcla.py:70-404):
`from_pretrained / from_merged_pretrained / from_vision_text_pretrained`, `forward`, `generate`,
`get/set_{input,output}_embeddings`, the attributes the reference's callers read (`.tokenizer`,
`.image_processor`, `.num_patch`, `.image_at_head`, `.device`, `.config`, `.text_model`, `.vision_model`,
`.visual_resampler`, `.image_projection_layer`) and the nn.Module verbs they use (`.eval() .float() .half()
.to() .state_dict() .resize_token_embeddings()`).  Underneath there are no nn.Modules: all arithmetic runs in
hand-written sm_100a kernels behind the C ABI in include/vcla.h (see engine.py / _native.py).
"""
from __future__ import annotations

import copy
import glob
import json
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Union

import torch

from . import _native as N
from .configuration_visualcla import VisualCLAConfig
from .engine import Engine, path_config_7b


# --------------------------------------------------------------------------------------------------
# checkpoint shards
# --------------------------------------------------------------------------------------------------
def _iter_checkpoint(directory: str):
    """Yield (name, tensor) from every `pytorch_model*.bin` / `*.safetensors` shard of an HF-style directory."""
    files = sorted(glob.glob(os.path.join(directory, "pytorch_model*.bin"))) + \
        sorted(glob.glob(os.path.join(directory, "*.safetensors")))
    if not files:
        raise ValueError(f"no checkpoint shards (pytorch_model*.bin / *.safetensors) under {directory}")
    for f in files:
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(f, device="cpu")
        else:
            sd = torch.load(f, map_location="cpu", weights_only=True)
        for k, v in sd.items():
            yield k, v
        del sd


class _EmbeddingView:
    """What `get_input_embeddings()` / `get_output_embeddings()` return: `.weight` materialises the table."""

    def __init__(self, model: "VisualCLAModel", name: str):
        self._model, self._name = model, name

    @property
    def weight(self) -> torch.Tensor:
        return self._model._engine.read_weight(self._name).to(self._model.device)

    def __call__(self, input_ids: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.embedding(input_ids.to(self._model.device), self.weight)


class _SubModel:
    """Lightweight handle standing in for the reference's nn.Module children (only what callers touch)."""

    def __init__(self, model: "VisualCLAModel", prefix: str, config):
        self._model, self._prefix = model, prefix
        self.config = config

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k[len(self._prefix):]: v for k, v in self._model.state_dict().items() if k.startswith(self._prefix)}

    def get_input_embeddings(self):
        return self._model.get_input_embeddings()

    def get_output_embeddings(self):
        return self._model.get_output_embeddings()


class VisualCLAModel:
    config_class = VisualCLAConfig
    base_model_prefix = "visualcla"

    def __init__(self, config: VisualCLAConfig = None, vision_model=None, text_model=None, device=None,
                 max_batch: int = 8, max_seq: int = 1024, max_prefill_tokens: Optional[int] = None,
                 torch_dtype=torch.bfloat16):
        if config is None:
            raise ValueError("VisualCLAModel needs a VisualCLAConfig")
        if vision_model is not None or text_model is not None:
            raise NotImplementedError("pre-built nn.Module sub-models are not used on the B200 path; load weights with "
                                      "from_merged_pretrained / from_vision_text_pretrained / load_state_dict")
        self.config = config
        self._engine = Engine(config.to_path_config(), max_batch=max_batch, max_seq=max_seq,
                              max_prefill_tokens=max_prefill_tokens, device=device)
        self.dtype = torch.bfloat16      # compute dtype of the path (bf16 operands, fp32 accumulate / residual stream)
        self.requested_dtype = torch_dtype
        self.image_at_head = True        # constructor default of the reference (:108); the loader flips it (:134)
        self.tokenizer = None
        self.image_processor = None
        self.num_patch = self._engine.nq
        self.vision_embed_dim = config.vision_config["hidden_size"]
        self.text_embed_dim = config.text_config["hidden_size"]
        self.text_model = _SubModel(self, "text_model.", SimpleNamespace(**config.text_config))
        self.vision_model = _SubModel(self, "vision_model.", SimpleNamespace(**config.vision_config))
        self.visual_resampler = _SubModel(self, "visual_resampler.", SimpleNamespace(**(config.visual_resampler_config or {})))
        self.image_projection_layer = _SubModel(self, "image_projection_layer.", None)
        self.generation_config = None
        self._tok_buf = {}

    # ---- nn.Module-ish verbs used by the reference's scripts ------------------------------------
    @property
    def device(self) -> torch.device:
        return self._engine.device

    def eval(self): return self
    def float(self): return self
    def half(self): return self
    def bfloat16(self): return self
    def requires_grad_(self, *_a, **_k): return self
    def train(self, mode: bool = False):
        if mode:
            raise NotImplementedError("the B200 path is inference-only (the reference ships no training code)")
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise N.NativeError("the B200 path has no CPU fallback: model.to('cpu') is not supported")
        return self

    def parameters(self):
        return iter(())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {n: self._engine.read_weight(n) for n, _s, _k in self._engine.weight_table()}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        loaded, unexpected = self._engine.load_state_dict(sd)
        expected = {n for n, _s, _k in self._engine.weight_table()}
        missing = sorted(expected - loaded)
        # tolerated extras: the dead pooler (ref: modeling_visual_resampler.py:725) and position_ids buffers
        unexpected = [k for k in unexpected if "pooler" not in k and "position_ids" not in k and "rotary_emb.inv_freq" not in k]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]}... ({len(missing)}), unexpected {unexpected[:5]}... ({len(unexpected)})")
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    def get_input_embeddings(self): return _EmbeddingView(self, "text_model.model.embed_tokens.weight")
    def get_output_embeddings(self): return _EmbeddingView(self, "text_model.lm_head.weight")
    def set_input_embeddings(self, new): self._engine.load_weight("text_model.model.embed_tokens.weight", new.weight)
    def set_output_embeddings(self, new): self._engine.load_weight("text_model.lm_head.weight", new.weight)

    def resize_token_embeddings(self, new_num_tokens: Optional[int] = None):
        """ref: scripts/inference/inference.py:69 -- grow the vocab (49954 -> 49958) before LoRA weights are applied."""
        old = self._engine.vocab
        if new_num_tokens is None or new_num_tokens == old:
            return self.get_input_embeddings()
        sd = self.state_dict()
        cfg = copy.deepcopy(self.config)
        cfg.text_config["vocab_size"] = int(new_num_tokens)
        e = self._engine
        new = Engine(cfg.to_path_config(), max_batch=e.max_batch, max_seq=e.max_seq, max_prefill_tokens=e.max_prefill_tokens, device=e.device)
        for k in ("text_model.model.embed_tokens.weight", "text_model.lm_head.weight"):
            w = sd.pop(k).float()
            grown = torch.zeros(new_num_tokens, w.shape[1])
            n = min(old, new_num_tokens)
            grown[:n] = w[:n]
            if new_num_tokens > old:
                grown[old:] = w[:old].mean(0, keepdim=True)     # HF initialises new rows from the mean embedding
            new.load_weight(k, grown)
        new.load_state_dict(sd)
        e.close()
        self._engine, self.config = new, cfg
        self.text_model.config = SimpleNamespace(**cfg.text_config)
        self._tok_buf = {}
        return self.get_input_embeddings()

    # ---- constructors -----------------------------------------------------------------------------
    @classmethod
    def from_synthetic(cls, path_cfg: Union[str, Dict] = "7b", seed: int = 0, **kw) -> "VisualCLAModel":
        """Random-init weights of the given architecture, generated on the device (no checkpoint files offline)."""
        p = path_config_7b() if path_cfg == "7b" else dict(path_cfg)
        model = cls(VisualCLAConfig.from_path_config(p), **kw)
        model._engine.init_synthetic(seed)
        return model

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, *args, **kwargs) -> "VisualCLAModel":
        """ref: modeling_visualcla.py:110-118.  Accepts a merged directory (text_encoder/ + vision_encoder/ +
        pytorch_model.bin) or a flat HF directory holding the whole composite state dict."""
        path = pretrained_model_name_or_path
        if not os.path.isdir(path):
            raise ValueError(f"{path} is not a local directory (no network access on this path)")
        kw = dict(torch_dtype=kwargs.pop("torch_dtype", torch.bfloat16), default_device=kwargs.pop("default_device", None),
                  device_map=kwargs.pop("device_map", None), load_in_8bit=kwargs.pop("load_in_8bit", False))
        kwargs.pop("_fast_init", None)
        if os.path.isdir(os.path.join(path, "text_encoder")):
            return cls.from_merged_pretrained(path, **kw, **kwargs)
        config = VisualCLAConfig.from_pretrained(path)
        model = cls(config, device=kw["default_device"], torch_dtype=kw["torch_dtype"], **kwargs)
        model.load_state_dict(dict(_iter_checkpoint(path)))
        return model

    @classmethod
    def from_merged_pretrained(cls, visualcla_model_name_or_path: str = None, *args, **kwargs) -> "VisualCLAModel":
        """ref: modeling_visualcla.py:120-181.  The four kwargs are REQUIRED there (popped without default); same here."""
        path = visualcla_model_name_or_path
        if not os.path.isdir(path):
            raise ValueError(f"{path} is not a local directory (the reference's hub path is broken too, SURVEY 3.6)")
        torch_dtype = kwargs.pop("torch_dtype")
        default_device = kwargs.pop("default_device")
        _device_map = kwargs.pop("device_map")          # whole model lives on one GPU (14.5 GB of 180 GB); DP replicates it
        load_in_8bit = kwargs.pop("load_in_8bit")
        if load_in_8bit:
            raise NotImplementedError("load_in_8bit (bitsandbytes) is out of scope of the B200 path (bf16 weights)")
        config = VisualCLAConfig.from_pretrained(path)
        text_dir, vision_dir = os.path.join(path, "text_encoder"), os.path.join(path, "vision_encoder")
        with open(os.path.join(text_dir, "config.json")) as f:
            config.text_config = json.load(f)
        with open(os.path.join(vision_dir, "config.json")) as f:
            vc = json.load(f)
            config.vision_config = vc.get("vision_config", vc) if "hidden_size" not in vc else vc
        model = cls(config, device=default_device, torch_dtype=torch_dtype, **kwargs)
        eng = model._engine
        seen = set()
        for k, v in _iter_checkpoint(text_dir):
            if "rotary_emb.inv_freq" in k:
                continue
            eng.load_weight("text_model." + k, v); seen.add("text_model." + k)
        for k, v in _iter_checkpoint(vision_dir):
            if "position_ids" in k:
                continue
            eng.load_weight("vision_model." + k, v); seen.add("vision_model." + k)
        for k, v in _iter_checkpoint(path):
            if k.startswith("visual_resampler.") and "pooler" not in k or k.startswith("image_projection_layer."):
                eng.load_weight(k, v); seen.add(k)
        missing = sorted({n for n, _s, _k in eng.weight_table()} - seen)
        if missing:
            raise RuntimeError(f"merged checkpoint {path} lacks {len(missing)} tensors, e.g. {missing[:4]}")
        return model

    @classmethod
    def from_vision_text_pretrained(cls, vision_model_name_or_path: str = None, text_model_name_or_path: str = None,
                                    visualcla_config: Union[str, VisualCLAConfig] = None, torch_dtype=torch.float16,
                                    default_device=None, device_map=None, load_in_8bit=False, **kwargs) -> "VisualCLAModel":
        """ref: modeling_visualcla.py:183-261: base CLIP + base LLaMA, resampler/projector randomly initialised
        (the LoRA path then overwrites them; LoRA folding itself is a 'next' row)."""
        if vision_model_name_or_path is None:
            raise ValueError("If `vision_model` is not defined as an argument, a `vision_model_name_or_path` has to be defined")
        if text_model_name_or_path is None:
            raise ValueError("If `text_model` is not defined as an argument, a `text_model_name_or_path` has to be defined")
        if load_in_8bit:
            raise NotImplementedError("load_in_8bit is out of scope of the B200 path")
        if isinstance(visualcla_config, str):
            visualcla_config = VisualCLAConfig.from_pretrained(visualcla_config)
        config = copy.deepcopy(visualcla_config)
        with open(os.path.join(text_model_name_or_path, "config.json")) as f:
            config.text_config = json.load(f)
        with open(os.path.join(vision_model_name_or_path, "config.json")) as f:
            vc = json.load(f)
            config.vision_config = vc.get("vision_config", vc) if "hidden_size" not in vc else vc
        model = cls(config, device=default_device, torch_dtype=torch_dtype, **kwargs)
        eng = model._engine
        eng.init_synthetic(0)   # resampler + projector: fresh init, as in the reference
        known = {n for n, _s, _k in eng.weight_table()}
        seen = set()
        for k, v in _iter_checkpoint(text_model_name_or_path):
            if "rotary_emb.inv_freq" not in k:
                eng.load_weight("text_model." + k, v); seen.add("text_model." + k)
        for k, v in _iter_checkpoint(vision_model_name_or_path):
            # a full CLIPModel checkpoint also carries the text tower / projections: only the vision tower is on the path
            name = "vision_model." + k
            if name in known:
                eng.load_weight(name, v); seen.add(name)
        # only the resampler and the projector may keep their fresh initialisation: a missing shard or a renamed key must
        # not leave LLaMA / CLIP layers at random weights
        missing = sorted(n for n in known - seen if n.startswith(("text_model.", "vision_model.")))
        if missing:
            raise RuntimeError(f"base checkpoints lack {len(missing)} tensors of the text/vision towers, e.g. {missing[:4]}")
        return model

    def save_merged_pretrained(self, output_dir: str):
        """Write the merged-directory layout of ref: scripts/merge_llama_with_visualcla_lora.py:87-97."""
        os.makedirs(os.path.join(output_dir, "text_encoder"), exist_ok=True)
        os.makedirs(os.path.join(output_dir, "vision_encoder"), exist_ok=True)
        sd = self.state_dict()
        text = {k[len("text_model."):]: v for k, v in sd.items() if k.startswith("text_model.")}
        vision = {k[len("vision_model."):]: v for k, v in sd.items() if k.startswith("vision_model.")}
        rest = {k: v for k, v in sd.items() if not k.startswith(("text_model.", "vision_model."))}
        torch.save(text, os.path.join(output_dir, "text_encoder", "pytorch_model.bin"))
        torch.save(vision, os.path.join(output_dir, "vision_encoder", "pytorch_model.bin"))
        torch.save(rest, os.path.join(output_dir, "pytorch_model.bin"))
        with open(os.path.join(output_dir, "text_encoder", "config.json"), "w") as f:
            json.dump(self.config.text_config, f)
        with open(os.path.join(output_dir, "vision_encoder", "config.json"), "w") as f:
            json.dump(self.config.vision_config, f)
        self.config.save_pretrained(output_dir)

    @torch.no_grad()
    def embed_images(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """pixels (N,3,I,I) -> image embeddings (N, num_query_tokens, text_hidden): the tg-webui pipeline's entry point
        (ref: scripts/inference/text_generation_webui/visualcla/visualcla.py:116-129)."""
        return self._engine.vision_encode(pixel_values, return_embeds=True)

    # ---- prompt assembly (host side of ref :290-312 / :356-377) -----------------------------------
    def _image_layout(self, input_ids: torch.Tensor, pixel_values):
        """-> (image_mode, img_rows or None).  Replicates the reference's per-sample checks for the placeholder layout."""
        if pixel_values is None:
            return N.TEXT_ONLY, None
        if self.image_at_head:
            return N.IMAGE_AT_HEAD, None
        tok = self.tokenizer
        if tok is None:
            raise AttributeError("model.tokenizer (img_start_token_id / img_end_token_id) is required when image_at_head is False")
        ids = input_ids.detach().cpu()
        nq = self._engine.nq
        rows = []
        for cur in ids:
            pos = torch.where(cur == tok.img_start_token_id)[0]
            if len(pos) == 0:
                rows.append(-1)
                continue
            p = int(pos[0])
            if p + nq + 1 >= cur.shape[0] or int(cur[p + nq + 1]) != tok.img_end_token_id:
                raise ValueError(f"Num of patch ({nq}) is not equal to the length of pre-filled image patch tokens.")
            rows.append(p + 1)
        return N.IMAGE_PLACEHOLDER, torch.tensor(rows, dtype=torch.int32)

    def _left_pad(self, attention_mask, image_mode):
        """attention_mask -> per-sequence left-padding counts (or None).  HF batches prompts of different lengths by LEFT
        padding; any other mask shape is rejected."""
        if attention_mask is None or bool((attention_mask != 0).all()):
            return None
        m = (attention_mask != 0).to("cpu")
        pads = (~m).sum(1)
        T = m.shape[1]
        expect = torch.arange(T)[None, :] >= pads[:, None]
        if not torch.equal(m, expect):
            raise NotImplementedError("only left padding (attention_mask = [0]*p + [1]*(T-p)) is supported on the B200 path")
        if bool((pads >= T).any()):
            raise ValueError("attention_mask masks a whole sequence")
        if image_mode == N.IMAGE_AT_HEAD:
            raise NotImplementedError("padded batches need the placeholder layout (image_at_head=False): the reference's at-head splice "
                                      "ignores the mask (ref modeling_visualcla.py:291)")
        return pads.to(torch.int32)

    # ---- forward: logits for every position (ref :264-330) ---------------------------------------
    @torch.no_grad()
    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, position_ids=None, past_key_values=None,
                labels=None, use_cache=None, return_loss=None, return_dict=None, **kwargs):
        from transformers.modeling_outputs import CausalLMOutputWithPast
        if past_key_values is not None:
            raise NotImplementedError("forward(past_key_values=...) is not supported; use generate()")
        mode, rows = self._image_layout(input_ids, pixel_values)
        pads = self._left_pad(attention_mask, mode)
        eng = self._engine
        if mode != N.TEXT_ONLY:
            if mode == N.IMAGE_AT_HEAD and labels is None:
                # ref quirk (:313-315): labels[:, [0]] is indexed unconditionally in this layout
                raise TypeError("'NoneType' object is not subscriptable (labels are required with image_at_head=True)")
            eng.vision_encode(pixel_values)
        _, _, logits = eng.prefill(input_ids, mode, rows, all_logits=True, last_logits=False, left_pad=pads, pos_from_mask=False)
        loss = None
        if labels is not None:
            lab = labels.to(logits.device)
            if mode == N.IMAGE_AT_HEAD:
                fill = torch.full((lab.shape[0], eng.nq), -100, dtype=lab.dtype, device=lab.device)
                lab = torch.cat([lab[:, :1], fill, lab[:, 1:]], dim=1)
            loss = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), lab[:, 1:].reshape(-1), ignore_index=-100)
        if return_dict is False:
            return ((loss,) if loss is not None else ()) + (logits,)
        return CausalLMOutputWithPast(loss=loss, logits=logits)

    __call__ = forward

    # ---- generate: only the NEW tokens are returned (ref :333-392) --------------------
    def _resolve_generation_config(self, generation_config, kwargs):
        from transformers import GenerationConfig
        gc = copy.deepcopy(generation_config) if generation_config is not None else GenerationConfig()
        unused = gc.update(**kwargs)
        for k in list(unused):
            if k in ("output_scores", "output_logits", "return_dict_in_generate", "use_cache", "tfs", "top_a", "mirostat_mode", "mirostat_tau", "mirostat_eta"):
                setattr(gc, k, unused.pop(k))
        if unused:
            raise ValueError(f"generate(): unsupported arguments {sorted(unused)}")
        if getattr(gc, "num_beams", 1) not in (None, 1) or getattr(gc, "num_return_sequences", 1) not in (None, 1):
            raise NotImplementedError("beam search / num_return_sequences > 1 are not on the B200 path")
        return gc

    @staticmethod
    def _eos_set(gc):
        e = gc.eos_token_id
        if e is None:
            return []
        return [int(x) for x in (e if isinstance(e, (list, tuple)) else [e])]

    @torch.no_grad()
    def generate(self, input_ids=None, pixel_values=None, attention_mask=None, generation_config=None,
                 logits_processor=None, stopping_criteria=None, prefix_allowed_tokens_fn=None, synced_gpus=False, **kwargs):
        gc = self._resolve_generation_config(generation_config, kwargs)
        if prefix_allowed_tokens_fn is not None:
            raise NotImplementedError("prefix_allowed_tokens_fn is not supported on the B200 path")
        eng = self._engine
        B = input_ids.shape[0]
        if B > eng.max_batch:
            outs = []
            for s in range(0, B, eng.max_batch):
                sl = slice(s, s + eng.max_batch)
                outs.append(self.generate(input_ids[sl], None if pixel_values is None else pixel_values[sl],
                                          None if attention_mask is None else attention_mask[sl],
                                          gc, logits_processor, stopping_criteria, None, synced_gpus))
            width = max(o.shape[1] for o in outs)
            pad = gc.pad_token_id if gc.pad_token_id is not None else 0
            outs = [torch.nn.functional.pad(o, (0, width - o.shape[1]), value=pad) for o in outs]
            return torch.cat(outs, 0)

        mode, rows = self._image_layout(input_ids, pixel_values)
        pads = self._left_pad(attention_mask, mode)
        S = input_ids.shape[1] + (eng.nq if mode == N.IMAGE_AT_HEAD else 0)
        max_new = gc.max_new_tokens if gc.max_new_tokens is not None else max(int(gc.max_length or 20), 1)
        min_new = int(getattr(gc, "min_new_tokens", 0) or 0)
        if S + max_new > eng.max_seq:
            raise ValueError(f"prompt ({S}) + max_new_tokens ({max_new}) exceeds the context capacity max_seq={eng.max_seq} "
                             f"of this model instance")
        eos = self._eos_set(gc)
        pad = gc.pad_token_id if gc.pad_token_id is not None else (eos[0] if eos else 0)
        processors = self._build_processors(gc, logits_processor)
        sampling = bool(gc.do_sample)
        need_logits = sampling or len(processors) > 0 or bool(getattr(gc, "output_logits", False) or getattr(gc, "output_scores", False))
        crit = list(stopping_criteria) if stopping_criteria is not None else []

        if mode != N.TEXT_ONLY:
            eng.vision_encode(pixel_values)
        dev = eng.device
        key = (B, need_logits)
        if key not in self._tok_buf:
            # persistent step buffers (their addresses key the captured CUDA graph).  chat() runs under torch.inference_mode and
            # chat_in_stream's worker thread does not: allocate them as ordinary tensors so both may update them in place.
            with torch.inference_mode(False):
                self._tok_buf[key] = (torch.zeros(B, dtype=torch.int32, device=dev),
                                      torch.empty(B, eng.vocab, dtype=torch.float32, device=dev) if need_logits else None)
        tok, logits = self._tok_buf[key]
        out = torch.full((B, max_new), pad, dtype=torch.int64, device=dev)
        all_logits: List[torch.Tensor] = []

        spec = self._device_sampler_spec(gc, eos, pad, min_new, logits_processor, crit, processors)
        if spec is not None:
            return self._generate_on_device(spec, input_ids, mode, rows, pads, B, max_new, eos, pad, gc, tok)
        last, first_tok, _ = eng.prefill(input_ids, mode, rows, all_logits=False, last_logits=need_logits, left_pad=pads, pos_from_mask=True)
        if not need_logits and not eos and not crit:
            # pure greedy, fixed length: graph replays only; tokens come from the device-side history the graph appends to
            tok.copy_(first_tok)
            eng.decode_many(tok, max_new - 1)
            result = eng.read_history(B, max_new).t().to(torch.int64)
            if getattr(gc, "return_dict_in_generate", False):
                return SimpleNamespace(sequences=result, logits=None, scores=None)
            return result
        finished = torch.zeros(B, dtype=torch.bool, device=dev)
        n_done = 0
        for step in range(max_new):
            if step == 0:
                cur_logits, greedy = last, first_tok
            else:
                eng.decode_step(tok, tok, logits)
                cur_logits, greedy = logits, tok
            if need_logits:
                if getattr(gc, "output_logits", False):
                    all_logits.append(cur_logits.clone())
                scores = cur_logits
                if processors:
                    scores = cur_logits.clone()
                    if eos and step < min_new:
                        scores[:, eos] = -float("inf")
                    hist = out[:, :step]
                    for p in processors:
                        scores = p(hist, scores)
                if sampling:
                    probs = torch.softmax(scores.float(), dim=-1)
                    nxt = torch.multinomial(probs, num_samples=1).squeeze(1)
                else:
                    nxt = scores.argmax(dim=-1)
                nxt = nxt.to(torch.int32)
            else:
                nxt = greedy
            if eos:
                nxt = torch.where(finished, torch.full_like(nxt, pad), nxt)
            out[:, step] = nxt
            if nxt.data_ptr() != tok.data_ptr():
                tok.copy_(nxt)
            n_done = step + 1
            stop = False
            if eos:
                is_eos = torch.zeros_like(finished)
                for e in eos:
                    is_eos |= nxt == e
                finished |= is_eos
                # host poll (one sync) only every 8 steps: sequences are independent, so running a few extra
                # steps past the last EOS cannot change any retained token (HF syncs every step)
                if (step & 7) == 7 or step == max_new - 1:
                    stop = bool(finished.all())
            for cfn in crit:
                r = cfn(out[:, :n_done], cur_logits if need_logits else None)
                if isinstance(r, torch.Tensor):
                    r = bool(r.all())
                stop = stop or bool(r)
            if stop:
                break
        result = out[:, :n_done]
        if eos:
            # cut at the step where the last sequence finished (what HF's per-step check would have produced)
            hit = torch.zeros(B, n_done, dtype=torch.bool, device=dev)
            for e in eos:
                hit |= result == e
            first = torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full((B,), n_done, device=dev))
            keep = int(first.max())
            result = result[:, :keep]
            idx = torch.arange(keep, device=dev)[None, :]
            result = torch.where(idx < first[:, None], result, torch.full_like(result, pad))
        if getattr(gc, "return_dict_in_generate", False):
            return SimpleNamespace(sequences=result, logits=tuple(all_logits) if all_logits else None, scores=None)
        return result

    # ---- sampling / EOS on the device: one fused kernel per step inside the decode graph ---------------------
    def _device_sampler_spec(self, gc, eos, pad, min_new, extra_processors, crit, processors):
        """-> a native sampler spec when this call can run entirely on the device, else None (host logits-processor path).
        On the device: greedy or sampling with repetition_penalty / no_repeat_ngram_size / temperature / top_k (1..1024) / top_p and
        up to 4 EOS ids -- the reference's DEFAULT_GENERATION_CONFIG (ref modeling_utils.py:36-47) is such a call.  Anything else
        (custom processors, stopping criteria / streaming, TFS / Top-A / Mirostat, top_k disabled, logits or scores requested)
        keeps the per-step host path."""
        eng = self._engine
        if os.environ.get("VCLA_HOST_SAMPLER") == "1" or not hasattr(eng, "sampler_supported") or not eng.sampler_supported():
            return None
        if extra_processors or crit or getattr(gc, "output_logits", False) or getattr(gc, "output_scores", False) or len(eos) > 4:
            return None
        sampling = bool(gc.do_sample)
        rp = getattr(gc, "repetition_penalty", None) or 1.0
        ng = getattr(gc, "no_repeat_ngram_size", None) or 0
        if not sampling and rp == 1.0 and ng == 0 and not eos:
            return None                                   # plain fixed-length greedy: the argmax graphs
        if getattr(gc, "mirostat_mode", 0) == 2:
            return None
        for name, on in (("tfs", lambda v: 0.0 <= v < 1.0), ("top_a", lambda v: 0.0 < v <= 1.0), ("typical_p", lambda v: v < 1.0),
                         ("min_p", lambda v: v > 0.0), ("epsilon_cutoff", lambda v: v > 0.0), ("eta_cutoff", lambda v: v > 0.0)):
            v = getattr(gc, name, None)
            if v is not None and on(v):
                return None
        temperature = gc.temperature if (sampling and gc.temperature is not None) else 1.0
        top_k = gc.top_k if sampling else 0
        top_p = gc.top_p if (sampling and gc.top_p is not None) else 1.0
        if sampling and (top_k is None or not 1 <= top_k <= 1024 or temperature <= 0 or not 0.0 < top_p <= 1.0):
            return None
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # from torch's global generator: torch.manual_seed reproduces a run
        return eng.sampler_spec(do_sample=sampling, repetition_penalty=rp, no_repeat_ngram_size=ng, temperature=temperature, top_k=top_k or 0,
                                top_p=top_p, min_new_tokens=min_new, eos_token_id=eos, pad_token_id=pad, seed=seed)

    def _generate_on_device(self, spec, input_ids, mode, rows, pads, B, max_new, eos, pad, gc, tok):
        eng = self._engine
        eng.set_sampler(spec)
        try:
            _, first_tok, _ = eng.prefill(input_ids, mode, rows, all_logits=False, last_logits=False, left_pad=pads, pos_from_mask=True)
            tok.copy_(first_tok)
            n_done = 1
            if not eos:
                eng.decode_many(tok, max_new - 1)
                n_done = max_new
            else:
                # EOS: graphs of 8 steps, one host poll of the per-sequence finished flags between them (sequences are independent
                # and a finished one only emits pad, so running a few steps past the last EOS cannot change a retained token)
                while n_done < max_new:
                    if bool(eng.read_finished(B).bool().all()):
                        break
                    k = min(8, max_new - n_done)
                    eng.decode_many(tok, k)
                    n_done += k
            result = eng.read_history(B, n_done).t().to(torch.int64)
        finally:
            eng.set_sampler(None)
        if eos:
            hit = torch.zeros_like(result, dtype=torch.bool)
            for e in eos:
                hit |= result == e
            first = torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full((B,), n_done, device=result.device))
            result = result[:, : int(first.max())]        # cut where the last sequence finished (HF's per-step check)
        if getattr(gc, "return_dict_in_generate", False):
            return SimpleNamespace(sequences=result, logits=None, scores=None)
        return result

    @staticmethod
    def _build_processors(gc, extra):
        """HF logits processors for the sampling knobs of DEFAULT_GENERATION_CONFIG
        (ref: models/visualcla/modeling_utils.py:36-47).  Order follows HF's _get_logits_processor."""
        from transformers.generation import logits_process as lp
        procs = []
        rp = getattr(gc, "repetition_penalty", None)
        if rp is not None and rp != 1.0:
            procs.append(lp.RepetitionPenaltyLogitsProcessor(penalty=rp))
        ng = getattr(gc, "no_repeat_ngram_size", None)
        if ng is not None and ng > 0:
            procs.append(lp.NoRepeatNGramLogitsProcessor(ng))
        if extra:
            procs.extend(list(extra))
        if gc.do_sample:
            if gc.temperature is not None and gc.temperature != 1.0:
                procs.append(lp.TemperatureLogitsWarper(gc.temperature))
            if gc.top_k is not None and gc.top_k != 0:
                procs.append(lp.TopKLogitsWarper(top_k=gc.top_k, min_tokens_to_keep=1))
            if gc.top_p is not None and gc.top_p < 1.0:
                procs.append(lp.TopPLogitsWarper(top_p=gc.top_p, min_tokens_to_keep=1))
            from . import modeling_utils as mu
            if getattr(gc, "mirostat_mode", 0) == 2:
                # ref modeling_utils.py:366-371: Mirostat v2 joins the warpers and "disables samplers other than temperature" with a
                # remove-while-iterating loop -- which skips the element after every removal (temperature, top-k, top-p -> top-p
                # survives).  Replayed literally so the same generation config samples from the same distribution.
                n_fixed = len(procs) - sum(isinstance(p, (lp.TemperatureLogitsWarper, lp.TopKLogitsWarper, lp.TopPLogitsWarper)) for p in procs)
                warpers = procs[n_fixed:]
                for w in warpers:
                    if not isinstance(w, lp.TemperatureLogitsWarper):
                        warpers.remove(w)
                procs = procs[:n_fixed] + warpers
                procs.append(mu.MirostatLogitsWarper(mirostat_mode=2, mirostat_tau=getattr(gc, "mirostat_tau", 5), mirostat_eta=getattr(gc, "mirostat_eta", 0.1)))
                return procs
            for name, kw in (("tfs", "tfs"), ("top_a", "top_a")):
                val = getattr(gc, name, None)
                if val is not None and ((name == "tfs" and 0.0 <= val < 1.0) or (name == "top_a" and 0.0 < val <= 1.0)):
                    procs.append(mu.TailFreeLogitsWarper(tfs=val) if name == "tfs" else mu.TopALogitsWarper(top_a=val))
        return procs
