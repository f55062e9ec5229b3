#!/usr/bin/env python
"""SURVEY.md section 8f rank 1 at real size: a 32-layer Mistral-7B-shaped HF decoder (random weights, quantised on the GPU
without an fp32 host copy) goes through the reference's construction sequence (/root/reference/cullavo/load_cullavo.py:
73-126: NF4 double-quant Linear4bit everywhere but lm_head, prepare_model_for_kbit_training with NON-reentrant gradient
checkpointing, LoRA r=64 alpha=16 dropout=0.05 on all 7 projections, fp32 -> bf16 sweep) and trains for a few steps on
batch 8 x seq 2048 the way /root/reference/cullavo/arch_cullavo.py:638-647 drives `self.language_model(...)`, with
GradSync buckets + FusedLoraAdamW (clip 1.0).  Prints ONE JSON line: model tokens/s and the share of the step spent inside
libb2q's five hot-path entry points (CUDA events around every call).

    python tools/hf_real_size.py [--layers 32] [--batch 8] [--seq 2048] [--steps 3] [--warmup 2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--attn", default="sdpa")
    args = ap.parse_args()

    import torch
    import torch.nn as nn
    from importlib import import_module
    from transformers import LlamaConfig, LlamaForCausalLM

    import b200qlora as q

    lora = import_module("causal-unified-language-vision_b200.lora")
    stackmod = import_module("causal-unified-language-vision_b200.stack")
    par = import_module("causal-unified-language-vision_b200.parallel")
    optim = import_module("causal-unified-language-vision_b200.optim")
    F = q.functional
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    t0 = time.perf_counter()
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=args.layers, num_attention_heads=32,
                      num_key_value_heads=8, vocab_size=32000, max_position_embeddings=max(4096, args.seq),
                      rms_norm_eps=1e-5, attn_implementation=args.attn)
    with torch.device("meta"):
        model = LlamaForCausalLM(cfg)
    # every nn.Linear but lm_head -> Linear4bit with a weight drawn and NF4-quantised ON the GPU (load_cullavo.py:73-86)
    gen = torch.Generator(device=dev).manual_seed(0)
    n_lin = 0
    for name, mod in list(model.named_modules()):
        if isinstance(mod, nn.Linear) and not name.endswith("lm_head"):
            parent = model.get_submodule(name.rpartition(".")[0])
            setattr(parent, name.rpartition(".")[2], stackmod.make_quantized_linear(mod.out_features, mod.in_features, dev, gen))
            n_lin += 1
    # the rest (embeddings, norms, lm_head) materialised in bf16 on the GPU
    for name, mod in model.named_modules():
        for pn, p in list(mod.named_parameters(recurse=False)):
            if p.is_meta:
                w = torch.empty(p.shape, dtype=torch.bfloat16, device=dev)
                if p.dim() >= 2:
                    w.normal_(0.0, 0.02, generator=gen)
                else:
                    w.fill_(1.0)
                setattr(mod, pn, nn.Parameter(w, requires_grad=False))
        for bn, b in list(mod.named_buffers(recurse=False)):
            if b.is_meta:
                setattr(mod, bn, None)   # re-created below
    # rotary buffers live on meta after the meta construction: rebuild the rotary module on the device
    rot_cls = type(model.model.rotary_emb)
    model.model.rotary_emb = rot_cls(config=cfg, device=dev)
    lora.prepare_model_for_kbit_training(model, use_gradient_checkpointing=True,
                                         gradient_checkpointing_kwargs={"use_reentrant": False})      # :91-93
    names = lora.find_all_linear_names(model)                                                          # :8-20
    lora.add_adapter(model, lora.LoraConfig(r=64, lora_alpha=16, target_modules=names, lora_dropout=0.05, bias="none",
                                            task_type="CAUSAL_LM"), adapter_name="step1")             # :94-112
    for p in model.parameters():                                                                       # :124-126
        if p.dtype == torch.float32:
            p.data = p.data.to(torch.bfloat16)
    mods = [m for m in model.modules() if isinstance(m, lora.LoraLinear4bit)]
    with torch.no_grad():
        for m in mods:
            m.lora_B["step1"].weight.normal_(0.0, 0.02)
    sync = par.GradSync(mods, "step1")
    opt = optim.FusedLoraAdamW(sync, lr=2e-5, weight_decay=0.0)
    model.train()
    print(f"[hf] built: {n_lin} Linear4bit, {len(mods)} LoRA modules, {torch.cuda.memory_allocated() / 2**30:.1f} GiB, "
          f"{time.perf_counter() - t0:.1f}s", file=sys.stderr, flush=True)

    # CUDA events around every hot-path entry point
    events, on = [], [False]

    def timed(fn):
        def w(*a, **k):
            if not on[0]:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            events.append((e0, e1))
            return out
        return w

    auto = import_module("causal-unified-language-vision_b200.autograd")
    for name in ("lora_down", "qlora_fwd", "lora_bwd_du", "qlora_bwd_dx", "lora_grads"):
        setattr(F, name, timed(getattr(F, name)))
    assert auto.F is F

    ids = torch.randint(0, 32000, (args.batch, args.seq), device=dev, generator=gen)

    def step():
        sync.begin_step()
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        norm = opt.clip_grad_norm_(1.0)
        opt.step()
        opt.zero_grad()
        return loss, norm

    for i in range(args.warmup):
        loss, norm = step()
        torch.cuda.synchronize()
        print(f"[hf] warm-up step {i}: loss {float(loss):.4f} grad norm {float(norm):.4f}", file=sys.stderr, flush=True)
    launches0 = F.launch_count()
    on[0] = True
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        loss, norm = step()
    s1.record()
    torch.cuda.synchronize()
    on[0] = False
    ms = s0.elapsed_time(s1) / args.steps
    in_lib = sum(a.elapsed_time(b) for a, b in events) / args.steps
    toks = args.batch * args.seq
    print(json.dumps({
        "what": "HF LlamaForCausalLM, Mistral-7B shapes (32 x [q 4096, k/v 1024 (GQA), o 4096, gate/up 14336, down 4096]), NF4 "
                "double-quant base + LoRA r=64 dropout 0.05 on all projections, non-reentrant gradient checkpointing, "
                f"attention {args.attn}, GradSync buckets + FusedLoraAdamW, one GPU",
        "layers": args.layers, "tokens_per_step": toks, "ms_per_step": ms, "model_tokens_per_s": toks / (ms / 1e3),
        "ms_in_libb2q_per_step": in_lib, "share_in_libb2q": in_lib / ms, "libb2q_launches_per_step": (F.launch_count() - launches0) / args.steps,
        "loss": float(loss), "grad_norm": float(norm), "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2**30,
        "stall_records": q._lib.load().b2q_debug_stall_count(),
    }), flush=True)


if __name__ == "__main__":
    main()
