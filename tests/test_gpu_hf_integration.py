"""Drop-in test on a real HF decoder (pytest -m gpu): SURVEY.md section 8f rank 1.

A tiny ``LlamaForCausalLM`` goes through exactly the reference's construction sequence
(/root/reference/cullavo/load_cullavo.py:73-126): 4-bit replacement of every ``nn.Linear`` except ``lm_head``
-> ``prepare_model_for_kbit_training`` (non-reentrant gradient checkpointing) -> ``add_adapter(LoraConfig)`` ->
fp32->bf16 sweep, and is trained for one step.  A twin with the de-quantised weights in plain ``nn.Linear`` modules and
the LoRA branch written in stock PyTorch ops (no checkpointing) provides the expected loss and LoRA gradients.
"""
import copy
import importlib

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu


class TorchLora(nn.Module):
    """Stock-PyTorch statement of peft.tuners.lora.bnb.Linear4bit.forward (dropout off)."""

    def __init__(self, w_bf16, A, B, s):
        super().__init__()
        self.w = nn.Parameter(w_bf16, requires_grad=False)
        self.A = nn.Parameter(A.detach().clone())
        self.B = nn.Parameter(B.detach().clone())
        self.s = s

    def forward(self, x):
        return TF.linear(x, self.w) + TF.linear(TF.linear(x, self.A), self.B) * self.s


def _build(cuda_dev):
    from transformers import LlamaConfig, LlamaForCausalLM

    lora = importlib.import_module("causal-unified-language-vision_b200.lora")
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=256, attn_implementation="eager")
    torch.manual_seed(0)
    model = LlamaForCausalLM(cfg)
    # --- the reference's sequence ---------------------------------------------------------------
    lora.replace_with_4bit_linear(model, modules_to_not_convert=["lm_head"])          # load_cullavo.py:73-86
    model.to(cuda_dev)                                                                 # quantises on the move
    lora.prepare_model_for_kbit_training(model, use_gradient_checkpointing=True,
                                         gradient_checkpointing_kwargs={"use_reentrant": False})  # :91-93
    names = lora.find_all_linear_names(model)                                          # :8-20
    assert set(names) == {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"}
    lora.add_adapter(model, lora.LoraConfig(r=64, lora_alpha=16, target_modules=names, lora_dropout=0.0,
                                            bias="none", task_type="CAUSAL_LM"), adapter_name="step1")  # :103-112
    for p in model.parameters():                                                       # :124-126
        if p.dtype == torch.float32:
            p.data = p.data.to(torch.bfloat16)
    with torch.no_grad():  # B = 0 at init would make every LoRA gradient but dB vanish
        for m in model.modules():
            if isinstance(m, lora.LoraLinear4bit):
                m.lora_B["step1"].weight.normal_(0, 0.02)
    return model, lora


def test_hf_llama_one_training_step_matches_torch_twin(lib_built, cuda_dev):
    import b200qlora as q

    model, lora = _build(cuda_dev)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    assert sum(isinstance(m, lora.LoraLinear4bit) for m in model.modules()) == 14
    # twin: same modules, de-quantised weights, stock torch LoRA, no checkpointing
    twin = copy.deepcopy(model)
    twin.gradient_checkpointing_disable()
    pairs = []
    for name, mod in list(model.named_modules()):
        if isinstance(mod, lora.LoraLinear4bit):
            base = mod.base_layer
            w = q.dequantize_4bit(base.weight.data, base.weight.quant_state)
            t = TorchLora(w, mod.lora_A["step1"].weight, mod.lora_B["step1"].weight, mod.scaling["step1"])
            parent = twin.get_submodule(name.rpartition(".")[0])
            setattr(parent, name.rpartition(".")[2], t)
            pairs.append((name, mod, t))
    torch.manual_seed(1)
    ids = torch.randint(0, 512, (2, 96), device=cuda_dev)
    model.train()
    twin.train()
    out = model(input_ids=ids, labels=ids)
    out.loss.backward()
    ref = twin(input_ids=ids, labels=ids)
    ref.loss.backward()
    torch.cuda.synchronize()
    assert abs(float(out.loss.detach()) - float(ref.loss.detach())) <= 2e-2 * abs(float(ref.loss.detach()))
    rel = lambda a, b: float((a.detach().float() - b.detach().float()).abs().max() / b.detach().float().abs().max())
    assert rel(out.logits, ref.logits) <= 3e-2
    worst = 0.0
    for name, mod, t in pairs:
        ga, gb = mod.lora_A["step1"].weight.grad, mod.lora_B["step1"].weight.grad
        assert ga is not None and gb is not None, name
        assert mod.base_layer.weight.grad is None
        worst = max(worst, rel(ga, t.A.grad), rel(gb, t.B.grad))
    assert worst <= 5e-2, worst   # bf16 differences accumulate through two decoder layers of recompute


def test_second_adapter_step2_is_the_only_trainable_one(lib_built, cuda_dev):
    """add_adapter_for_step2 semantics (load_cullavo.py:23-59): step1 stays resident, step2 is active and trainable."""
    model, lora = _build(cuda_dev)
    lora.add_adapter(model, lora.LoraConfig(r=64, lora_alpha=16, target_modules=["q_proj", "v_proj"], lora_dropout=0.05,
                                            bias="none", task_type="CAUSAL_LM"), adapter_name="step2")
    for p in model.parameters():
        if p.dtype == torch.float32:
            p.data = p.data.to(torch.bfloat16)
    trainable = [n for n, p in model.named_parameters() if p.requires_grad]
    assert trainable and all(".step2." in n for n in trainable)
    ids = torch.randint(0, 512, (1, 40), device=cuda_dev)
    model.train()
    model(input_ids=ids, labels=ids).loss.backward()
    q0 = model.model.layers[0].self_attn.q_proj
    assert q0.lora_A["step2"].weight.grad is not None and q0.lora_A["step1"].weight.grad is None
    # modules without a step2 adapter (k_proj) fall back to the frozen base path
    k0 = model.model.layers[0].self_attn.k_proj
    assert "step2" not in k0.lora_A and k0.lora_A["step1"].weight.grad is None


def test_hf_llama_training_loop_with_gradsync_and_fused_adamw(lib_built, cuda_dev):
    """The whole replaced slice in a real decoder for a few optimizer steps: LoRA gradients written by the kernels into
    GradSync's flat buckets (no autograd accumulation), device-side global-norm clip + fused bf16 AdamW on the buckets,
    cosine LR schedule (trainer/cullavo_trainer.py:13-14, pipeline/CuLLaVOPipeline.py:87-92).  The loss on a fixed batch
    must go down and the LoRA weights must equal a torch.optim.AdamW twin fed with the same gradients (bf16 tolerance)."""
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    optim = importlib.import_module("causal-unified-language-vision_b200.optim")
    model, lora = _build(cuda_dev)
    mods = [m for m in model.modules() if isinstance(m, lora.LoraLinear4bit)]
    sync = par.GradSync(mods, "step1")
    opt = optim.FusedLoraAdamW(sync, lr=2e-3, weight_decay=0.0, state_dtype=torch.float32)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=6, eta_min=1e-5)
    # twin optimizer state in fp32 on copies of the LoRA weights
    twin_p = [p.detach().float().clone().requires_grad_(True) for (_, p, _) in sync.slots]
    twin = torch.optim.AdamW(twin_p, lr=2e-3, weight_decay=0.0)
    twin_sched = torch.optim.lr_scheduler.CosineAnnealingLR(twin, T_max=6, eta_min=1e-5)
    torch.manual_seed(3)
    ids = torch.randint(0, 512, (2, 64), device=cuda_dev)
    model.train()
    losses = []
    for step in range(6):
        sync.begin_step()
        loss = model(input_ids=ids, labels=ids).loss
        loss.backward()
        sync.finish()
        losses.append(float(loss.detach()))
        # the kernels wrote straight into the buckets: param.grad IS a bucket view
        g0 = mods[0].lora_A["step1"].weight.grad
        assert g0 is not None and float(g0.float().abs().sum()) > 0
        for tp, (_, p, _) in zip(twin_p, sync.slots):
            tp.grad = p.grad.detach().float().clone()
        total = torch.nn.utils.clip_grad_norm_(twin_p, 1.0)
        got_total = opt.clip_grad_norm_(1.0)
        assert abs(float(got_total) - float(total)) <= 2e-3 * float(total)
        opt.step()
        opt.zero_grad()
        sched.step()
        twin.step()
        twin_sched.step()
        # the bf16 parameter tracks the fp32 twin to bf16 resolution of the largest weight; re-sync the twin to the
        # rounded values so the comparison stays a per-step one
        for tp, (_, p, _) in zip(twin_p, sync.slots):
            err = float((p.detach().float() - tp.detach()).abs().max())
            assert err <= 2 ** -7 * float(tp.detach().abs().max()), (step, err)
            with torch.no_grad():
                tp.copy_(p.detach().float())
    assert losses[-1] < losses[0] - 0.05, losses


def test_non_lora_trainables_are_clipped_and_stepped_with_the_lora_buckets(lib_built, cuda_dev):
    """The reference trains `multi_modal_projector` next to the LoRA weights and clips the global norm over ALL trainable
    parameters (/root/reference/cullavo/load_cullavo.py:128-130, pipeline/CuLLaVOPipeline.py:90-91).  Here a projector in
    front of the decoder plays that role: registered as GradSync(extra_params=...), its gradient must enter the clip norm,
    be scaled by the same coefficient, and be stepped by FusedLoraAdamW's companion AdamW like a stock optimizer would."""
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    optim = importlib.import_module("causal-unified-language-vision_b200.optim")
    model, lora = _build(cuda_dev)
    torch.manual_seed(11)
    proj = nn.Linear(32, 256, bias=True).to(cuda_dev).to(torch.bfloat16)       # "multi_modal_projector": bf16 after the sweep
    mods = [m for m in model.modules() if isinstance(m, lora.LoraLinear4bit)]
    sync = par.GradSync(mods, "step1", extra_params=list(proj.parameters()))
    opt = optim.FusedLoraAdamW(sync, lr=1e-3, weight_decay=0.0, state_dtype=torch.float32)
    assert opt.extra_optimizer is not None and len(opt.extra_params) == 2
    twin_proj = copy.deepcopy(proj).float()
    twin_opt = torch.optim.AdamW(twin_proj.parameters(), lr=1e-3, weight_decay=0.0)
    feats = torch.randn(2, 24, 32, device=cuda_dev, dtype=torch.bfloat16)
    model.train()
    for step in range(3):
        sync.begin_step()
        emb = proj(feats)                                                       # [2, 24, 256] image-token embeddings
        out = model(inputs_embeds=emb, labels=torch.randint(0, 512, (2, 24), device=cuda_dev))
        out.loss.backward()
        # reference: ONE global norm over LoRA grads + projector grads
        lora_sq = sum(float(p.grad.float().pow(2).sum()) for (_, p, _) in sync.slots)
        extra_sq = sum(float(p.grad.float().pow(2).sum()) for p in proj.parameters())
        want = (lora_sq + extra_sq) ** 0.5
        assert extra_sq > 0
        max_norm = 0.5 * want                                                   # make the clip bite
        g_before = [p.grad.detach().float().clone() for p in proj.parameters()]
        got = float(opt.clip_grad_norm_(max_norm))
        assert abs(got - want) <= 5e-3 * want, (got, want)
        opt.step()
        coef = max_norm / (want + 1e-6)
        for tp, g in zip(twin_proj.parameters(), g_before):
            tp.grad = g * coef
        twin_opt.step()
        for p, tp in zip(proj.parameters(), twin_proj.parameters()):
            err = float((p.detach().float() - tp.detach()).abs().max())
            assert err <= 2 ** -7 * float(tp.detach().abs().max()) + 1e-6, (step, err)
            with torch.no_grad():
                tp.copy_(p.detach().float())
        opt.zero_grad()
        assert all(p.grad is None for p in proj.parameters())
