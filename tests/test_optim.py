"""AdamW + clip: the oracle is pinned against torch.optim.AdamW / clip_grad_norm_ (CPU); the fused kernel against the
oracle (GPU).  Host logic (flat parameter buckets) on CPU."""
import importlib
import math

import pytest
import torch
import torch.nn as nn


def test_oracle_matches_torch_adamw_and_clip():
    from oracle import adamw

    torch.manual_seed(0)
    p = nn.Parameter(torch.randn(257))
    ref = torch.optim.AdamW([p], lr=2e-5 * 500, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    po, m, v = p.detach().clone(), torch.zeros(257), torch.zeros(257)
    for step in range(1, 6):
        g = torch.randn(257) * 3.0
        p.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([p], 10.0)
        coef = adamw.clip_coef([g], 10.0)
        assert abs(coef - min(1.0, 10.0 / (float(total) + 1e-6))) < 1e-6
        ref.step()
        po, m, v = adamw.adamw_step(po, g, m, v, step, 2e-5 * 500, (0.9, 0.999), 1e-8, 0.01, clip=coef,
                                    param_dtype=torch.float32, state_dtype=torch.float32)
        assert float((po - p.detach()).abs().max()) <= 2e-6 * float(p.detach().abs().max())


class _FakeLora(nn.Module):
    def __init__(self, n, k, r, name="a"):
        super().__init__()
        self.lora_A = nn.ModuleDict({name: nn.Linear(k, r, bias=False)})
        self.lora_B = nn.ModuleDict({name: nn.Linear(r, n, bias=False)})
        self._grad_sinks = {}


def test_flat_parameter_buckets_alias_the_modules():
    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    opt = importlib.import_module("causal-unified-language-vision_b200.optim")
    torch.manual_seed(0)
    mods = [_FakeLora(64, 32, 8).bfloat16(), _FakeLora(32, 64, 8).bfloat16()]
    before = [m.lora_A["a"].weight.detach().clone() for m in mods]
    gs = par.GradSync(mods, "a", bucket_bytes=2000)
    o = opt.FusedLoraAdamW(gs, lr=1e-3, weight_decay=0.0)
    assert len(o.pflat) == len(gs.buckets) > 1
    for m, b in zip(mods, before):
        assert torch.equal(m.lora_A["a"].weight.detach(), b)          # values survive the re-homing
    # parameter storage is the flat bucket: writing the bucket changes the module weight
    o.pflat[0].zero_()
    assert float(mods[-1].lora_B["a"].weight.detach().abs().sum()) == 0.0
    assert sum(p.numel() for p in o.param_groups[0]["params"]) == sum(f.numel() for f in o.pflat)
    # it is a torch optimizer: LR schedulers drive it (trainer/cullavo_trainer.py:14)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(o, T_max=10, eta_min=1e-6)
    assert o.param_groups[0]["lr"] == 1e-3 and sched is not None


@pytest.mark.gpu
@pytest.mark.parametrize("state_dtype", [torch.bfloat16, torch.float32])
def test_fused_adamw_matches_oracle(lib_built, cuda_dev, state_dtype):
    from oracle import adamw

    par = importlib.import_module("causal-unified-language-vision_b200.parallel")
    opt = importlib.import_module("causal-unified-language-vision_b200.optim")
    torch.manual_seed(0)
    mods = [_FakeLora(256, 128, 64).bfloat16().to(cuda_dev), _FakeLora(128, 256, 64).bfloat16().to(cuda_dev),
            _FakeLora(72, 40, 8).bfloat16().to(cuda_dev)]   # last one: sizes that are not multiples of 8 per bucket tail
    gs = par.GradSync(mods, "a", bucket_bytes=40000)
    o = opt.FusedLoraAdamW(gs, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, state_dtype=state_dtype)
    ref_p = [f.float().cpu().clone() for f in o.pflat]
    ref_m = [torch.zeros_like(t) for t in ref_p]
    ref_v = [torch.zeros_like(t) for t in ref_p]
    for step in range(1, 4):
        for b in gs.buckets:
            b.flat.copy_((torch.randn(b.flat.numel(), device=cuda_dev) * (4.0 if step == 2 else 0.01)).bfloat16())
        grads = [b.flat.float().cpu() for b in gs.buckets]
        total = o.clip_grad_norm_(10.0)
        coef = adamw.clip_coef(grads, 10.0)
        expect_total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads))
        assert abs(float(total) - expect_total) <= 1e-3 * expect_total
        assert (coef < 1.0) == (step == 2)
        o.step()
        torch.cuda.synchronize()
        for i in range(len(ref_p)):
            p_, m_, v_ = adamw.adamw_step(ref_p[i], grads[i], ref_m[i], ref_v[i], step, 1e-2, (0.9, 0.999), 1e-8, 0.01,
                                          clip=coef, state_dtype=state_dtype)
            ref_p[i], ref_m[i], ref_v[i] = p_.float(), m_.float(), v_.float()
            got = o.pflat[i].float().cpu()
            # bf16 parameters: agreement to one bf16 ulp of the largest parameter (clip coefficient differs in the
            # last fp32 bits between the device's fp32 reduction and the oracle's fp64 one)
            assert float((got - ref_p[i]).abs().max()) <= 2 ** -7 * float(ref_p[i].abs().max())
            assert float((o.mflat[i].float().cpu() - ref_m[i]).abs().max()) <= 2 ** -7 * float(ref_m[i].abs().max()) + 1e-12
        # module weights ARE the bucket storage
        w = mods[0].lora_A["a"].weight
        pf = o.pflat[gs._slot_bucket[(id(mods[0]), "A")]]
        assert pf.data_ptr() <= w.data_ptr() < pf.data_ptr() + pf.numel() * 2
    o.zero_grad()
    assert all(float(b.flat.float().abs().sum()) == 0.0 for b in gs.buckets)
