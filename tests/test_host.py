"""Host-side mirror of the reference's module surface (CPU: construction, naming, state, errors)."""
import importlib
import os
import sys

import pytest
import torch
import torch.nn as nn

import b200qlora as q

lora = importlib.import_module("causal-unified-language-vision_b200.lora")
qnn = importlib.import_module("causal-unified-language-vision_b200.nn")


class Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.o_proj = nn.Linear(d, d, bias=False)


class Mlp(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.gate_proj = nn.Linear(d, f, bias=False)
        self.up_proj = nn.Linear(d, f, bias=False)
        self.down_proj = nn.Linear(f, d, bias=False)


class Layer(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.self_attn = Attn(d)
        self.mlp = Mlp(d, f)


class Toy(nn.Module):
    def __init__(self, d=256, f=512, n=3):
        super().__init__()
        self.layers = nn.ModuleList([Layer(d, f) for _ in range(n)])
        self.lm_head = nn.Linear(d, 1000, bias=False)


def _toy4bit():
    m = Toy()
    lora.replace_with_4bit_linear(m, modules_to_not_convert=["multi_modal_projector", "lm_head"])
    return m


def test_replace_keeps_skip_modules_and_linear_subclassing():
    m = _toy4bit()
    assert type(m.lm_head) is nn.Linear
    assert isinstance(m.layers[0].self_attn.q_proj, qnn.Linear4bit)
    assert isinstance(m.layers[0].self_attn.q_proj, nn.Linear)  # load_cullavo.py:9-14 relies on this
    assert isinstance(m.layers[0].self_attn.q_proj.weight, qnn.Params4bit)
    # the reference's own discovery helper finds exactly the 7 projection names
    assert lora.find_all_linear_names(m) == sorted(["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj",
                                                    "down_proj"])


def test_add_adapter_names_and_trainability():
    m = _toy4bit()
    cfg = lora.LoraConfig(r=64, lora_alpha=16, target_modules=lora.find_all_linear_names(m), lora_dropout=0.05,
                          bias="none", task_type="CAUSAL_LM")
    lora.add_adapter(m, cfg, adapter_name="step1")
    names = dict(m.named_parameters())
    assert "layers.0.self_attn.q_proj.base_layer.weight" in names
    assert "layers.0.self_attn.q_proj.lora_A.step1.weight" in names
    assert "layers.2.mlp.down_proj.lora_B.step1.weight" in names
    assert names["layers.0.self_attn.q_proj.lora_A.step1.weight"].shape == (64, 256)
    assert names["layers.2.mlp.down_proj.lora_B.step1.weight"].shape == (256, 64)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert trainable and all("lora" in n for n in trainable)
    assert len(trainable) == 3 * 7 * 2
    mod = m.layers[0].self_attn.q_proj
    assert mod.scaling["step1"] == 0.25 and mod.r["step1"] == 64
    assert isinstance(mod.lora_dropout["step1"], nn.Dropout) and mod.lora_dropout["step1"].p == 0.05
    assert float(mod.lora_B["step1"].weight.detach().abs().max()) == 0.0  # PEFT init: B = 0
    # second adapter (load_cullavo.py:23-42): resident, and only the new one is active / trainable
    cfg2 = lora.LoraConfig(r=64, lora_alpha=16, target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj",
                                                                "up_proj", "down_proj"], lora_dropout=0.05)
    lora.add_adapter(m, cfg2, adapter_name="step2")
    assert mod.active_adapters == ["step2"]
    assert not mod.lora_A["step1"].weight.requires_grad and mod.lora_A["step2"].weight.requires_grad
    assert set(m.peft_config) == {"step1", "step2"}


def test_layers_to_transform():
    m = _toy4bit()
    cfg = lora.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj"], layers_to_transform=[1, 2])
    lora.add_adapter(m, cfg, "a")
    assert isinstance(m.layers[0].self_attn.q_proj, qnn.Linear4bit)
    assert isinstance(m.layers[1].self_attn.q_proj, lora.LoraLinear4bit)


def test_bf16_sweep_leaves_packed_weights_alone():
    m = _toy4bit()
    lora.add_adapter(m, lora.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj"]), "a")
    for p in m.parameters():  # load_cullavo.py:124-126
        if "float32" in str(p.dtype):
            p.data = p.data.to(torch.bfloat16)
    assert m.layers[0].self_attn.q_proj.lora_A["a"].weight.dtype == torch.bfloat16


def test_adapter_state_dict_uses_peft_naming(tmp_path):
    m = _toy4bit()
    lora.add_adapter(m, lora.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "down_proj"]), "step1")
    sd = lora.get_adapter_state_dict(m, "step1")
    assert "base_model.model.layers.0.self_attn.q_proj.lora_A.weight" in sd
    with torch.no_grad():
        for p in m.parameters():
            if p.requires_grad:
                p.normal_()
    path = str(tmp_path / "adapter_model.safetensors")
    lora.save_adapter(m, path, "step1")
    m2 = _toy4bit()
    lora.add_adapter(m2, lora.LoraConfig(r=8, lora_alpha=16, target_modules=["q_proj", "down_proj"]), "step1")
    lora.load_adapter(m2, path, "step1")
    a = dict(m.named_parameters())
    b = dict(m2.named_parameters())
    for n in a:
        if "lora" in n:
            assert torch.equal(a[n], b[n]), n


def test_quant_state_dict_round_trip_keys():
    code = torch.tensor(q.functional.NF4_CODE)
    s2 = q.QuantState(torch.rand(2), code=q.functional.create_dynamic_map(), blocksize=256, quant_type="dynamic8",
                      dtype=torch.float32)
    qs = q.QuantState(torch.randint(0, 255, (512,), dtype=torch.uint8), (128, 256), code, 64, "nf4", torch.bfloat16,
                      offset=torch.tensor(0.03), state2=s2)
    d = qs.as_dict(packed=True)
    assert set(d) == {"absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state.bitsandbytes__nf4"}
    back = q.QuantState.from_dict(d, device="cpu")
    assert back.nested and tuple(back.shape) == (128, 256) and back.blocksize == 64
    assert torch.equal(back.absmax, qs.absmax) and torch.equal(back.state2.absmax, s2.absmax)
    assert abs(float(back.offset) - 0.03) < 1e-7


def test_dynamic_map_matches_oracle():
    import numpy as np
    from oracle import nf4

    assert np.array_equal(q.functional.create_dynamic_map().numpy(), nf4.create_dynamic_map())
    assert np.array_equal(np.array(q.functional.NF4_CODE, dtype=np.float32), nf4.NF4_CODE)


def test_cpu_tensors_fail_loudly(lib_built):
    """No CPU fallback: the product path raises instead of computing on the host."""
    x = torch.randn(8, 256).bfloat16()
    with pytest.raises(RuntimeError, match="GPU only"):
        q.functional.quantize_4bit(torch.randn(64, 64))
    m = _toy4bit()
    with pytest.raises(RuntimeError, match="quantization state not initialized"):
        m.layers[0].self_attn.q_proj(x)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "causal-unified-language-vision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("extra", [[], ["--world", "2"], ["--hang", "modules"], ["--world", "2", "--hang", "modules"],
                                   ["--hang", "raise"]])
def test_bench_control_flow_dry_run(extra):
    """bench.py's phases, watchdog and JSON line with every device call stubbed (tests/dev_bench_dryrun.py): N = 1,
    N > 1, a module-surface step that never returns and one that dies with a CUDA error -- at every N the complete
    device-timed line must still be printed, with "e2e": null and a note, and the process must exit non-zero quickly."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "dev_bench_dryrun.py"), *extra], capture_output=True,
                       text=True, timeout=300, cwd=root)
    if "--hang" in extra:
        assert r.returncode == 5, (r.returncode, r.stderr[-2000:])
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        assert d["e2e"] is None and "did not complete" in d["note"] and d["value"] > 0 and d["roofline"]["achieved"] >= 0
    else:
        assert r.returncode == 0, r.stderr[-2000:]
        assert "dry run OK" in r.stdout


def test_pipeline_protocol_model():
    """tests/sim_pipeline_protocol.py: the barrier protocols of the kernel family (operand ring + accumulator, two epilogue
    sets, transform groups, decode groups + packed ring) run to completion under random interleavings AND out-of-order
    load completion with no phase aliasing and no operand / accumulator hazard -- and the three round-1 protocols that
    stalled or corrupted on hardware fail in the model (its own mutation checks)."""
    import importlib.util

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim_pipeline_protocol.py")
    spec = importlib.util.spec_from_file_location("sim_pipeline_protocol", path)
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    assert sim.run_all(seeds=6) == 9 * 4 * 6


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle on the host cores) runs without a GPU and prints the JSON contract."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["metric"] == "qlora_linear_stack_train_tokens_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] and d["higher_is_better"] is True and d["vs_baseline"] is None
