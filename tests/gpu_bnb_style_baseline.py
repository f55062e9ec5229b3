"""Comparison point on the same B200 (NOT the judged reference arm): the reference's ALGORITHM as bitsandbytes + PEFT
would run it on a GPU -- dequantise the NF4 weight to a bf16 scratch (our own decode kernel standing in for
kDequantizeBlockwise), cuBLAS for the base GEMMs, stock torch ops for dropout and the LoRA branch, autograd for the
backward (which dequantises again) -- on the same linear-stack workload as bench.py.

    python tests/gpu_bnb_style_baseline.py [layers]
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

import b200qlora as q  # noqa: E402

F = q.functional
stackmod = importlib.import_module("causal-unified-language-vision_b200.stack")


class MatMul4BitStyle(torch.autograd.Function):
    """bitsandbytes.autograd._functions.MatMul4Bit: dequantise + F.linear forward, dequantise + matmul backward."""

    @staticmethod
    def forward(ctx, x, packed, qs):
        ctx.qs = qs
        ctx.save_for_backward(packed)
        return TF.linear(x, F.dequantize_4bit(packed, qs))

    @staticmethod
    def backward(ctx, dy):
        (packed,) = ctx.saved_tensors
        return dy @ F.dequantize_4bit(packed, ctx.qs), None, None


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    M = 16384
    st = stackmod.QLoRALinearStack(layers, stackmod.MISTRAL_LITERAL, M, r=64, dropout=0.05, device=dev, seed=0)
    mods = st.mods

    def step():
        outs = []
        for m in mods:
            x = st.inputs[m.in_features].detach().requires_grad_(True)
            base = m.base_layer
            A, B = m.lora_A["step1"].weight, m.lora_B["step1"].weight
            y = MatMul4BitStyle.apply(x, base.weight.data, base.weight.quant_state)
            y = y + TF.linear(TF.linear(TF.dropout(x, 0.05, True), A), B) * 0.25
            outs.append(y)
        for i in range(len(mods) - 1, -1, -1):
            torch.autograd.backward(outs[i], st.grads_out[outs[i].shape[-1]])
            outs[i] = None
        for m in mods:
            m.lora_A["step1"].weight.grad = None
            m.lora_B["step1"].weight.grad = None

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fpt = st.flops_per_token()
    print(json.dumps({"what": "bnb-style GPU path (dequant kernel + cuBLAS + torch LoRA/dropout + autograd), same stack",
                      "layers": layers, "ms_per_step": ms, "tokens_per_s": M / (ms / 1e3),
                      "step_tflops": fpt * M / (ms / 1e3) / 1e12 * layers / 32.0 if layers != 32 else fpt * M / (ms / 1e3) / 1e12,
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
