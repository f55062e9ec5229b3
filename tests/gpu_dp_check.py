"""Multi-GPU data-parallel parity check (run under torchrun on N >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_dp_check.py

Every rank runs the linear stack on its own shard of the batch with the overlapped NCCL gradient sync;
rank 0 then replays ALL shards locally (single process, same weights) and checks that the all-reduced
(mean) LoRA gradient buckets equal the mean of the per-shard gradients (SURVEY.md section 8c test (10)).
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stackmod = importlib.import_module("causal-unified-language-vision_b200.stack")
    shapes = [("q_proj", 512, 512), ("up_proj", 1024, 512), ("down_proj", 512, 1024)]
    M = 640
    st = stackmod.QLoRALinearStack(2, shapes, M, r=64, dropout=0.05, device=dev, seed=5, bucket_bytes=200_000)
    assert len(st.sync.buckets) > 1
    # are the replicas identical before any broadcast?  (base weights are seeded; LoRA init uses the global RNGs)
    def digest(t):
        return float(t.float().double().sum()), float(t.float().double().abs().sum())
    for name, t in (("packed[0]", st.mods[0].base_layer.weight.data), ("A[0]", st.mods[0].lora_A["step1"].weight.data),
                    ("B[0]", st.mods[0].lora_B["step1"].weight.data)):
        print(f"[rank {rank}] digest {name}: {digest(t)}", flush=True)
    st.sync.broadcast_parameters(0)
    # (a) local gradients of this rank's shard, no communication
    st.sync.world = 1
    st.step_direct()
    torch.cuda.synchronize()
    local = [f.float().clone() for f in st.sync.flat_grads()]
    # (b) the same step with the overlapped NCCL all-reduce(mean)
    st.sync.world = world
    st._step = 0
    st.step_direct()
    torch.cuda.synchronize()
    got = [f.float().clone() for f in st.sync.flat_grads()]
    ok = torch.ones(1, device=dev)
    worst = 0.0
    for li, g in zip(local, got):
        ref = li.clone()
        dist.all_reduce(ref)          # fp32 sum of the local gradients
        ref /= world
        err = float((g - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
    print(f"[rank {rank}] all-reduced (mean, bf16 buckets) vs fp32 mean of local grads: max rel err {worst:.3e}", flush=True)
    if not worst < 2e-2:
        ok.zero_()
    # (c) rank 0 replays every rank's shard locally: checks that weights / seeds are identical across ranks
    ins = {k: [torch.empty_like(v) for _ in range(world)] for k, v in st.inputs.items()}
    gos = {k: [torch.empty_like(v) for _ in range(world)] for k, v in st.grads_out.items()}
    for k, v in st.inputs.items():
        dist.all_gather(ins[k], v)
    for k, v in st.grads_out.items():
        dist.all_gather(gos[k], v)
    locs = [[torch.empty_like(t) for _ in range(world)] for t in local]
    for t, dst in zip(local, locs):
        dist.all_gather(dst, t)
    if rank == 0:
        st.sync.world = 1
        for r in range(world):
            st.inputs = {k: ins[k][r] for k in ins}
            st.grads_out = {k: gos[k][r] for k in gos}
            st._step = 0
            st.step_direct()
            torch.cuda.synchronize()
            w2 = 0.0
            for b, f in enumerate(st.sync.flat_grads()):
                ref = locs[b][r]
                w2 = max(w2, float((f.float() - ref).abs().max() / ref.abs().max()))
            print(f"[rank 0] replay of rank {r}'s shard vs its own local grads: max rel err {w2:.3e}", flush=True)
            if not w2 < 1e-6:
                ok.zero_()
    dist.broadcast(ok, 0)
    dist.destroy_process_group()
    if float(ok) != 1.0:
        sys.exit(1)
    if rank == 0:
        print("dp check OK")


if __name__ == "__main__":
    main()
