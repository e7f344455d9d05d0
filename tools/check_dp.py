"""Data-parallel correctness on N GPUs (torchrun): the all-reduced gradients and the SGD-updated weights of the
FlatDataParallel / FlatSGD path must equal the single-process result on the concatenated global batch (SURVEY.md §8e:
mean over ranks of per-rank mean losses == mean loss of the global batch when every rank has the same batch size).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_dp.py
    ... tools/check_dp.py --one-gpu      both ranks on cuda:0 over gloo (NCCL refuses two ranks on one device): the same
                                         FlatDataParallel / FlatSGD code path - bf16 gradient payload, reduce-scatter of the
                                         AVT-h gradients, sharded update, all-gather of the bf16 weights at the next step's
                                         start, per-layer backbone all-reduces - on a single-GPU box
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from avt_b200.model import AVTModel
from avt_b200.optim import FlatSGD
from avt_b200.parallel import FlatDataParallel


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_gpu = "--one-gpu" in sys.argv
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_MAX_CTAS", "16")
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    steps = 3
    hk = dict(n_head=2, n_layer=2, inter_dim=64, n_positions=32, embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    B, T = 2, 4

    def build():
        torch.manual_seed(0)
        return AVTModel("vit_test_patch16_32", 64, 32, dropout=0.0, head_kwargs=hk).to(dev).train()

    def loss_of(m, video):
        out, aux = m(video, target_shape=(video.shape[0],))
        return out["logits/action"].square().mean() + out["past_logits/action"].square().mean() + aux["feat"].mean()

    g = torch.Generator().manual_seed(7)
    videos = [torch.randn(B, T, 3, 1, 32, 32, generator=g) for _ in range(world)]   # every rank knows every shard

    # data-parallel model: this rank's shard only
    m = build()
    dp = FlatDataParallel(m)
    opt = None
    for _ in range(steps):
        dp.begin_step()          # all-gather of the AVT-h bf16 weights the sharded optimizer updated in the previous step
        loss = loss_of(m, videos[rank].to(dev))
        if opt is None:
            dp.broadcast_parameters()
            opt = FlatSGD([dp.vit, dp.head], dp.other, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3,
                          bias_bn_wd_scale=0.5)
        opt.zero_grad()
        loss.backward()
        dp.finish_backward(opt)
    sd = m.state_dict()          # (every rank: the head's state_dict hook all-gathers the fp32 master shards)
    assert not dp._master_stale
    # reference: same weights, the global batch on one GPU, stock autograd + torch SGD
    ref = build()
    # the reference's parameter groups (func/train.py:704-731): names ending in 'bias' decay with wd * bias_bn_wd_scale
    named = list(ref.named_parameters())
    ropt = torch.optim.SGD([dict(params=[p for n, p in named if not n.endswith("bias")], weight_decay=1e-3),
                            dict(params=[p for n, p in named if n.endswith("bias")], weight_decay=0.5e-3)],
                           lr=0.05, momentum=0.9, nesterov=True)
    for _ in range(steps):
        ropt.zero_grad()
        loss_of(ref, torch.cat(videos, 0).to(dev)).backward()     # one forward: the global batch
        ropt.step()
    torch.cuda.synchronize()
    worst = 0.0
    rp = dict(ref.named_parameters())
    errs = {}
    for n, p in sd.items():
        errs[n] = ((p.detach() - rp[n].detach()).norm() / (rp[n].detach().norm() + 1e-30)).item()
        worst = max(worst, errs[n])
    if rank == 0 and "--verbose" in sys.argv:
        for n in sorted(errs, key=errs.get, reverse=True)[:8]:
            print(f"  {n}: {errs[n]:.3e}  (norm {rp[n].detach().norm().item():.3e})", flush=True)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"check_dp: world {world} ({'gloo, one GPU' if one_gpu else 'nccl'}), worst relative weight difference after {steps} steps "
              f"{t.item():.3e}", flush=True)
    # bf16 gradient payload (matrices): with this deliberately huge lr (0.05, 500x the reference's 1e-4) the weights move by
    # O(1) of their norm in 3 steps, so the 2^-9 rounding of the gradients shows up as ~1.5e-3 of the weights (measured
    # 1.75e-3 worst: cls_token); fp32 payload gave 4.6e-5 in round 1. A wrong reduction (sum instead of mean, a stale
    # shard, a missing all-gather) is O(1).
    ok = t.item() < 5e-3
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
