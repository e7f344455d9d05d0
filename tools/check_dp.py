"""Data-parallel correctness on N GPUs (torchrun): the all-reduced gradients and the SGD-updated weights of the
FlatDataParallel / FlatSGD path must equal the single-process result on the concatenated global batch (SURVEY.md §8e:
mean over ranks of per-rank mean losses == mean loss of the global batch when every rank has the same batch size).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_dp.py
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_MAX_CTAS", "16")
    dist.init_process_group("nccl", device_id=dev)
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
    for _ in range(2):
        loss = loss_of(m, videos[rank].to(dev))
        if opt is None:
            dp.broadcast_parameters()
            opt = FlatSGD([dp.vit, dp.head], dp.other, lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
        for p in dp.other:
            p.grad = None
        loss.backward()
        dp.finish_backward(opt)
    # reference: same weights, the global batch on one GPU, stock autograd + torch SGD
    ref = build()
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-3)
    for _ in range(2):
        ropt.zero_grad()
        loss_of(ref, torch.cat(videos, 0).to(dev)).backward()     # one forward: the global batch
        ropt.step()
    torch.cuda.synchronize()
    worst = 0.0
    rp = dict(ref.named_parameters())
    for n, p in m.named_parameters():
        e = ((p.detach() - rp[n].detach()).norm() / (rp[n].detach().norm() + 1e-30)).item()
        worst = max(worst, e)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"check_dp: world {world}, worst relative weight difference after 2 steps {t.item():.3e}", flush=True)
    ok = t.item() < 5e-4
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
