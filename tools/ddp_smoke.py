"""Two-rank (or N-rank) DDP training smoke step: the reference's evit_tiny_p16 (oracle/_ref/models) with the drop-in EVA attention,
fp16 autocast + GradScaler-free SGD on a fixed synthetic batch (vit/engine.py:47-62, vit/main.py:286-288).  Checks that the NCCL
gradient all-reduce leaves every rank with identical parameters and that the loss goes down.  Launch with torchrun."""
import os
import sys
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from oracle import ref_loader
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        vm = ref_loader.vit_models()
        torch.manual_seed(0)
        args = ref_loader.deit_args('eva', num_classes=10)
        args.drop_path_rate = 0.0
        model = vm.evit_tiny_p16(args).to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(ddp.parameters(), lr=0.05, momentum=0.9)
    torch.manual_seed(100 + rank)                                   # every rank its own shard of the (synthetic) batch
    x = torch.randn(8, 3, 224, 224, device=dev)
    y = torch.randint(0, 10, (8,), device=dev)
    losses = []
    for step in range(6):
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.float16):
            logits = ddp(x)
        loss = torch.nn.functional.cross_entropy(logits.float(), y)
        loss.backward()
        opt.step()
        t = loss.detach().clone()
        dist.all_reduce(t)
        losses.append(float(t) / world)
    checksum = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum()
    gathered = [torch.zeros_like(checksum) for _ in range(world)]
    dist.all_gather(gathered, checksum)
    same = all(float(g) == float(gathered[0]) for g in gathered)
    if rank == 0:
        print('losses', ' '.join(f'{l:.4f}' for l in losses), 'params identical across ranks:', same)
        ok = same and all(map(lambda v: v == v, losses)) and losses[-1] < losses[0]
        print('DDP_SMOKE_OK' if ok else 'DDP_SMOKE_FAILED')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
