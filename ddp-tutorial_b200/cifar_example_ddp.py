"""Drop-in counterpart of /root/reference/cifar_example_ddp.py on the B200-native path.

Same entry contract (argparse with the unused ``--world_size``; RANK / WORLD_SIZE / LOCAL_RANK
from ``torchrun --nproc_per_node=N``), same construction order and step body
(cifar_example_ddp.py:82-114), but the model / wrapper / criterion / optimizer are the
``ddp_tutorial_b200`` objects, so every kernel of the step is launched by libb2ddp.so.

    torchrun --nproc_per_node=8 ddp-tutorial_b200/cifar_example_ddp.py --model resnet18 --batch-size 512

Input is synthetic CIFAR-10-shaped data (3x32x32, N(0,1); uniform labels): the dataset download
and DataLoader of the reference are outside the hot path this build covers.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddp_tutorial_b200 as b2  # noqa: E402


def main(args):
    batch_size = args.batch_size
    b2.init_distributed(args)                                                 # :69 / :42-58
    rank = int(os.environ.get("RANK", "0"))
    torch.manual_seed(args.seed)
    if args.model == "net":
        net = b2.Net().to(args.gpu)                                           # :82
    else:
        net = getattr(b2, args.model)(num_classes=10, precision=args.precision).to(args.gpu)
    if args.distributed:
        net = b2.DistributedDataParallel(net, device_ids=[args.gpu])          # :83
    criterion = b2.CrossEntropyLoss()                                         # :85
    optimizer = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)              # :86

    gen = torch.Generator().manual_seed(args.seed + 1 + rank)
    for epoch in range(args.epochs):                                          # :90
        running_loss = 0.0
        for i in range(args.steps):                                           # :94
            inputs = torch.randn(batch_size, 3, 32, 32, generator=gen).pin_memory()
            labels = torch.randint(0, 10, (batch_size,), generator=gen).pin_memory()
            inputs = inputs.to(args.gpu, non_blocking=True)                   # :97
            labels = labels.to(args.gpu, non_blocking=True)                   # :98
            optimizer.zero_grad()                                             # :101
            outputs = net(inputs)                                             # :104
            loss = criterion(outputs, labels)                                 # :105
            loss.backward()                                                   # :106
            optimizer.step()                                                  # :107
            running_loss += loss.item()                                       # :110
            if i % args.print_every == args.print_every - 1:                  # :111-114
                print('[%d, %5d] loss: %.3f' % (epoch + 1, i + 1, running_loss / args.print_every))
                running_loss = 0.0
    print('Finished Training')
    if rank == 0 and args.save:
        torch.save(net.state_dict(), args.save)                               # :118-119 (single writer)
    if args.distributed:
        b2.destroy()


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument("--world_size", help="Total number of nodes")         # kept; overwritten from the env (:44)
    parser.add_argument("--model", default="net", choices=["net", "resnet18", "resnet50"])
    parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--batch-size", type=int, default=4)                   # reference: batch_size = 4 (:65)
    parser.add_argument("--epochs", type=int, default=2)                       # reference: range(2) (:90)
    parser.add_argument("--steps", type=int, default=200, help="synthetic mini-batches per epoch")
    parser.add_argument("--print-every", type=int, default=100)               # reference: 2000
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--save", default="", help="path for state_dict (reference: ./cifar_net.pth)")
    main(parser.parse_args())
