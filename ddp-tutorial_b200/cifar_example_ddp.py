"""Drop-in counterpart of /root/reference/cifar_example_ddp.py on the B200-native path.

Same entry contract (argparse with the unused ``--world_size``; RANK / WORLD_SIZE / LOCAL_RANK
from ``torchrun --nproc_per_node=N``), same construction order and step body
(cifar_example_ddp.py:82-114), but the model / wrapper / criterion / optimizer are the
``ddp_tutorial_b200`` objects, so every kernel of the step is launched by libb2ddp.so.

    torchrun --nproc_per_node=8 ddp-tutorial_b200/cifar_example_ddp.py --model resnet18 --batch-size 512

With ``./data/cifar-10-batches-py`` present the uint8 images live in HBM, are indexed with the
reference's ``DistributedSampler`` semantics and normalised on the device;
otherwise (no network here, so no download) the input is synthetic CIFAR-10-shaped data.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddp_tutorial_b200 as b2  # noqa: E402


def load_cifar10(root, train):
    """CIFAR-10 python batches as uint8 [N,32,32,3] + int64 labels (what torchvision.datasets.CIFAR10
    decodes, cifar_example_ddp.py:67-68); None when the files are not there (no network in this build)."""
    import pickle
    import numpy as np
    base = os.path.join(root, "cifar-10-batches-py")
    names = [f"data_batch_{i}" for i in range(1, 6)] if train else ["test_batch"]
    if not all(os.path.exists(os.path.join(base, n)) for n in names):
        return None
    xs, ys = [], []
    for n in names:
        with open(os.path.join(base, n), "rb") as f:
            d = pickle.load(f, encoding="latin1")
        xs.append(np.asarray(d["data"], dtype=np.uint8).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1))
        ys.extend(d["labels"])
    return torch.from_numpy(np.ascontiguousarray(np.concatenate(xs))), torch.tensor(ys, dtype=torch.int64)


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
    running = b2.RunningLoss(args.gpu)                                        # :110 running_loss, kept on the device
    criterion = b2.CrossEntropyLoss(running=running)                          # :85
    optimizer = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)              # :86

    gen = torch.Generator().manual_seed(args.seed + 1 + rank)
    train = load_cifar10(args.data, True)
    test = load_cifar10(args.data, False)
    world = args.world_size if args.distributed else 1
    if train is not None:                                                     # :67-71, images resident in HBM
        sampler_train = b2.DistributedSampler(train[0].shape[0], num_replicas=world, rank=rank, shuffle=True)
        trainloader = b2.DeviceCifarLoader(train[0], train[1], batch_size, sampler_train, args.gpu)

    def synthetic():
        for _ in range(args.steps):
            x = torch.randn(batch_size, 3, 32, 32, generator=gen).pin_memory()
            y = torch.randint(0, 10, (batch_size,), generator=gen).pin_memory()
            yield x.to(args.gpu, non_blocking=True), y.to(args.gpu, non_blocking=True)          # :97-98

    for epoch in range(args.epochs):                                          # :90
        running.mean_and_reset()
        if train is not None:
            sampler_train.set_epoch(epoch)                                    # :92
        for i, (inputs, labels) in enumerate(trainloader if train is not None else synthetic()):   # :94
            optimizer.zero_grad()                                             # :101
            outputs = net(inputs)                                             # :104
            loss = criterion(outputs, labels)                                 # :105
            loss.backward()                                                   # :106
            optimizer.step()                                                  # :107
            # :110 `running_loss += loss.item()` happens inside the criterion kernel (no host sync per step)
            if i % args.print_every == args.print_every - 1:                  # :111-114
                print('[%d, %5d] loss: %.3f' % (epoch + 1, i + 1, running.mean_and_reset()))
                b2.check(net)                                                 # surfaces a timed-out barrier / peer wait
    print('Finished Training')
    if rank == 0 and args.save:
        torch.save(net.state_dict(), args.save)                               # :118-119 (single writer)
    if test is not None:                                                      # :121-136
        sampler_test = b2.DistributedSampler(test[0].shape[0], num_replicas=world, rank=rank, shuffle=False)
        accuracy = b2.Accuracy(args.gpu)
        for images, labels in b2.DeviceCifarLoader(test[0], test[1], batch_size, sampler_test, args.gpu):
            accuracy.update(net(images), labels)          # batch-stat BatchNorm, as the reference never calls .eval()
        print('Accuracy of the network on the 10000 test images: %d %%' % (100 * accuracy.compute()))
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
    parser.add_argument("--data", default="./data", help="root holding cifar-10-batches-py (synthetic input if absent)")
    parser.add_argument("--save", default="", help="path for state_dict (reference: ./cifar_net.pth)")
    main(parser.parse_args())
