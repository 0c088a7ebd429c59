"""B200-native drop-in for the DDP hot path of rensortino/DDP-Tutorial.

    import ddp_tutorial_b200 as b2
    b2.init_distributed(args)
    net = b2.Net().to(args.gpu)                                   # or b2.resnet18(num_classes=10)
    net = b2.DistributedDataParallel(net, device_ids=[args.gpu])
    criterion = b2.CrossEntropyLoss()
    optimizer = b2.SGD(net.parameters(), lr=0.001, momentum=0.9)
    optimizer.zero_grad(); loss = criterion(net(x), y); loss.backward(); optimizer.step(); loss.item()
"""
from ._lib import B2Error, LIB_PATH, lib  # noqa: F401
from .nn import CrossEntropyLoss, Loss, Module, Net, ResNet, RunningLoss, resnet18, resnet50  # noqa: F401
from .data import Accuracy, DeviceCifarLoader, DevicePrefetcher, DistributedSampler, normalize_u8  # noqa: F401
from .optim import SGD  # noqa: F401
from .parallel import DistributedDataParallel, barrier, check, destroy, init_distributed  # noqa: F401

__all__ = ["Accuracy", "DeviceCifarLoader", "DevicePrefetcher", "DistributedSampler", "normalize_u8", "B2Error", "CrossEntropyLoss", "DistributedDataParallel", "Net", "ResNet", "RunningLoss", "SGD", "barrier", "check",
           "destroy", "init_distributed", "lib", "resnet18", "resnet50"]
