"""Shared pieces of the training benchmark scripts (synthetic / torchvision batches, shape scaling,
per-step timing and the reference's console lines)."""
import time

import numpy as np
import torch


def batches(args, image_size, n_images, steps):
    """Yield (images, labels) host batches of `n_images` images.  APP 3 = synthetic (pinned when a
    GPU is present); APP 1 / 2 = ImageFolder / CIFAR10 through torchvision, like the reference."""
    if args.app == 3:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_images, 3, image_size, image_size, generator=g)
        y = torch.randint(0, args.num_classes, (n_images,), generator=g)
        if torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        for _ in range(steps):
            yield x, y
        return
    import torchvision
    import torchvision.transforms as transforms
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    torch.manual_seed(0)
    if args.app == 1:
        ds = torchvision.datasets.ImageFolder(args.datapath, transform=tf)
    else:
        ds = torchvision.datasets.CIFAR10(root=args.datapath, train=True, download=False, transform=tf)
    dl = torch.utils.data.DataLoader(ds, batch_size=n_images, shuffle=(args.app == 1), num_workers=args.num_workers,
                                     pin_memory=True, drop_last=True)
    yield from dl


def scale_shapes(shape_list, times):
    """Stage output shapes traced at a small image -> the real image (H, W x times); 2-D shapes pass."""
    def one(s):
        return tuple(s) if len(s) == 2 else (s[0], s[1], int(s[2] * times), int(s[3] * times))
    return [[one(t) for t in s] if isinstance(s, list) else one(s) for s in shape_list]


class StepTimer:
    """CUDA-event timing of a step when a GPU is present, wall clock otherwise."""

    def __init__(self):
        self.cuda = torch.cuda.is_available()

    def __enter__(self):
        if self.cuda:
            self.t0, self.t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.t0.record()
        else:
            self.w0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if self.cuda:
            self.t1.record()
            torch.cuda.synchronize()
            self.seconds = self.t0.elapsed_time(self.t1) / 1000
        else:
            self.seconds = time.perf_counter() - self.w0
        return False


def report(perf):
    if perf:
        print("Mean %s Median %s" % (sum(perf) / len(perf), np.median(perf)), flush=True)


def pop_model_flag(argv, default):
    if "--model" in argv:
        i = argv.index("--model")
        kind = argv[i + 1]
        del argv[i:i + 2]
        return kind
    return default


def build_sequential(kind, args, mb, image_size):
    """(small tracing model, its image size, full-size model) of the NON-spatial builders."""
    from mpi4dl_b200.torchgems.utils import get_depth
    if kind == "resnet":
        from mpi4dl_b200.models import resnet
        depth = get_depth(2, 12)
        return (resnet.get_resnet_v2((mb, 3, 32, 32), depth, num_classes=args.num_classes), 32,
                lambda: resnet.get_resnet_v2((mb, 3, image_size, image_size), depth, num_classes=args.num_classes))
    from mpi4dl_b200.models import amoebanet
    mk = lambda: amoebanet.amoebanetd(num_classes=args.num_classes, num_layers=args.num_layers,  # noqa: E731
                                      num_filters=args.num_filters)
    return mk(), min(512, image_size), mk
