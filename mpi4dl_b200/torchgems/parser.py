"""torchgems.parser mirror: the flag set shared by the reference's model benchmarks
(reference src/torchgems/parser.py:21-143) -- same names, types and defaults."""
import argparse


def get_parser():
    p = argparse.ArgumentParser(description="SP-MP-DP Configuration Script", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-v", "--verbose", help="Prints performance numbers or logs", action="store_true")
    p.add_argument("--batch-size", type=int, default=32, help="input batch size")
    p.add_argument("--parts", type=int, default=1, help="Number of parts for MP")
    p.add_argument("--split-size", type=int, default=2, help="Number of process for MP")
    p.add_argument("--num-spatial-parts", type=str, default="4", help="Number of partitions in spatial parallelism")
    p.add_argument("--spatial-size", type=int, default=1, help="Number splits for spatial parallelism")
    p.add_argument("--times", type=int, default=1, help="Number of times to repeat MASTER 1: 2 repications, 2: 4 replications")
    p.add_argument("--image-size", type=int, default=32, help="Image size for synthetic benchmark")
    p.add_argument("--num-epochs", type=int, default=1, help="Number of epochs")
    p.add_argument("--num-layers", type=int, default=18, help="Number of layers in amoebanet")
    p.add_argument("--num-filters", type=int, default=416, help="Number of layers in amoebanet")
    p.add_argument("--num-classes", type=int, default=10, help="Number of classes")
    p.add_argument("--balance", type=str, default=None,
                   help="length of list equals to number of partitions and sum should be equal to num layers")
    p.add_argument("--halo-D2", dest="halo_d2", action="store_true", default=False,
                   help="Enable design2 (do halo exhange on few convs) for spatial conv. ")
    p.add_argument("--fused-layers", type=int, default=1,
                   help="When D2 design is enables for halo exchange, number of blocks to fuse in ResNet model ")
    p.add_argument("--local-DP", dest="local_DP", type=int, default=1,
                   help="LBANN intergration of SP with MP. MP can apply data parallelism. 1: only one GPU for a given "
                        "split, 2: two gpus for a given split (uses DP)")
    p.add_argument("--slice-method", type=str, default="square",
                   help="Slice method (square, vertical, and horizontal) in Spatial parallelism")
    p.add_argument("--app", type=int, default=3,
                   help="Application type (1.medical, 2.cifar, and synthetic) in Spatial parallelism")
    p.add_argument("--datapath", type=str, default="./train", help="local Dataset path")
    p.add_argument("--enable-master-comm-opt", action="store_true", default=False,
                   help="Enable communication optimization for MASTER in Spatial")
    p.add_argument("--num-workers", type=int, default=0, help="Data loader workers")
    return p
