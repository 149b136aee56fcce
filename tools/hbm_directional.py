"""HBM bandwidth by direction on one GPU: write-only (fill), read-only (reduction), copy.  Measurement aid for DESIGN.md's
roofline discussion (kv_kernel is write-only, ks_kernel read-only, kh_kernel half and half).  torch kernels, CUDA events."""
import json
import torch


def timed(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = 1 << 29                                     # 4 GiB of float64: far beyond the 126 MB L2
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    x.fill_(1.0); y.fill_(2.0); torch.cuda.synchronize()
    out = {}
    out["write_only_GBs"] = n * 8 / timed(lambda: x.fill_(3.0)) / 1e6
    out["read_only_GBs"] = n * 8 / timed(lambda: x.sum()) / 1e6
    out["copy_GBs"] = 2 * n * 8 / timed(lambda: y.copy_(x)) / 1e6
    out["read2_write1_GBs"] = 3 * n * 8 / timed(lambda: torch.add(x, y, out=y)) / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
