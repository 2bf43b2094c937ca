"""Seeded driver-API allocation traces (SURVEY.md §8d cfg 2) in the grammar of oracle/trace_replay.c.

Shared by the CPU tests, the GPU tests and tests/golden/make_golden.py so that the reference binary, the new
library and the CPU restatement all see byte-identical op streams.
"""
import math
import random

EDGE_SIZES = [1, (2 << 20) - 1, 2 << 20, (2 << 20) + 1]


def gen_trace(n_ops, seed=0xB200, max_size=512 << 20, p_alloc=0.60, with_info=True, kinds="A", max_live=None):
    """60 % alloc / 40 % free of a uniformly random live id; sizes log-uniform in [256 B, max_size] rounded to
    256 B plus the edge sizes around IPCSIZE; every 16th op is followed by a cuMemGetInfo probe."""
    rng = random.Random(seed)
    live = []
    next_id = 0
    lines = []
    lo, hi = math.log(256.0), math.log(float(max_size))
    for i in range(n_ops):
        do_alloc = rng.random() < p_alloc or not live
        if max_live is not None and len(live) >= max_live:
            do_alloc = False
        if do_alloc:
            if rng.random() < 0.02:
                size = rng.choice(EDGE_SIZES)
            else:
                size = int(math.exp(rng.uniform(lo, hi)))
                size = max(256, (size + 255) & ~255)
            kind = rng.choice(kinds)
            if kind == "P":
                w = max(4, (size // 64) & ~3) or 4
                lines.append(f"P {next_id} {w} 64")
            else:
                lines.append(f"{kind} {next_id} {size}")
            # the id stays "live" in the generator even if the hook refuses it: freeing a refused id is
            # cuMemFree_v2(0) == CUDA_SUCCESS in every implementation (reference @0x32383)
            live.append(next_id)
            next_id += 1
        else:
            k = rng.randrange(len(live))
            live[k], live[-1] = live[-1], live[k]
            lines.append(f"F {live.pop()}")
        if with_info and i % 16 == 15:
            lines.append("I")
        if with_info and i % 4096 == 4095:
            lines.append("X 0x1234000")
            lines.append("T")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sys.stdout.write(gen_trace(n))
