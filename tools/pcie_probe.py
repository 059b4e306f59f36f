"""Host<->device copy rates of this box (pinned memory), alone and in both directions at once:
the ceiling of bench.py's `e2e` (21.0 MB in + 11.8 MB out per step)."""
import torch
dev = torch.device("cuda:0")
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    ms = timed(fn)
    print("%-5s 64 MiB: %.3f ms  %.1f GB/s%s" % (name, ms, n / ms / 1e6, " per direction" if name == "both" else ""))
