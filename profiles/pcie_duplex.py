"""Does this host overlap H2D and D2H copies (pinned memory, two streams)?  Prints GB/s alone and together."""
import time
import torch
n = 2 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


run(True, True, 1)
a, b, c = run(True, False), run(False, True), run(True, True)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both together %.1f + %.1f GB/s (%.3f s vs %.3f s serial)" % (
    n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9, c, a + b))
