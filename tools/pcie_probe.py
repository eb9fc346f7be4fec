"""Dev probe: H2D / D2H bandwidth of eye-sized images, contiguous vs pitched (2-D) copies, one and two directions."""
import time, torch
dev = torch.device("cuda:0")
def bw(fn, nbytes, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return nbytes * reps / (time.perf_counter() - t) / 1e9
H, W = 2492, 2244 * 4
hp = torch.empty((H, W), dtype=torch.uint8).pin_memory()
dc = torch.empty((H, W), dtype=torch.uint8, device=dev)
dp = torch.empty((H, 9216), dtype=torch.uint8, device=dev)[:, :W]
print("D2H contiguous GB/s", bw(lambda: hp.copy_(dc, non_blocking=True), H * W))
print("D2H pitched    GB/s", bw(lambda: hp.copy_(dp, non_blocking=True), H * W))
print("H2D contiguous GB/s", bw(lambda: dc.copy_(hp, non_blocking=True), H * W))
print("H2D pitched    GB/s", bw(lambda: dp.copy_(hp, non_blocking=True), H * W))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
hp2 = torch.empty((H, W), dtype=torch.uint8).pin_memory(); dc2 = torch.empty((H, W), dtype=torch.uint8, device=dev)
def both():
    with torch.cuda.stream(s1): dc.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2): hp2.copy_(dc2, non_blocking=True)
print("bidirectional contiguous GB/s (sum)", bw(both, 2 * H * W))

# pipeline shaped like ovrfsr_apply_host: per stream H2D(in eye) -> [kernel] -> D2H(out eye), round-robin over S streams
IH, IW, OH, OW = 1869, 1683 * 4, 2492, 2244 * 4
def pipeline(S, pitched, nframes=16, busy=False):
    ss = [torch.cuda.Stream() for _ in range(S)]
    hin = [torch.empty((IH, IW), dtype=torch.uint8).pin_memory() for _ in range(S)]
    hout = [torch.empty((OH, OW), dtype=torch.uint8).pin_memory() for _ in range(S)]
    if pitched:
        din = [torch.empty((IH, 6912), dtype=torch.uint8, device=dev)[:, :IW] for _ in range(S)]
        dout = [torch.empty((OH, 9216), dtype=torch.uint8, device=dev)[:, :OW] for _ in range(S)]
    else:
        din = [torch.empty((IH, IW), dtype=torch.uint8, device=dev) for _ in range(S)]
        dout = [torch.empty((OH, OW), dtype=torch.uint8, device=dev) for _ in range(S)]
    def run():
        for f in range(nframes * 2):
            k = f % S
            with torch.cuda.stream(ss[k]):
                din[k].copy_(hin[k], non_blocking=True)
                if busy: torch.cuda._sleep(250000)   # ~0.13 ms of "kernel"
                hout[k].copy_(dout[k], non_blocking=True)
    run(); torch.cuda.synchronize(); t = time.perf_counter()
    run(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    return nframes / dt
for S in (1, 2, 4, 8):
    for pitched in (False, True):
        for busy in (False, True):
            print(f"pipeline streams={S} pitched={pitched} kernel={busy}: {pipeline(S, pitched, busy=busy):.0f} pairs/s")
