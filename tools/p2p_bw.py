"""Developer probe: peer-to-peer copy bandwidth between two GPUs of the box (DMA engines), one and several streams, several sizes."""
import torch, time, json
assert torch.cuda.device_count() >= 2
res = {}
for mb in (8, 40, 83, 331, 1024):
    n = mb << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    torch.cuda.set_device(0)
    for nstream in (1, 2, 4):
        streams = [torch.cuda.Stream(device=0) for _ in range(nstream)]
        piece = n // nstream
        def go():
            for j, st in enumerate(streams):
                with torch.cuda.stream(st):
                    b[j * piece:(j + 1) * piece].copy_(a[j * piece:(j + 1) * piece], non_blocking=True)
        go(); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            go()
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = (time.perf_counter() - t0) / reps
        res[f"{mb}MB_x{nstream}"] = round(n / dt / 1e9, 1)
    del a, b
# a 2-D (pitched) copy like the way out of the exchange: 1609 rows of 25 KB at a 206-KB pitch
src = torch.empty(1609 * 25744, dtype=torch.uint8, device="cuda:0")
dst = torch.empty(1609 * 205920, dtype=torch.uint8, device="cuda:1")
d2 = dst.view(1609, 205920)[:, :25744]
s2 = src.view(1609, 25744)
d2.copy_(s2); torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t0 = time.perf_counter()
for _ in range(10):
    d2.copy_(s2, non_blocking=True)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
res["pitched_41MB_torch_copy"] = round(src.numel() / ((time.perf_counter() - t0) / 10) / 1e9, 1)
print(json.dumps(res))
