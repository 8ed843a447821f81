"""Where does the per-chunk cost of a chunked PCIe pipeline come from?  32 MiB each way split into c chunks:
A  H2D copies back to back on one stream          B  A + an event record after every copy
C  H2D on one stream, D2H on another, independent D  H2D -> event -> small kernel -> event -> D2H (three streams)"""
import torch, time
n = 32 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, c, reps=10):
    sz = n // c
    evs = [(torch.cuda.Event(), torch.cuda.Event()) for _ in range(c)]
    def once():
        for k in range(c):
            sl = slice(k * sz, (k + 1) * sz)
            with torch.cuda.stream(s1):
                d_in[sl].copy_(h_in[sl], non_blocking=True)
                if mode in "BD": evs[k][0].record(s1)
            if mode == "C":
                with torch.cuda.stream(s3): h_out[sl].copy_(d_out[sl], non_blocking=True)
            if mode == "D":
                with torch.cuda.stream(s2):
                    s2.wait_event(evs[k][0]); d_out[sl].copy_(d_in[sl]); evs[k][1].record(s2)
                with torch.cuda.stream(s3):
                    s3.wait_event(evs[k][1]); h_out[sl].copy_(d_out[sl], non_blocking=True)
    once(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        once(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for c in (1, 2, 4, 8, 16, 32, 64):
    print("chunks %2d: " % c + "  ".join("%s %.3f ms" % (m, run(m, c) * 1e3) for m in "ABCD"))
