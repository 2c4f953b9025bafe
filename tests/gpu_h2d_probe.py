"""H2D bandwidth of pinned / write-combined / portable host memory, 1 and 4 streams; tuning helper.
B200 box (PCIe 5 x16): 55.5 GB/s for every kind."""
import ctypes as C, torch, time
rt = C.CDLL('libcudart.so.12')
torch.cuda.init(); torch.zeros(1, device='cuda')
N = 256 << 20
dev = torch.empty(N, dtype=torch.uint8, device='cuda')
def bw(ptr, label, streams=1):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = N // streams
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(4):
            for i, s in enumerate(ss):
                rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + i * chunk), C.c_void_p(ptr + i * chunk), C.c_size_t(chunk), 1, C.c_void_p(s.cuda_stream))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(label, streams, 'streams', round(4 * N / dt / 1e9, 1), 'GB/s')
for flags, name in ((0, 'pinned'), (4, 'pinned write-combined'), (1, 'pinned portable')):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), flags) == 0
    C.memset(p, 1, N)
    bw(p.value, name, 1); bw(p.value, name, 4)
    rt.cudaFreeHost(p)
t = torch.empty(N, dtype=torch.uint8).pin_memory()
bw(t.data_ptr(), 'torch pin_memory', 1)
