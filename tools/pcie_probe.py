"""Host<->device copy rates of the box (pinned memory), contiguous and as the 2-D copy of the unpadded segmentation."""
import time, torch, ctypes
from cuda.bindings import runtime as rt
n = 16
disp = torch.empty((n, 1024, 2048), dtype=torch.float32).pin_memory()
seg = torch.empty((n, 256, 21, 256), dtype=torch.int32).pin_memory()
d_disp = torch.empty_like(disp, device="cuda"); d_seg = torch.empty_like(seg, device="cuda")
s = torch.cuda.Stream()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
with torch.cuda.stream(s):
    t = timeit(lambda: d_disp.copy_(disp, non_blocking=True)); print("H2D disparity %.1f GB/s" % (disp.numel() * 4 / t / 1e9))
    t = timeit(lambda: d_seg.copy_(seg, non_blocking=True)); print("H2D padded seg %.1f GB/s (%.2f ms)" % (seg.numel() * 4 / t / 1e9, t * 1e3))
    def c2d():
        err, = rt.cudaMemcpy2DAsync(d_seg.data_ptr(), 1024, seg.data_ptr(), 1024, 512, n * 256 * 21, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, s.cuda_stream)
        assert err == rt.cudaError_t.cudaSuccess, err
    t = timeit(c2d); print("H2D unpadded seg as 2-D copy: %.2f ms (%.1f GB/s of payload)" % (t * 1e3, seg.numel() * 2 / t / 1e9))
    t = timeit(lambda: disp.copy_(d_disp, non_blocking=True)); print("D2H %.1f GB/s" % (disp.numel() * 4 / t / 1e9))
