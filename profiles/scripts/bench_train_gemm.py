"""Times the tensor-core products of the training step at the shapes of BASELINE config 3 (524 288 (point, view)
rows): dW += dZ^T X (grad_w_tc_kernel), dIn = dZ W (pack + linear_tc_kernel).  CUDA events, 10 repetitions."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from dynibar_b200._lib import lib, ptr, check, stream
DEV = 'cuda:0'
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
def timeit(fn, n=10):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n
nb = int(lib.dyn_debug_tc_grad_in_scratch_bytes())
scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
for out, width in [(128, 256), (256, 256), (128, 128), (129, 128), (256, 70), (64, 128), (35, 256)]:
  dz = torch.randn(rows, out, device=DEV); x = torch.randn(rows, width, device=DEV)
  dW = torch.zeros(out, width, device=DEV)
  ms = timeit(lambda: check(lib.dyn_debug_tc_grad_w(ptr(dz), out, out, rows, ptr(x), width, width, None, ptr(dW), width, stream())))
  gb = rows * (out + width) * 4 / 1e9
  print('grad_w  out %3d width %3d: %.3f ms  %6.1f TFLOP/s  %6.0f GB/s' % (out, width, ms, 2 * rows * out * width / ms / 1e9, gb / ms * 1e3))
  if width >= 16 and out >= 16:
    W = torch.randn(out, width, device=DEV) * 0.1
    din = torch.empty(rows, width, device=DEV)
    ms = timeit(lambda: check(lib.dyn_debug_tc_grad_in(ptr(dz), out, out, rows, ptr(W), width, width, ptr(din), width, scratch.data_ptr(), nb, stream())))
    print('grad_in out %3d width %3d: %.3f ms  %6.1f TFLOP/s  %6.0f GB/s' % (out, width, ms, 2 * rows * out * width / ms / 1e9, gb / ms * 1e3))
