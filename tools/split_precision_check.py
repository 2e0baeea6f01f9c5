"""MEASUREMENT (not an argument): can the Gaussian column-sum pass (a pure GEMM: colsum_s = sum_n x_n . A_s + N c_s,
sparsevi.py:71-72 with model_gaussian.py:4-10) run on the float32-accumulating tensor cores (what tcgen05 offers:
TF32 / BF16 kinds accumulate in float32) with split-precision operands, and still meet the 1e-9 column-sum bar that
the float64 DMMA kernel meets?

Engine for the tensor-core products: cuBLAS through torch.matmul (TF32 mode = the same tensor-core datapath and the same
float32 accumulation a hand-written tcgen05 kind::tf32 kernel would use), so the ACCURACY it measures is the
accuracy a tcgen05 version could reach at best; the time is a lower bound for three tensor-core GEMMs of this shape.

  3xTF32: x = xh + xl, A = Ah + Al (h = TF32-rounded, l = residual rounded to TF32); x.A ~ xh.Ah + xh.Al + xl.Ah
  3xBF16: the same with three-way bf16 splits (6 products, float32 accumulation)
Prints relative errors of the per-column sums against float64 and the bar."""
import json, sys, time
import numpy as np
import torch

N, d, S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000, 200, 512
dev = torch.device('cuda')
g = torch.Generator(device='cpu').manual_seed(0)
x = (torch.randn(N, d, generator=g, dtype=torch.float64) + 1.).to(dev)          # examples/gaussian/main.py:72,82
A = torch.randn(S, d, generator=g, dtype=torch.float64).to(dev)
ref_full = x @ A.T                                                                # float64 N x S
ref = ref_full.sum(dim=0)
# what the bar is relative to (tests: rtol 1e-9, atol 1e-9 * sum |ll|)
scale = ref_full.abs().sum(dim=0).max().item()
del ref_full


def tf32_round(t):                       # round-to-nearest-even to 10 explicit mantissa bits
  i = t.float().contiguous().view(torch.int32)
  i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
  return i.view(torch.float32)


def timed(fn):
  torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
  e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
  return out, e0.elapsed_time(e1)


res = {}
torch.backends.cuda.matmul.allow_tf32 = True
xf = x.float()
xh = tf32_round(xf); xl = tf32_round(xf - xh)
Af = A.float()
Ah = tf32_round(Af); Al = tf32_round(Af - Ah)


def three_tf32():
  acc = (xh @ Ah.T).double().sum(dim=0)            # float32 accumulation inside each product, float64 column reduce
  acc += (xh @ Al.T).double().sum(dim=0)
  acc += (xl @ Ah.T).double().sum(dim=0)
  return acc


def one_tf32():
  return (xh @ Ah.T).double().sum(dim=0)


three_tf32()
out, ms = timed(three_tf32)
res['3xTF32 (fp32 accumulate, f64 column reduce)'] = (out, ms)
out, ms = timed(one_tf32)
res['1xTF32'] = (out, ms)
torch.backends.cuda.matmul.allow_tf32 = False
out, ms = timed(lambda: (xf @ Af.T).double().sum(dim=0))
res['fp32 FFMA GEMM (cuBLAS sgemm)'] = (out, ms)
# bf16 x 3
xb = [None]*3; r = xf.clone()
for i in range(3):
  xb[i] = r.bfloat16(); r = r - xb[i].float()
Ab = [None]*3; r = Af.clone()
for i in range(3):
  Ab[i] = r.bfloat16(); r = r - Ab[i].float()


def bf16x3():
  acc = torch.zeros(S, dtype=torch.float64, device=dev)
  for i in range(3):
    for j in range(3 - i):
      acc += (xb[i] @ Ab[j].T).double().sum(dim=0)   # bf16 products accumulate in float32 and ROUND the output to bf16 in torch
  return acc


out, ms = timed(bf16x3)
res['3xBF16 splits, 6 products (torch rounds each product to bf16: lower bound on accuracy)'] = (out, ms)
out, ms = timed(lambda: (x @ A.T).sum(dim=0))
res['float64 cuBLAS dgemm'] = (out, ms)
print(json.dumps({'N': N, 'd': d, 'S': S, 'bar_rel_to_sum_abs': 1e-9, 'sum_abs_scale': scale}))
for k, (o, ms) in res.items():
  err = (o - ref).abs().max().item()
  print(json.dumps({'variant': k, 'ms': round(ms, 3), 'max_abs_err_colsum': err, 'err_over_sum_abs': err/scale,
                    'max_rel_err': ((o - ref).abs()/ref.abs()).max().item(), 'meets_1e-9_bar': bool(err <= 1e-9*scale)}))
