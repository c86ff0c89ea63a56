"""Time small_exp_batched_kernel (the device exponential!, full matrix) for one and for 128 matrices of size 30."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg as sl
import eu_b200 as eu

eng = eu.get_engine()
out = {}
for nb in (1, 128):
    rng = np.random.default_rng(3)
    H = np.zeros((nb, 30, 30))
    for b in range(nb):  # Hessenberg-like test matrices with ||.||_1 ~ 8 (one squaring), as tH of the C2 operator
        M = np.triu(rng.standard_normal((30, 30)), -1)
        H[b] = M / np.abs(M).sum(0).max() * 8.0
    ref = np.stack([sl.expm(H[b]) for b in range(nb)])
    At0 = torch.from_numpy(np.ascontiguousarray(H.transpose(0, 2, 1))).cuda()
    At = At0.clone()
    eu.exponential_batched_(At)
    err = float(np.abs(At.cpu().numpy().transpose(0, 2, 1) - ref).max() / np.abs(ref).max())
    ts = []
    for _ in range(20):
        At.copy_(At0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eu.exponential_batched_(At); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    out[f"nb{nb}"] = {"us_median": float(np.median(ts)), "us_min": float(np.min(ts)), "max_rel_err_vs_scipy_expm": err}
print(json.dumps(out))
